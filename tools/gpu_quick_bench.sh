#!/bin/bash
# usage: tools/gpu_quick_bench.sh TAG [pytest -k expression]
TAG=${1:-rXX}
mkdir -p gpurun_out
if [ -n "$2" ]; then timeout 1200 python -m pytest tests -m gpu -q -x -k "$2" 2>&1 | tail -8; fi
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python - "$TAG" <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/{sys.argv[1]}_bench.json"))
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["native_kernel_ms_per_step"], d["gpu_launches"], d["e2e"]["value"])
for k, v in sorted(d["kernels"].items(), key=lambda kv: -kv[1]["ms_per_step"]):
    print(f"  {k:22s} {v['ms_per_step']:.3f} ms x{v['launches_per_step']}")
PY
