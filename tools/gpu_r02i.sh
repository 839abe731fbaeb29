#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x -k "classifier_chain or module or hot_path or engine_style" > gpurun_out/r02i_pytest1.log 2>&1; tail -5 gpurun_out/r02i_pytest1.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r02i_bench.json 2> gpurun_out/r02i_bench.err; python - <<'PY'
import json
d = json.load(open('gpurun_out/r02i_bench.json'))
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['native_kernel_ms_per_step'])
for k, v in sorted(d['kernels'].items(), key=lambda kv: -kv[1]['ms_per_step']): print(f"  {k:20s} {v['ms_per_step']:.3f} ms x{v['launches_per_step']}")
PY
