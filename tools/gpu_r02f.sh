#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x -k "classifier_chain" > gpurun_out/r02f_pytest1.log 2>&1; tail -15 gpurun_out/r02f_pytest1.log
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r02f_pytest.log 2>&1; tail -8 gpurun_out/r02f_pytest.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err; python - <<'PY'
import json
d = json.load(open('gpurun_out/r02f_bench.json'))
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['native_kernel_ms_per_step'])
for k, v in sorted(d['kernels'].items(), key=lambda kv: -kv[1]['ms_per_step']): print(f"  {k:20s} {v['ms_per_step']:.3f} ms x{v['launches_per_step']}")
PY
