#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}; TAG=${2:-r02j}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err
tail -3 gpurun_out/${TAG}_bench_n$N.err
python - <<PY
import json
d = json.load(open('gpurun_out/${TAG}_bench_n$N.json'))
print("weak", d['n_gpus'], d['value'], d['ms_per_step'], "e2e", d['e2e']['value'])
print("strong", d['strong_scaling'] and {k: d['strong_scaling'][k] for k in ('global_batch','batch_per_gpu','ms_per_step','frames_per_s','native_kernel_ms_per_step_rank0')})
PY
