#!/bin/bash
# ncu full captures of the two classifier middle kernels
mkdir -p gpurun_out
timeout 120 python tools/clf_bench.py 2>&1 | tail -3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"clf_mid_act_bwd_kernel|clf_mid_act_fwd_kernel" --launch-skip 2 -c 2 -f \
  -o gpurun_out/r02af_clf python tools/clf_bench.py --iters 2 > gpurun_out/r02af_ncu.log 2>&1
tail -3 gpurun_out/r02af_ncu.log
