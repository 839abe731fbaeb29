#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gcn_tc_kernel --launch-skip 3 -c 1 -f \
  -o gpurun_out/r02d_main python tools/kernel_bench.py --main-only --batch 92 --only gcn_conv_fwd --iters 1 > gpurun_out/r02d_ncu_main.log 2>&1
tail -2 gpurun_out/r02d_ncu_main.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gcn_tc_kernel --launch-skip 3 -c 1 -f \
  -o gpurun_out/r02d_full python tools/kernel_bench.py --only gcn_conv_fwd --iters 1 > gpurun_out/r02d_ncu_full.log 2>&1
tail -2 gpurun_out/r02d_ncu_full.log
ls -la gpurun_out/*.ncu-rep
