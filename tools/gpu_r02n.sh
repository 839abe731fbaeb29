#!/bin/bash
# ncu full captures (source-level stall samples) of the patch-mode fused kernel: main-only graph and default graph
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gcn_patch_kernel --launch-skip 3 -c 1 -f \
  -o gpurun_out/r02n_main python tools/kernel_bench.py --main-only --batch 92 --only gcn_conv_fwd --iters 1 > gpurun_out/r02n_ncu_main.log 2>&1
tail -2 gpurun_out/r02n_ncu_main.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gcn_patch_kernel --launch-skip 3 -c 1 -f \
  -o gpurun_out/r02n_full python tools/kernel_bench.py --only gcn_conv_fwd --iters 1 > gpurun_out/r02n_ncu_full.log 2>&1
tail -2 gpurun_out/r02n_ncu_full.log
ls -la gpurun_out/*.ncu-rep
