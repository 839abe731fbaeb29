#!/bin/bash
# ncu --set full (with source) of the tensor-core kernels via the microbench.  usage: tools/gpu_ncu.sh TAG [kernel regex] [cases]
TAG=${1:-n}; KRE=${2:-gcn_tc_kernel}; CASES=${3:-gcn_conv_fwd}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KRE --launch-skip 3 -c 1 -f \
  -o gpurun_out/${TAG} python tools/kernel_bench.py --only $CASES --iters 1 > gpurun_out/${TAG}_ncu.log 2>&1
tail -3 gpurun_out/${TAG}_ncu.log
