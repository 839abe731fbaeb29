#!/bin/bash
mkdir -p gpurun_out
EG_LIB_PATH=$PWD/echoglad_b200/variants/libeg_timing.so timeout 300 python tools/kernel_bench.py --main-only --batch 92 --only gcn_conv_fwd 2>&1 | grep -v "^{" | tee gpurun_out/r02c_kb_timing.log
