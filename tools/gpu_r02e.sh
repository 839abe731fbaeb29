#!/bin/bash
mkdir -p gpurun_out
for v in pf1 pf2 pf3 pf5; do
  echo "=== $v"
  EG_LIB_PATH=$PWD/echoglad_b200/variants/libeg_$v.so timeout 300 python tools/kernel_bench.py --only gcn_conv_fwd,gcn_conv_bwd 2>&1 | grep -v "^{" | tee gpurun_out/r02e_kb_$v.log
  EG_LIB_PATH=$PWD/echoglad_b200/variants/libeg_$v.so timeout 300 python tools/kernel_bench.py --main-only --batch 92 --only gcn_conv_fwd 2>&1 | grep -v "^{" | tee -a gpurun_out/r02e_kb_$v.log
done
EG_LIB_PATH=$PWD/echoglad_b200/variants/libeg_pf2.so timeout 600 python -m pytest tests -m gpu -q -x -k "gcn_conv or fused_gcn" 2>&1 | tail -3
