#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 -k "config3 or data_batch or node_labels or engine_style" > gpurun_out/r02b_pytest.log 2>&1; tail -5 gpurun_out/r02b_pytest.log
EG_LIB_PATH=$PWD/echoglad_b200/variants/libeg_timing.so timeout 300 python tools/kernel_bench.py --only gcn_conv_fwd,gcn_conv_bwd,linear128 2>&1 | grep -v "^{" | tee gpurun_out/r02b_kb_timing.log
EG_LIB_PATH=$PWD/echoglad_b200/variants/libeg_timing.so timeout 300 python tools/kernel_bench.py --main-only --batch 92 --only gcn_conv_fwd 2>&1 | grep -v "^{" | tee -a gpurun_out/r02b_kb_timing.log
