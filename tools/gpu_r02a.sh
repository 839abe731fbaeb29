#!/bin/bash
# r02a: full GPU parity tier (incl. the new real-size tests) + mbarrier suspend-hint sweep of the tcgen05 kernels
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv,noheader
free -g | head -2
timeout 2400 python -m pytest tests -m gpu -q --timeout 1200 > gpurun_out/r02a_pytest.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/r02a_pytest.log
tail -15 gpurun_out/r02a_pytest.log
for v in sus20000 sus2000 sus200 sus20 spin; do
  echo "=== $v"
  EG_LIB_PATH=$PWD/echoglad_b200/variants/libeg_$v.so timeout 300 python tools/kernel_bench.py --only gcn_conv_fwd,gcn_conv_bwd,linear128,wgrad128 2>&1 | grep -v "^{" | tee gpurun_out/r02a_kb_$v.log
done
