#!/bin/bash
# knock-outs of the backward launch of the fused kernel (WRONG results, timing only): side-output stores, addend loads,
# epilogue stores
mkdir -p gpurun_out
for rep in 1 2; do
for v in base noagg nores both nostore; do
  a=$(EG_LIB_PATH=echoglad_b200/variants/libeg_$v.so timeout 120 python tools/kernel_bench.py --only gcn_bwd_nowgrad,gcn_layer_eval,gcn_conv_fwd 2>&1 | grep -v "^{" | awk '{printf "%s %s  ", $1, $2}')
  echo "$v $a" | tee -a gpurun_out/r02ae_pd.log
done; done
