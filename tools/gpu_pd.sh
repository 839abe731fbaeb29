#!/bin/bash
# A/B timing of prebuilt library variants (tools/build_variants.sh): forward / backward / linear128 microbench on the
# default graph and the forward on the main-only graph.  usage: tools/gpu_pd.sh TAG name...
TAG=$1; shift
mkdir -p gpurun_out
for name in "$@"; do
  L=echoglad_b200/variants/libeg_${name}.so
  a=$(EG_LIB_PATH=$L timeout 40 python tools/kernel_bench.py --only ${PD_ONLY:-gcn_conv_fwd} 2>&1 | grep -v "^{" | grep -E "gcn_conv|linear128" | awk '{printf "%s %s  ", $1, $2}')
  b=$(EG_LIB_PATH=$L timeout 40 python tools/kernel_bench.py --only gcn_conv_fwd --main-only --batch 92 2>&1 | grep -v "^{" | grep gcn_conv_fwd | awk '{print $2}')
  echo "$name default $a ms  main-only(b92) $b ms" | tee -a gpurun_out/${TAG}_pd.log
done
