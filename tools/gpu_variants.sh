#!/bin/bash
# A/B of compile-time variants on the GPU box: for each "NAME:FLAGS" argument rebuild the library with
# EG_NVCC_EXTRA=FLAGS, run the tensor-core parity tests and the kernel microbench.
# usage: tools/gpu_variants.sh TAG "base:" "trunc:-DEG_TF32_TRUNC" ...
TAG=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  name=${v%%:*}; flags=${v#*:}
  echo "=== variant $name ($flags)"
  EG_NVCC_EXTRA="$flags" python echoglad_b200/build.py --force > /dev/null || { echo build failed; continue; }
  [[ $name == dbg* ]] || timeout 600 python -m pytest tests -m gpu -x -q -k "linear128 or gcn_conv or aggregate_full or default_yml or unet_variant" > gpurun_out/${TAG}_${name}_pytest.log 2>&1; [[ $name == dbg* ]] || tail -3 gpurun_out/${TAG}_${name}_pytest.log
  timeout 300 python tools/kernel_bench.py --only gcn_conv_fwd,gcn_conv_bwd,linear128 > gpurun_out/${TAG}_${name}_kb.log 2>&1; grep -v "^{" gpurun_out/${TAG}_${name}_kb.log | tail -4
done
python echoglad_b200/build.py --force > /dev/null
