#!/bin/bash
# A/B on the GPU box: for each "NAME:ENV:FLAGS" argument rebuild the library with EG_NVCC_EXTRA=FLAGS, export ENV
# (space-separated VAR=value pairs), optionally run the tensor-core parity tests (names starting with "t") and the
# kernel microbench.  usage: tools/gpu_ab.sh TAG "t_new::" "old:EG_TILE_ORDER=coarse-first:" "nopipe::-DEG_NO_CHILD_PIPE"
TAG=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  name=${v%%:*}; rest=${v#*:}; envs=${rest%%:*}; flags=${rest#*:}
  echo "=== variant $name env[$envs] flags[$flags]"
  EG_NVCC_EXTRA="$flags" python echoglad_b200/build.py --force > /dev/null || { echo build failed; continue; }
  if [[ $name == t* ]]; then
    env $envs timeout 900 python -m pytest tests -m gpu -x -q -k "linear128 or gcn_conv or aggregate_full or default_yml or unet_variant or fused_gcn or deeper" > gpurun_out/${TAG}_${name}_pytest.log 2>&1
    tail -3 gpurun_out/${TAG}_${name}_pytest.log
  fi
  env $envs timeout 300 python tools/kernel_bench.py --only ${KB_ONLY:-gcn_conv_fwd,gcn_conv_bwd} > gpurun_out/${TAG}_${name}_kb.log 2>&1; grep -v "^{" gpurun_out/${TAG}_${name}_kb.log | tail -4
done
python echoglad_b200/build.py --force > /dev/null
