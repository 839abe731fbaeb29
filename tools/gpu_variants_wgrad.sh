#!/bin/bash
# wgrad_tc ring / block-shape sweep: "NAME:FLAGS" ...
TAG=$1; shift
mkdir -p gpurun_out; exec > >(tee gpurun_out/${TAG}_sweep.log) 2>&1
for v in "$@"; do
  name=${v%%:*}; flags=${v#*:}
  echo "=== variant $name ($flags)"
  EG_NVCC_EXTRA="$flags" python echoglad_b200/build.py --force > /dev/null || { echo build failed; continue; }
  timeout 300 python -m pytest tests -m gpu -q -x -k "wgrad" 2>&1 | tail -1
  timeout 300 python tools/kernel_bench.py --only wgrad128 2>&1 | grep -v "^{" | tail -1
done
python echoglad_b200/build.py --force > /dev/null
