#!/bin/bash
# knock-out variants on the main-only graph (all tiles are lattice tiles), timing only
TAG=$1; shift
for v in "$@"; do
  name=${v%%:*}; flags=${v#*:}
  echo "=== variant $name ($flags)"
  EG_NVCC_EXTRA="$flags" python echoglad_b200/build.py --force > /dev/null || { echo build failed; continue; }
  timeout 300 python tools/kernel_bench.py --main-only --batch 92 --only gcn_conv_fwd 2>&1 | grep -v "^{" | tail -2
done
python echoglad_b200/build.py --force > /dev/null
