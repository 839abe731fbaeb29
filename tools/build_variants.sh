#!/bin/bash
# Prebuild compile-time variants of the library HERE (nvcc cross-compiles without a GPU) so that a gpurun call only
# measures.  usage: tools/build_variants.sh "name:FLAGS" ...   -> echoglad_b200/variants/libeg_<name>.so
# run one on the GPU box with: EG_LIB_PATH=echoglad_b200/variants/libeg_<name>.so python tools/kernel_bench.py ...
mkdir -p echoglad_b200/variants
for v in "$@"; do
  name=${v%%:*}; flags=${v#*:}
  EG_LIB_OUT=$PWD/echoglad_b200/variants/libeg_${name}.so EG_NVCC_EXTRA="$flags" python echoglad_b200/build.py --force > /dev/null \
    && echo "built $name [$flags]" || echo "FAILED $name"
  rm -rf echoglad_b200/variants/libeg_${name}.so.obj
done
