#!/bin/bash
# quick GPU check: selected tests + kernel microbench.  usage: tools/gpu_quick.sh TAG "pytest -k expr" "kernel_bench --only list"
TAG=${1:-q}; KEXPR=${2:-}; ONLY=${3:-}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q ${KEXPR:+-k "$KEXPR"} > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?"
tail -15 gpurun_out/${TAG}_pytest.log
timeout 300 python tools/kernel_bench.py ${ONLY:+--only $ONLY} > gpurun_out/${TAG}_kernel_bench.log 2>&1; tail -12 gpurun_out/${TAG}_kernel_bench.log
