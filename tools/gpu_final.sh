#!/bin/bash
# final check at HEAD: full GPU parity tier, smoke, bench line
TAG=${1:-r02end}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; cut -c1-600 gpurun_out/${TAG}_bench.json
timeout 300 python tools/kernel_bench.py > gpurun_out/${TAG}_kernel_bench.txt 2>&1; grep -v "^{" gpurun_out/${TAG}_kernel_bench.txt | tail -16
