#!/bin/bash
# One gpurun call: GPU parity tests, per-kernel microbench, bench line, ncu launch list, ncu full capture.
# usage: tools/gpu_round.sh TAG   (outputs under gpurun_out/TAG_*)
TAG=${1:-rXX}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
timeout 300 python tools/kernel_bench.py > gpurun_out/${TAG}_kernel_bench.log 2>&1; tail -12 gpurun_out/${TAG}_kernel_bench.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 3 --ncu-range > gpurun_out/${TAG}_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gcn_tc_kernel -c 2 -f \
  -o gpurun_out/${TAG}_gcn_tc python tools/kernel_bench.py --only gcn_conv_fwd,linear128 --iters 1 > gpurun_out/${TAG}_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gcn_tc_kernel|wgrad_tc_kernel" --launch-skip 6 -c 2 -f \
  -o gpurun_out/${TAG}_bwd python tools/kernel_bench.py --only gcn_conv_bwd --iters 1 > gpurun_out/${TAG}_ncu3.log 2>&1
ls -la gpurun_out | tail -12
# BASELINE configs[1] (main graph only, batch 32) and configs[3] (448 px / 8 aux levels, batch 16): per-entry-point microbench
{ echo "== configs[1]: use_main_graph_only, batch 32"; timeout 300 python tools/kernel_bench.py --main-only --batch 32 --only gcn_conv_fwd,gcn_conv_bwd,aggregate | grep -v "^{";
  echo "== configs[3]: frame 448, 8 aux levels, batch 16"; timeout 300 python tools/kernel_bench.py --frame 448 --naux 8 --batch 16 --only gcn_conv_fwd,gcn_conv_bwd,aggregate | grep -v "^{"; } > gpurun_out/${TAG}_configs_c2_c4.txt 2>&1
cat gpurun_out/${TAG}_configs_c2_c4.txt
