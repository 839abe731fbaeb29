#!/bin/bash
# r02k: full GPU parity tier, C1' (coordinate graph) timing, bench line, ncu launch list, ncu full captures
TAG=r02k
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
timeout 600 python tools/coord_bench.py > gpurun_out/${TAG}_coord_bench.txt 2>&1; tail -3 gpurun_out/${TAG}_coord_bench.txt
timeout 300 python tools/kernel_bench.py > gpurun_out/${TAG}_kernel_bench.log 2>&1; tail -12 gpurun_out/${TAG}_kernel_bench.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench.json | cut -c1-600
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 3 --ncu-range --no-cpu-baseline > gpurun_out/${TAG}_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gcn_tc_kernel --launch-skip 3 -c 1 -f \
  -o gpurun_out/${TAG}_gcn_tc python tools/kernel_bench.py --only gcn_conv_fwd --iters 1 > gpurun_out/${TAG}_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gcn_tc_kernel|wgrad_tc_kernel" --launch-skip 6 -c 2 -f \
  -o gpurun_out/${TAG}_bwd python tools/kernel_bench.py --only gcn_conv_bwd --iters 1 > gpurun_out/${TAG}_ncu3.log 2>&1
{ echo "== configs[1]: use_main_graph_only, batch 32"; timeout 300 python tools/kernel_bench.py --main-only --batch 32 --only gcn_conv_fwd,gcn_conv_bwd,aggregate | grep -v "^{";
  echo "== configs[3]: frame 448, 8 aux levels, batch 16"; timeout 300 python tools/kernel_bench.py --frame 448 --naux 8 --batch 16 --only gcn_conv_fwd,gcn_conv_bwd,aggregate | grep -v "^{"; } > gpurun_out/${TAG}_configs_c2_c4.txt 2>&1
cat gpurun_out/${TAG}_configs_c2_c4.txt
ls -la gpurun_out | grep ${TAG}
