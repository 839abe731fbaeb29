#!/bin/bash
# r02 final state: full GPU parity tier, microbench, bench line (+ reference arm), ncu launch list, ncu full captures of
# the patch-mode fused kernel (forward, backward) and the weight gradient, configs[1] / configs[3] microbench.
TAG=${1:-r02zz}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
timeout 300 python tools/kernel_bench.py > gpurun_out/${TAG}_kernel_bench.log 2>&1; tail -14 gpurun_out/${TAG}_kernel_bench.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; cut -c1-700 gpurun_out/${TAG}_bench.json
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; cut -c1-400 gpurun_out/${TAG}_bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 3 --ncu-range --no-cpu-baseline > gpurun_out/${TAG}_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gcn_patch_kernel|gcn_tc_kernel" --launch-skip 3 -c 1 -f \
  -o gpurun_out/${TAG}_gcn_tc python tools/kernel_bench.py --only gcn_conv_fwd --iters 1 > gpurun_out/${TAG}_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gcn_patch_kernel|gcn_tc_kernel|wgrad_tc_kernel" --launch-skip 6 -c 2 -f \
  -o gpurun_out/${TAG}_bwd python tools/kernel_bench.py --only gcn_conv_bwd --iters 1 > gpurun_out/${TAG}_ncu3.log 2>&1
{ echo "== configs[1]: use_main_graph_only, batch 32"; timeout 300 python tools/kernel_bench.py --main-only --batch 32 --only gcn_conv_fwd,gcn_conv_bwd,aggregate | grep -v "^{";
  echo "== main graph only at the bytes of configs[2] (batch 92)"; timeout 300 python tools/kernel_bench.py --main-only --batch 92 --only gcn_conv_fwd,gcn_conv_bwd | grep -v "^{";
  echo "== configs[3]: frame 448, 8 aux levels, batch 16"; timeout 300 python tools/kernel_bench.py --frame 448 --naux 8 --batch 16 --only gcn_conv_fwd,gcn_conv_bwd,aggregate | grep -v "^{";
  echo "== use_connection_nodes (hub rows with 4..16384 neighbours: gather plan, CSR rows), batch 64"; timeout 300 python tools/kernel_bench.py --conn --only gcn_conv_fwd,gcn_conv_bwd,aggregate | grep -v "^{";
  echo "== grid-diagonal lattices (8-neighbour rows: gather plan), batch 64"; timeout 300 python tools/kernel_bench.py --diag --only gcn_conv_fwd,gcn_conv_bwd,aggregate | grep -v "^{";
  echo "== gather plan (EG_GCN_PLAN=gather: cp.async row copies, per-row slot plan) on configs[2]"; EG_GCN_PLAN=gather timeout 300 python tools/kernel_bench.py --only gcn_conv_fwd,gcn_conv_bwd | grep -v "^{"; } > gpurun_out/${TAG}_configs_c2_c4.txt 2>&1
cat gpurun_out/${TAG}_configs_c2_c4.txt
ls -la gpurun_out | grep ${TAG}
