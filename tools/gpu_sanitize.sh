#!/bin/bash
# compute-sanitizer over the fused kernel (patch plan: families on S32_n4 / S96_n6, direct child loads on S56_n5) through the
# GCN entry-point parity tests: memcheck (global / shared out-of-bounds) and racecheck (shared-memory hazards).
TAG=${1:-r02zz}
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --kernel-regex kns=gcn_ python -m pytest tests -m gpu -x -q -k "test_gcn_conv_fwd_bwd_entry_points" > gpurun_out/${TAG}_memcheck.log 2>&1
grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/${TAG}_memcheck.log | head -10
timeout 900 compute-sanitizer --tool racecheck --racecheck-report hazard --kernel-regex kns=gcn_patch python -m pytest tests -m gpu -x -q -k "test_gcn_conv_fwd_bwd_entry_points and S32" > gpurun_out/${TAG}_racecheck.log 2>&1
grep -E "RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/${TAG}_racecheck.log | head -10
