#!/bin/bash
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool racecheck --racecheck-report all --kernel-name kns=clf_mid_act_bwd python tools/dbg/clf_chain_dbg.py 9472 > gpurun_out/r02g_race.log 2>&1; grep -v "^=========     at\|^=========         in\|Saved host" gpurun_out/r02g_race.log | head -60
timeout 600 compute-sanitizer --tool memcheck --kernel-name kns=clf_ python tools/dbg/clf_chain_dbg.py 9472 > gpurun_out/r02g_mem.log 2>&1; tail -5 gpurun_out/r02g_mem.log
