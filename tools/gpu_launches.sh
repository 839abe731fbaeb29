#!/bin/bash
# ncu launch list of one steady-state training step (aggregated by profiles/summarize.py)
TAG=${1:-r02end}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 3 --ncu-range --no-cpu-baseline > gpurun_out/${TAG}_ncu1.log 2>&1
tail -2 gpurun_out/${TAG}_ncu1.log
