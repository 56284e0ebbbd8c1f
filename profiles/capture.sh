#!/bin/bash
# Round-1 profile capture (run on the GPU box through gpurun; outputs land in gpurun_out/, summaries are copied to profiles/).
#   1. launch list of the bench command (per-launch durations, serialised + cold cache: compare SHARES, not absolutes)
#   2. one full-metric capture of the dominant kernel at the bench size (DRAM traffic for roofline.traffic)
set -e
W=${1:-c2}
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$W.csv \
  python bench.py --workload $W --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/launches_$W.log 2>&1 || true
timeout 300 ncu --set full --clock-control none --import-source on -k regex:scan_tile -c 1 -o gpurun_out/full_$W \
  python bench.py --workload $W --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/full_$W.log 2>&1 || true
tail -2 gpurun_out/full_$W.log
