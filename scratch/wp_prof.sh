#!/bin/bash
# source-level counters of one kernel for one workload: W=<workload> K=<kernel regex>
set -u
mkdir -p gpurun_out
for spec in ${SPECS:-c2-counter:scan_wp_ctr c5:scan_wp_ctr}; do
  W=${spec%%:*}; K=${spec##*:}
  rm -f gpurun_out/src_$W.ncu-rep
  timeout 300 ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --section MemoryWorkloadAnalysis_Tables --section LaunchStats --section Occupancy --section SpeedOfLight --import-source on --clock-control none -k regex:$K -c 1 -o gpurun_out/src_$W python bench.py --workload $W --series 2960000 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/src_$W.log 2>&1
  tail -1 gpurun_out/src_$W.log
done
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_c5.csv python bench.py --workload c5 --series 2960000 --steps 2 --warmup 1 --no-e2e --no-cpu > /dev/null 2>&1
grep -v "^==" gpurun_out/launches_c5.csv | awk -F'","' '{print $5, $NF}' | tail -14
