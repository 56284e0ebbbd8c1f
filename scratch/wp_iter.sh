#!/bin/bash
# one iteration of the wp-kernel loop: GPU parity (fast subset unless FULL=1), timing, source-level instruction counters of the wp kernel
set -u
mkdir -p gpurun_out
W=${WL:-c2}
K=${KREGEX:-scan_wp}
P="python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu --workload $W"
if [ "${FULL:-0}" = "1" ]; then timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3; else timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "per_series or edge or golden or aggregates" 2>&1 | tail -3; fi
timeout 90 $P 2>gpurun_out/it.err | tail -1 > gpurun_out/it_$W.json
python - <<PY
import json
d = json.load(open("gpurun_out/it_$W.json")); print("$W %8.2f ms/step  frac %.3f kernel_ms %.2f" % (d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["kernel_ms"]))
PY
tail -3 gpurun_out/it.err
rm -f gpurun_out/src_$W.ncu-rep
timeout 300 ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --section MemoryWorkloadAnalysis_Tables --section LaunchStats --section Occupancy --section SpeedOfLight --import-source on --clock-control none -k regex:$K -c 1 -o gpurun_out/src_$W python bench.py --workload $W --series 2960000 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/src_$W.log 2>&1
tail -2 gpurun_out/src_$W.log; ls -la gpurun_out/src_$W.ncu-rep
