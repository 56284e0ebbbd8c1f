#!/bin/bash
# A/B on one box: scratch/base_wp.so (baseline build) against the current build, interleaved, kernel-only C2 timings
set -u
mkdir -p gpurun_out
W=${WL:-c2}
P="python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --workload $W"
if [ "${FULL:-0}" = "1" ]; then timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3; else timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "per_series or edge or golden or aggregates" 2>&1 | tail -3; fi
for r in 1 2; do
  FILO_LIB_PATH=$PWD/scratch/base_wp.so timeout 90 $P 2>/dev/null | tail -1 > gpurun_out/ab_base_$r.json
  timeout 90 $P 2>gpurun_out/ab.err | tail -1 > gpurun_out/ab_new_$r.json
  if [ -n "${VARIANT:-}" ]; then env $VARIANT timeout 90 $P 2>/dev/null | tail -1 > gpurun_out/ab_var_$r.json; fi
done
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/ab_*.json")):
    try:
        d = json.load(open(f)); print("%-32s %8.2f ms/step  kernel_ms %.2f frac %.3f" % (f, d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"]))
    except Exception as e: print(f, "unreadable", e)
PY
tail -3 gpurun_out/ab.err
if [ "${NCU:-1}" = "1" ]; then
rm -f gpurun_out/src_$W.ncu-rep
timeout 300 ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --section MemoryWorkloadAnalysis_Tables --section LaunchStats --section Occupancy --section SpeedOfLight --import-source on --clock-control none -k regex:${KREGEX:-scan_wp} -c 1 -o gpurun_out/src_$W python bench.py --workload $W --series 2960000 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/src_$W.log 2>&1
tail -1 gpurun_out/src_$W.log
fi
