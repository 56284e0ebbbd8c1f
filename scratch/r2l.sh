#!/bin/bash
set -u
mkdir -p gpurun_out
P="python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --no-c5"
for r in 1 2; do
  for w in c2 c2-raw; do
    FILO_LIB_PATH=$PWD/scratch/base_wp.so timeout 90 $P --workload $w 2>gpurun_out/w22.err | tail -1 > gpurun_out/w22_base_${w}_$r.json
    FILO_LIB_PATH=$PWD/scratch/var_w22.so timeout 90 $P --workload $w 2>>gpurun_out/w22.err | tail -1 > gpurun_out/w22_var_${w}_$r.json
  done
done
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/w22_*.json")):
    try:
        d = json.load(open(f)); print("%-40s %8.2f ms/step  kernel_ms %.2f frac %.3f" % (f, d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"]))
    except Exception as e: print(f, "unreadable", e)
PY
tail -3 gpurun_out/w22.err
FILO_DEBUG_TIMING=1 timeout 300 python bench.py --no-cpu --no-c5 --steps 3 --warmup 3 --e2e-steps 2 2>gpurun_out/timing.err | tail -1 > gpurun_out/timing.json
grep "scan_series" gpurun_out/timing.err | tail -4
