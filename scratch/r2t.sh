#!/bin/bash
set -u
mkdir -p gpurun_out
P="python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu --no-c5 --no-extra"
FILO_LIB_PATH=$PWD/scratch/var_lock.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "per_series or golden or edge" 2>&1 | tail -2
for r in 1 2; do for w in c2 c2-counter; do
  FILO_LIB_PATH=$PWD/scratch/base_wp.so timeout 120 $P --workload $w 2>/dev/null | tail -1 > gpurun_out/lock_base_${w}_$r.json
  FILO_LIB_PATH=$PWD/scratch/var_lock.so timeout 120 $P --workload $w 2>/dev/null | tail -1 > gpurun_out/lock_var_${w}_$r.json
done; done
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/lock_*.json")):
    try:
        d = json.load(open(f)); print("%-44s %8.2f ms/step  kernel_ms %.2f frac %.3f" % (f, d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"]))
    except Exception as e: print(f, "unreadable", e)
PY
