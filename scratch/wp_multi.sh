#!/bin/bash
# several workloads, new build vs FILO_KERNEL=v3 (tile kernel) on one box
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for W in ${WLS:-c2-counter c5 c3-const c3}; do
  P="python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu --workload $W"
  timeout 120 $P 2>gpurun_out/multi.err | tail -1 > gpurun_out/multi_${W}_new.json
  FILO_KERNEL=v3 timeout 120 $P 2>/dev/null | tail -1 > gpurun_out/multi_${W}_v3.json
done
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/multi_*.json")):
    try:
        d = json.load(open(f)); print("%-40s %8.2f ms/step  kernel_ms %.2f frac %.3f" % (f, d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"]))
    except Exception as e: print(f, "unreadable", e)
PY
tail -3 gpurun_out/multi.err
