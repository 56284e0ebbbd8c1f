#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
P="python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu --no-c5 --no-extra"
for w in c3 c3-const c5 c2-counter; do timeout 120 $P --workload $w 2>gpurun_out/r2m.err | tail -1 > gpurun_out/r2m_$w.json; done
FILO_DEBUG_TIMING=1 timeout 300 python bench.py --no-cpu --no-c5 --no-extra --steps 3 --warmup 3 --e2e-steps 2 2>gpurun_out/timing2.err | tail -1 > gpurun_out/r2m_c2.json
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/r2m_*.json")):
    try:
        d = json.load(open(f)); print("%-36s %8.2f ms/step  kernel_ms %.2f frac %.3f e2e %s" % (f, d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], (d.get("e2e") or {}).get("value")))
    except Exception as e: print(f, "unreadable", e)
PY
tail -2 gpurun_out/r2m.err; grep scan_series gpurun_out/timing2.err | tail -2
