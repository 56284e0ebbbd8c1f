#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 120 python bench.py --workload c1 --steps 20 --warmup 5 2>gpurun_out/r2j_c1.err | tail -1 > gpurun_out/r2j_c1.json; tail -2 gpurun_out/r2j_c1.err
timeout 200 python bench.py --workload c4 --steps 6 --warmup 3 2>gpurun_out/r2j_c4.err | tail -1 > gpurun_out/r2j_c4.json; tail -2 gpurun_out/r2j_c4.err
timeout 400 python bench.py --no-c5 --no-cpu 2>gpurun_out/r2j_c2.err | tail -1 > gpurun_out/r2j_c2.json; tail -2 gpurun_out/r2j_c2.err
python - <<PY
import json
for w in ("c1", "c4", "c2"):
    try:
        d = json.load(open("gpurun_out/r2j_%s.json" % w))
        print(w, d["ms_per_step"], d["roofline"]["frac"], "e2e", (d.get("e2e") or {}).get("value"), "resident", (d.get("e2e_resident") or {}), "parity", d.get("parity_check"))
    except Exception as e: print(w, "unreadable", e)
PY
