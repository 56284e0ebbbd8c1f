#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 300 python bench.py --workload c4 --steps 5 --warmup 3 2>gpurun_out/r2o.err | tail -1 > gpurun_out/r2o_c4.json; tail -3 gpurun_out/r2o.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2o_c4.json"))
print("c4", d["ms_per_step"], d["roofline"]["frac"], d["config"]["table_gen_s"], "e2e", (d.get("e2e") or {}).get("value"), "cpu", (d.get("cpu_baseline") or {}).get("value"), "parity", d.get("parity_check"))
PY
