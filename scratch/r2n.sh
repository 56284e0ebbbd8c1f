#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
P="python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu --no-c5 --no-extra"
for w in c3 c3-const c5 c2-counter; do timeout 120 $P --workload $w 2>gpurun_out/r2n.err | tail -1 > gpurun_out/r2n_$w.json; done
for v in base_wp var_h2b10 var_h2b20; do FILO_LIB_PATH=$PWD/scratch/$v.so timeout 150 python bench.py --workload c4 --steps 5 --warmup 2 --no-e2e --no-cpu 2>>gpurun_out/r2n.err | tail -1 > gpurun_out/r2n_c4_$v.json; done
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/r2n_*.json")):
    try:
        d = json.load(open(f)); print("%-40s %8.2f ms/step  kernel_ms %.2f frac %.3f" % (f, d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"]))
    except Exception as e: print(f, "unreadable", e)
PY
tail -2 gpurun_out/r2n.err
