#!/bin/bash
# v4 warp-pipeline kernel: GPU parity, then C2 timings against the tile kernel (FILO_KERNEL=v3) and over warps per CTA
set -u
mkdir -p gpurun_out
P="python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu"
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/wp_pytest.txt; cat gpurun_out/wp_pytest.txt
timeout 90 $P 2>gpurun_out/wp_c2.err | tail -1 > gpurun_out/wp_c2.json
FILO_KERNEL=v3 timeout 90 $P 2>/dev/null | tail -1 > gpurun_out/wp_c2_v3.json
for w in 8 12 14; do FILO_WP_WARPS=$w timeout 90 $P 2>/dev/null | tail -1 > gpurun_out/wp_c2_w$w.json; done
timeout 90 $P --workload c2-raw 2>/dev/null | tail -1 > gpurun_out/wp_c2-raw.json
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/wp_*.json")):
    try:
        d = json.load(open(f)); print("%-40s %8.2f ms/step  frac %.3f" % (f, d["ms_per_step"], d["roofline"]["frac"]))
    except Exception as e:
        print(f, "unreadable:", e)
PY
tail -5 gpurun_out/wp_c2.err
