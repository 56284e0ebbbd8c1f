#!/bin/bash
# First GPU call of the next round: everything that was changed after the last measurement of round 1, in one short run.
#   scratch/gpuretry.sh 400 'bash scratch/round2_measure.sh'
# Writes gpurun_out/r2_*.json / r2_*.txt.  Roughly 3 minutes of box time.
set -u
mkdir -p gpurun_out
P="python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu"
# 1. parity first (the post-measurement kernel changes have only run on the CPU emulator)
timeout 150 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r2_pytest.txt; cat gpurun_out/r2_pytest.txt
# 2. C2 tile kernel: junction blocks on / off (same box, same build)
timeout 60 $P 2>/dev/null | tail -1 > gpurun_out/r2_c2_junction_on.json
FILO_TILE_JUNCTION=0 timeout 60 $P 2>/dev/null | tail -1 > gpurun_out/r2_c2_junction_off.json
# 2b. experimental per-warp decode (no mid-decode barrier, 16-byte row stores): parity under the switch, then its timing
FILO_TILE_WARPDEC=1 timeout 120 python -m pytest tests/test_gpu_parity.py -q -x -k "per_series_bit_exact or golden or edge or aggregates" 2>&1 | tail -2 > gpurun_out/r2_pytest_warpdec.txt; cat gpurun_out/r2_pytest_warpdec.txt
FILO_TILE_WARPDEC=1 timeout 60 $P 2>/dev/null | tail -1 > gpurun_out/r2_c2_warpdec.json
for w in c2-counter c5; do FILO_TILE_WARPDEC=1 timeout 90 $P --workload $w 2>/dev/null | tail -1 > gpurun_out/r2_${w}_warpdec.json; done
# 3. the other workloads, kernel only
for w in c2-counter c2-raw c3-const c5; do timeout 90 $P --workload $w 2>/dev/null | tail -1 > gpurun_out/r2_$w.json; done
# 4. C4 histogram kernels A/B
timeout 60 python bench.py --workload c4 --steps 6 --warmup 3 --no-e2e --no-cpu 2>/dev/null | tail -1 > gpurun_out/r2_c4_v2.json
FILO_HIST_V2=0 timeout 90 python bench.py --workload c4 --steps 4 --warmup 2 --no-e2e --no-cpu 2>/dev/null | tail -1 > gpurun_out/r2_c4_v1.json
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_*.json")):
    try:
        d = json.load(open(f)); print("%-40s %8.2f ms/step  frac %.3f" % (f, d["ms_per_step"], d["roofline"]["frac"]))
    except Exception as e:
        print(f, "unreadable:", e)
PY
# 5. per-phase cycle profiles (need scratch/libfilo_b200_prof.so: FILO_NVCC_EXTRA="-DFILO_HIST_PROF -DFILO_TILE_PROF" FILO_BUILD_OUT=scratch/libfilo_b200_prof.so python -m filodb_b200.build --force)
if [ -f scratch/libfilo_b200_prof.so ]; then
  timeout 120 python scratch/hist_prof.py > gpurun_out/r2_hist_phases.txt 2>&1; cat gpurun_out/r2_hist_phases.txt
  for w in c2 c2-counter; do timeout 90 python scratch/tile_prof.py $w 2>/dev/null | tail -12 > gpurun_out/r2_tile_phases_$w.txt; cat gpurun_out/r2_tile_phases_$w.txt; done
fi
