#!/bin/bash
# fused / per-series counter kernel A/B on one box: base build (scratch/base_wp.so) vs variant builds (scratch/var_*.so), warps-per-CTA sweep
set -u
mkdir -p gpurun_out
P="python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu --no-c5"
run() { # tag workload env...
  local tag=$1 wl=$2; shift 2
  env "$@" timeout 120 $P --workload $wl 2>/dev/null | tail -1 > gpurun_out/ctr_${tag}_${wl}.json
}
for wl in ${WLS:-c5 c2-counter}; do
  run base $wl FILO_LIB_PATH=$PWD/scratch/base_wp.so
  for v in scratch/var_*.so; do t=$(basename $v .so); run $t $wl FILO_LIB_PATH=$PWD/$v; done
  run base_w12 $wl FILO_LIB_PATH=$PWD/scratch/base_wp.so FILO_WP_WARPS=12
  run base_w10 $wl FILO_LIB_PATH=$PWD/scratch/base_wp.so FILO_WP_WARPS=10
done
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/ctr_*.json")):
    try:
        d = json.load(open(f)); print("%-44s %8.2f ms/step  kernel_ms %.2f frac %.3f" % (f, d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"]))
    except Exception as e: print(f, "unreadable", e)
PY
