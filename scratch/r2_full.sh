#!/bin/bash
# Round-2 full measurement of HEAD on one box: GPU parity, the default bench line (what the driver runs), the reference arm, every
# other workload kernel-only, the ncu launch list of the bench command and one --set full capture of the dominant kernels.
#   scratch/gpuretry.sh 1500 'bash scratch/r2_full.sh'
set -u
mkdir -p gpurun_out
T=${TAG:-r2h}
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/${T}_pytest.txt; cat gpurun_out/${T}_pytest.txt
timeout 500 python bench.py 2>gpurun_out/${T}_default.err | tail -1 > gpurun_out/${T}_default.json
tail -2 gpurun_out/${T}_default.err
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/${T}_reference.json
P="python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu --no-c5"
for w in c2-counter c2-raw c3-const c3 c5; do timeout 120 $P --workload $w 2>/dev/null | tail -1 > gpurun_out/${T}_$w.json; done
timeout 120 python bench.py --workload c4 --steps 6 --warmup 3 --no-e2e --no-cpu 2>/dev/null | tail -1 > gpurun_out/${T}_c4.json
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${T}_*.json")):
    try:
        d = json.load(open(f)); r = d.get("roofline", {})
        print("%-36s %9.2f ms/step  kernel_ms %8.2f frac %.3f  e2e %s  parity %s" % (f, d["ms_per_step"], r.get("kernel_ms", 0), r.get("frac", 0), (d.get("e2e") or {}).get("value"), d.get("parity_check")))
        if "c5" in d: print("   c5:", d["c5"].get("ms_per_step"), d["c5"].get("roofline", {}).get("frac"), d["c5"].get("error"))
        if "cpu_baseline" in d: print("   cpu:", d["cpu_baseline"].get("value"), d["cpu_baseline"].get("cores"))
    except Exception as e: print(f, "unreadable", e)
PY
# launch list of the bench command (serialised, cold cache: shares, not absolutes)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_c2.csv \
  python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/${T}_launches_c2.log 2>&1
# full captures: the SUM-class kernel on C2, the counter kernel on C5 (fused) -- 2.96 M series = 148 x 20,000
for spec in c2:scan_wp_sum c5:scan_wp_ctr; do
  W=${spec%%:*}; K=${spec##*:}
  rm -f gpurun_out/${T}_full_$W.ncu-rep
  timeout 400 ncu --set full --import-source on --clock-control none -k regex:$K -c 1 -o gpurun_out/${T}_full_$W \
    python bench.py --workload $W --series 2960000 --steps 1 --warmup 1 --no-e2e --no-cpu --no-c5 > gpurun_out/${T}_full_$W.log 2>&1
  tail -1 gpurun_out/${T}_full_$W.log; ls -la gpurun_out/${T}_full_$W.ncu-rep
done
