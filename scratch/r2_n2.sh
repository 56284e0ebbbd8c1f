#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 2>gpurun_out/n2.err | tail -1 > gpurun_out/r2_n2.json
tail -3 gpurun_out/n2.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2_n2.json"))
print("n_gpus", d["n_gpus"], "ms", d["ms_per_step"], "per rank", d.get("ms_per_rank"), "value", d["value"], "frac", d["roofline"]["frac"])
print("c5", {k: d.get("c5", {}).get(k) for k in ("ms_per_step", "ms_per_rank", "value", "collective", "error")})
print("e2e", d.get("e2e"), "resident", d.get("e2e_resident"))
PY
