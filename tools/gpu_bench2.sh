#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_mega.json 2> gpurun_out/bench_mega.err; echo "mega exit $?"; tail -3 gpurun_out/bench_mega.err
MPDB_MEGA=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_layers.json 2> gpurun_out/bench_layers.err; echo "layers exit $?"
python - <<'PY'
import json
for f in ("gpurun_out/bench_mega.json","gpurun_out/bench_layers.json"):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, round(d["value"]), d["ms_per_step"], round(d["e2e"]["value"]), d["gpu_launches"], d["roofline"]["frac"], d.get("roofline_sdf",{}).get("ms_per_launch"))
PY
