#!/bin/bash
set -u
mkdir -p gpurun_out
for w in cfg2 cfg5; do
  timeout 400 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "$w exit $?"; tail -2 gpurun_out/bench_$w.err
done
python - <<'PY'
import json
for w in ("cfg2", "cfg5"):
    try:
        d = json.load(open(f"gpurun_out/bench_{w}.json"))
        r = d["roofline"]
        print(w, d["value"], d["ms_per_step"], "fwd_ms", r["unet_forward_ms"], "TF", r["achieved"], "frac", r["frac"], "guide", d["roofline_sdf"]["ms_per_launch"], d["roofline_sdf"]["frac"])
        print(r["per_launch_us"])
    except Exception as e:
        print(w, "ERR", e)
PY
