#!/bin/bash
# Round 2, session C: per-layer (persistent) path — tests, cfg5 / cfg3 bench with the two RTB-fusion rules
set -u
mkdir -p gpurun_out
echo "== pytest (per-layer related + all)"; timeout 2400 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -s -x > gpurun_out/pytest.log 2>&1; echo "pytest exit $?"
grep -E "passed|failed|error" gpurun_out/pytest.log | tail -3
grep -E "^\[panda_opt1_h128|FAILED|Error|assert" gpurun_out/pytest.log | head -30
for wl in cfg5 cfg3; do
 for fr in 1 2 0; do
  echo "== bench $wl fuse_rtb=$fr"; MPDB_FUSE_RTB=$fr timeout 600 python bench.py --workload $wl --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${wl}_rtb$fr.json 2> gpurun_out/bench_${wl}_rtb$fr.err; echo "exit $?"; tail -2 gpurun_out/bench_${wl}_rtb$fr.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_${wl}_rtb$fr.json"))
    r, s = d["roofline"], d["roofline_sdf"]
    print("$wl rtb=$fr value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]),
          "| unet us", r["forward_us_by_precision"], "useful TF", round(r["achieved"], 1), "| guide ms/launch", round(s["ms_per_launch"], 4))
    print("   per-launch us", r.get("per_launch_us"))
except Exception as e:
    print("$wl parse error", e)
PY
 done
done
echo "== bench cfg4 mega=0 (per-layer path at B=100)"; MPDB_MEGA=0 timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg4_layers.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_cfg4_layers.json')); print('cfg4 per-layer value', round(d['value']), d['roofline']['forward_us_by_precision'])"
bash tools/gpu_sanitize.sh
