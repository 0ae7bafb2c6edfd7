#!/bin/bash
# Round 2, session B: GPU test-suite, the other bench workloads, ncu --set full of the guide and cluster kernels.
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 2400 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -s > gpurun_out/pytest.log 2>&1; echo "pytest exit $?"
grep -E "passed|failed|error" gpurun_out/pytest.log | tail -3
grep -E "^\[|FAILED|Error|assert" gpurun_out/pytest.log | grep -v "mega vs oracle" | head -60
for wl in cfg4 cfg5 cfg3 cfg2 cfg4_ddim; do
  echo "== bench $wl"; timeout 600 python bench.py --workload $wl --steps 9 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; echo "exit $?"; tail -2 gpurun_out/bench_$wl.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_$wl.json"))
    r, s = d["roofline"], d["roofline_sdf"]
    print("$wl value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"],
          "| unet us", r["forward_us_by_precision"], "useful TF", round(r["achieved"], 1), "issued frac", round(r["issued_frac"], 4),
          "| guide ms/launch", round(s["ms_per_launch"], 4), "evals", s["evaluations_per_launch"], "frac", round(s["frac"], 4))
except Exception as e:
    print("$wl parse error", e)
PY
done
echo "== ncu full guide"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"guide_step" -s 1 -c 2 -f -o gpurun_out/prof_guide python tools/profile_loop.py > gpurun_out/ncu_guide.log 2>&1; echo "ncu guide exit $?"
echo "== ncu full mega"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"unet_mega" -s 8 -c 2 -f -o gpurun_out/prof_mega python tools/profile_loop.py > gpurun_out/ncu_mega.log 2>&1; echo "ncu mega exit $?"
ls -la gpurun_out | head -40
