#!/bin/bash
# Round 2, session G: cluster-kernel changes: parity of the cluster kernel (vs oracle, vs per-layer path), benched-config parity,
# cfg4 / cfg2 bench lines, per-layer timeline of a one-product forward.
set -u
mkdir -p gpurun_out
echo "== pytest"; timeout 1500 python -m pytest tests/test_gpu_mega.py tests/test_gpu_benched.py tests/test_gpu_parity.py -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest.log 2>&1; echo "pytest exit $?"
grep -E "passed|failed|error" gpurun_out/pytest.log | tail -3
grep -E "FAILED|Error|assert " gpurun_out/pytest.log | head -20
for wl in cfg4 cfg2; do
  echo "== bench $wl"; timeout 600 python bench.py --workload $wl --steps 9 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; echo "exit $?"; tail -3 gpurun_out/bench_$wl.err
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
echo "== mega timeline t=5"; timeout 300 python tools/mega_timeline.py --t 5 > gpurun_out/mega_timeline_t5.txt 2>&1; grep -E "phase sums|total|mega=" gpurun_out/mega_timeline_t5.txt; sed -n 6,46p gpurun_out/mega_timeline_t5.txt | cut -c1-75
