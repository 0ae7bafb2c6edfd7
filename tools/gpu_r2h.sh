#!/bin/bash
# Round 2, session H: full GPU suite + cfg4 bench + end-to-end breakdown after the one-launch hard-condition normaliser
set -u
mkdir -p gpurun_out
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest.log 2>&1; echo "pytest exit $?"
grep -E "passed|failed|error" gpurun_out/pytest.log | tail -3
grep -E "FAILED|Error|assert " gpurun_out/pytest.log | head -20
for wl in cfg4; do
  echo "== bench $wl"; timeout 600 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; echo "exit $?"; tail -3 gpurun_out/bench_$wl.err
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_$wl.json"))
print("$wl value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), "e2e ms", round(d["e2e"]["ms_per_step"], 3), "launches", d["gpu_launches"])
PY
done
timeout 300 python tools/e2e_breakdown.py > gpurun_out/e2e_breakdown.txt 2>&1; head -4 gpurun_out/e2e_breakdown.txt
