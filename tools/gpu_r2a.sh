#!/bin/bash
# Round 2, session A: the whole GPU test-suite, smoke, the cfg4 bench (with the guide's phase stamps), the cluster kernel's
# timeline at both precisions. Outputs -> gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.csv 2>&1
echo "== pytest -m gpu"; timeout 2400 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -s > gpurun_out/pytest.log 2>&1; echo "pytest exit $?"
grep -E "passed|failed|error" gpurun_out/pytest.log | tail -5
grep -E "^\[|^t=|FAILED|Error|assert" gpurun_out/pytest.log | head -80
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke.log
echo "== bench"; MPDB_GUIDE_TIMELINE=1 timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -12 gpurun_out/bench.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench.json"))
    print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"])
    print("roofline", d["roofline"]["achieved"], d["roofline"]["frac"], d["roofline"]["forward_us_by_precision"], d["roofline"]["forwards_by_precision"])
    print("sdf", d["roofline_sdf"]["ms_per_launch"], d["roofline_sdf"]["evaluations_per_launch"], d["roofline_sdf"]["ms_single_evaluation_launch"], d["roofline_sdf"]["frac"])
    print("cpu", d.get("cpu_baseline"))
except Exception as e:
    print("bench parse error", e)
PY
echo "== bench force (22-bit split everywhere)"; timeout 300 python bench.py --steps 10 --warmup 3 --tc force --no-cpu-baseline > gpurun_out/bench_force.json 2> gpurun_out/bench_force.err; python -c "
import json; d=json.load(open('gpurun_out/bench_force.json')); print('force value', d['value'], d['ms_per_step'])"
echo "== mega timeline t=5"; timeout 300 python tools/mega_timeline.py --t 5 > gpurun_out/mega_timeline_t5.txt 2>&1; head -3 gpurun_out/mega_timeline_t5.txt; grep -E "phase sums|total|mega=" gpurun_out/mega_timeline_t5.txt
echo "== mega timeline t=20"; timeout 300 python tools/mega_timeline.py --t 20 > gpurun_out/mega_timeline_t20.txt 2>&1; grep -E "phase sums|total|mega=" gpurun_out/mega_timeline_t20.txt
ls -la gpurun_out | head -30
