#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x --tb=short -p no:cacheprovider -k "guide or loop or trajectory or steps" > gpurun_out/pytest_guide.log 2>&1; echo "exit $?"; tail -5 gpurun_out/pytest_guide.log
MPDB_GUIDE_TIMELINE=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_g.json 2> gpurun_out/bench_g.err; grep "guide timeline" gpurun_out/bench_g.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_g.json").read().strip().splitlines()[-1])
print(round(d["value"]), d["ms_per_step"], round(d["e2e"]["value"]), d["e2e"]["ms_per_step"], d["gpu_launches"], d["roofline"]["frac"], d["roofline_sdf"]["ms_per_launch"])
PY
