#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest.log
for b in 100 400; do echo "== B=$b"; timeout 200 python tools/mega_timeline.py --batch $b 2>&1 | grep -E "mega="; done
echo "== cfg5"; timeout 300 python tools/mega_timeline.py --workload cfg5 2>&1 | grep -E "mega="
MPDB_MEGA=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_layers.json 2> gpurun_out/bench_layers.err; echo "layers exit $?"
timeout 300 python bench.py --workload cfg5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg5.json 2> gpurun_out/bench_cfg5.err; echo "cfg5 exit $?"
python - <<'PY'
import json
for f in ("gpurun_out/bench_layers.json","gpurun_out/bench_cfg5.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), d["ms_per_step"], round(d["e2e"]["value"]), d["gpu_launches"], d["roofline"]["frac"], d["roofline"]["achieved"])
    except Exception as e: print(f, "ERR", e)
PY
