#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest.log
timeout 300 python tools/mega_timeline.py 2>&1 | grep -v "models/temporal" > gpurun_out/mega_timeline.txt; grep -E "mega=|total" gpurun_out/mega_timeline.txt; grep -E "^ *(2|17|25) " gpurun_out/mega_timeline.txt | tail -3
MPDB_GUIDE_TIMELINE=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_g.json 2> gpurun_out/bench_g.err; grep "guide timeline" gpurun_out/bench_g.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_g.json").read().strip().splitlines()[-1])
print(round(d["value"]), d["ms_per_step"], round(d["e2e"]["value"]), d["e2e"]["ms_per_step"], d["gpu_launches"], d["roofline"]["frac"], d["roofline_sdf"]["ms_per_launch"])
PY
