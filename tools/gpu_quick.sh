#!/bin/bash
# quick check: cluster-kernel tests + cfg4 bench + timeline
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_mega.py tests/test_gpu_benched.py -m gpu -q --tb=short -p no:cacheprovider -x > gpurun_out/pytest_quick.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_quick.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; python -c "
import json; d=json.load(open('gpurun_out/bench_quick.json')); print('cfg4 value', round(d['value']), 'e2e', round(d['e2e']['value']), d['roofline']['forward_us_by_precision'], 'guide', d['roofline_sdf']['ms_per_launch'])"
for t in 5 20; do timeout 300 python tools/mega_timeline.py --t $t > gpurun_out/mega_timeline_t$t.txt 2>&1; grep -E "phase sums|total|mega=1" gpurun_out/mega_timeline_t$t.txt; done
