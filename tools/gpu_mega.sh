#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== mega tests"; timeout 600 python -m pytest tests/test_gpu_mega.py -q -s --tb=short -p no:cacheprovider > gpurun_out/pytest_mega.log 2>&1; echo "exit $?"
grep -E "passed|failed|Error|error" gpurun_out/pytest_mega.log | head -20
echo "== timeline"; timeout 300 python tools/mega_timeline.py 2>&1 | grep -v "models/temporal" > gpurun_out/mega_timeline.txt; echo "exit $?"; cat gpurun_out/mega_timeline.txt
if [ "${1:-}" == "probe" ]; then echo "== bulk probe"; timeout 120 tools/probes/bulk_probe > gpurun_out/bulk_probe.txt 2>&1; cat gpurun_out/bulk_probe.txt; fi
