#!/bin/bash
# One GPU-box session: parity tests, smoke, bench, (optionally) launch list and full ncu captures. Outputs -> gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.csv 2>&1
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest exit $?"
tail -30 gpurun_out/pytest.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke.log
echo "== bench"; timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -3 gpurun_out/bench.err
echo "== bench tc off"; timeout 300 python bench.py --steps 5 --warmup 3 --tc off --no-cpu-baseline > gpurun_out/bench_tcoff.json 2>> gpurun_out/bench.err
if [ "${1:-}" != "quick" ]; then
echo "== ncu launch list"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_loop.py --graph > gpurun_out/ncu_launches.log 2>&1; echo "ncu exit $?"
echo "== ncu full conv"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"conv5_tc|conv_kernel" -s 0 -c 40 -f -o gpurun_out/prof_conv python tools/profile_loop.py > gpurun_out/ncu_conv.log 2>&1; echo "ncu exit $?"
echo "== ncu full guide"
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"guide_step|final_kernel" -c 4 -f -o gpurun_out/prof_guide python tools/profile_loop.py > gpurun_out/ncu_guide.log 2>&1; echo "ncu exit $?"
fi
ls -la gpurun_out | head -30
