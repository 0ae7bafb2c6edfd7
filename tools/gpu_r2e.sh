#!/bin/bash
# Round 2, session E: ncu evidence of the round-2 build. Launch lists (cfg4 loop, cfg5 per-layer persistent kernel),
# --set full captures of the cluster kernel (one-product and three-product steps), the guide kernel and the persistent
# per-layer kernel at the cfg5 shape; summaries and source-level hot spots are extracted on the box (gpurun_out is capped
# at 64 MiB: the big reports are deleted after summarising). Numbers printed under ncu are never bench values.
set -u
mkdir -p gpurun_out
echo "== benched-config parity"; timeout 1200 python -m pytest tests/test_gpu_benched.py -m gpu -q --tb=short -p no:cacheprovider -s -x > gpurun_out/pytest_benched.log 2>&1; echo "exit $?"
grep -E "passed|failed|error|worst|amplification" gpurun_out/pytest_benched.log | tail -20
echo "== bench cfg4"; timeout 600 python bench.py --workload cfg4 --steps 9 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err; echo "exit $?"
python -c "
import json; d=json.load(open('gpurun_out/bench_cfg4.json')); r=d['roofline']
print('cfg4 value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'unet us', r['forward_us_by_precision'], 'precs', d['precision_policy']['products_per_mma_step_by_loop_step'])"
for wl in cfg4 cfg5; do
  timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$wl.csv python tools/profile_loop.py --workload $wl > gpurun_out/ncu_launches_$wl.log 2>&1; echo "ncu launches $wl exit $?"
  python tools/summarize_launches.py gpurun_out/launches_$wl.csv > gpurun_out/launches_${wl}_summary.txt 2>&1; cat gpurun_out/launches_${wl}_summary.txt
done
# cluster kernel: launches 0..5 of a loop are three-product steps (t = 24..19), later ones one-product
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"unet_mega" -s 1 -c 1 -f -o gpurun_out/prof_mega_p3 python tools/profile_loop.py > gpurun_out/ncu_mega_p3.log 2>&1; echo "ncu mega p3 exit $?"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"unet_mega" -s 12 -c 1 -f -o gpurun_out/prof_mega_p1 python tools/profile_loop.py > gpurun_out/ncu_mega_p1.log 2>&1; echo "ncu mega p1 exit $?"
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"guide_step" -s 2 -c 1 -f -o gpurun_out/prof_guide python tools/profile_loop.py > gpurun_out/ncu_guide.log 2>&1; echo "ncu guide exit $?"
# persistent per-layer kernel, cfg5 shape: one-product forward (skip 20 forwards x 39 conv launches), layers 1..12
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"conv5_tc" -s 780 -c 12 -f -o gpurun_out/prof_tc_cfg5 python tools/profile_loop.py --workload cfg5 > gpurun_out/ncu_tc_cfg5.log 2>&1; echo "ncu tc cfg5 exit $?"
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"guide_step" -s 2 -c 1 -f -o gpurun_out/prof_guide_cfg5 python tools/profile_loop.py --workload cfg5 > gpurun_out/ncu_guide_cfg5.log 2>&1; echo "ncu guide cfg5 exit $?"
for r in mega_p3 mega_p1 guide tc_cfg5 guide_cfg5; do
  python tools/ncu_summary.py gpurun_out/prof_$r.ncu-rep > gpurun_out/ncu_${r}_summary.csv 2> gpurun_out/ncu_${r}_traffic.txt; cat gpurun_out/ncu_${r}_traffic.txt
  python tools/ncu_hotspots.py gpurun_out/prof_$r.ncu-rep 60 > gpurun_out/ncu_${r}_hotspots.txt 2>&1
  [ $r = tc_cfg5 ] && python tools/ncu_hotspots.py gpurun_out/prof_$r.ncu-rep 60 7 > gpurun_out/ncu_${r}_hotspots_l8.txt 2>&1
done
rm -f gpurun_out/prof_tc_cfg5.ncu-rep gpurun_out/prof_guide_cfg5.ncu-rep gpurun_out/prof_mega_p3.ncu-rep
du -sh gpurun_out; ls -la gpurun_out | head -60
