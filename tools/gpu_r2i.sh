#!/bin/bash
# Round 2, session I (final build): full GPU suite, every bench workload (20 steps / 5 warm-up as the driver runs it, CPU baseline
# on cfg4), guide timeline, in-kernel timelines, launch lists and ncu --set full captures of the final guide / per-layer kernels.
set -u
mkdir -p gpurun_out
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest exit $?"
grep -E "passed|failed|error" gpurun_out/pytest.log | tail -3
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
echo "== bench cfg4 (with CPU baseline)"; MPDB_GUIDE_TIMELINE=1 timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err; echo "exit $?"; grep "guide timeline" gpurun_out/bench_cfg4.err | head -6
for wl in cfg5 cfg3 cfg2 cfg4_ddim; do
  echo "== bench $wl"; timeout 600 python bench.py --workload $wl --steps 18 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; echo "exit $?"; tail -2 gpurun_out/bench_$wl.err
done
python - <<'PY'
import json
for wl in ("cfg4", "cfg5", "cfg3", "cfg2", "cfg4_ddim"):
    try:
        d = json.load(open(f"gpurun_out/bench_{wl}.json"))
        r, s = d["roofline"], d["roofline_sdf"]
        print(wl, "value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"],
              "| unet us", r["forward_us_by_precision"], "useful TF", round(r["achieved"], 1), "frac", round(r["frac"], 4), "issued", round(r["issued_frac"], 4),
              "| guide ms/launch", round(s["ms_per_launch"], 4), "evals", s["evaluations_per_launch"], "frac", round(s["frac"], 4),
              "| cpu", (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e:
        print(wl, "parse error", e)
PY
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_arm.json 2> gpurun_out/bench_reference_arm.err; echo "exit $?"; cut -c1-200 gpurun_out/bench_reference_arm.json
timeout 300 python tools/mega_timeline.py --t 5 > gpurun_out/mega_timeline_t5.txt 2>&1; grep -E "phase sums|total|mega=" gpurun_out/mega_timeline_t5.txt
timeout 300 python tools/tc_timeline.py cfg5 5 > gpurun_out/tc_timeline_cfg5_t5.txt 2>&1; head -6 gpurun_out/tc_timeline_cfg5_t5.txt
for wl in cfg4 cfg5; do
  timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$wl.csv python tools/profile_loop.py --workload $wl > gpurun_out/ncu_launches_$wl.log 2>&1; echo "ncu launches $wl exit $?"
  python tools/summarize_launches.py gpurun_out/launches_$wl.csv > gpurun_out/launches_${wl}_summary.txt 2>&1; cat gpurun_out/launches_${wl}_summary.txt
done
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"guide_step" -s 2 -c 1 -f -o gpurun_out/prof_guide python tools/profile_loop.py > gpurun_out/ncu_guide.log 2>&1; echo "ncu guide exit $?"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"conv5_tc" -s 780 -c 12 -f -o gpurun_out/prof_tc_cfg5 python tools/profile_loop.py --workload cfg5 > gpurun_out/ncu_tc_cfg5.log 2>&1; echo "ncu tc cfg5 exit $?"
timeout 600 ncu --profile-from-start off --set full --clock-control none -k regex:"final_kernel" -s 2 -c 2 -f -o gpurun_out/prof_final_cfg5 python tools/profile_loop.py --workload cfg5 > gpurun_out/ncu_final_cfg5.log 2>&1; echo "ncu final cfg5 exit $?"
for r in guide tc_cfg5 final_cfg5; do
  python tools/ncu_summary.py gpurun_out/prof_$r.ncu-rep > gpurun_out/ncu_${r}_summary.csv 2> gpurun_out/ncu_${r}_traffic.txt; cat gpurun_out/ncu_${r}_traffic.txt
  python tools/ncu_hotspots.py gpurun_out/prof_$r.ncu-rep 40 > gpurun_out/ncu_${r}_hotspots.txt 2>&1
done
python tools/ncu_hotspots.py gpurun_out/prof_tc_cfg5.ncu-rep 40 7 > gpurun_out/ncu_tc_cfg5_hotspots_l8.txt 2>&1
rm -f gpurun_out/prof_tc_cfg5.ncu-rep gpurun_out/prof_final_cfg5.ncu-rep
cuobjdump -sass mpd_public_b200/libmpdb200.so | grep -oE "^\s+/\*[0-9a-f]+\*/\s+[A-Z0-9_.]+" | awk '{print $2}' | grep -E "^(UTCHMMA|UTCBAR|LDTM|UBLKCP|UTMALDG|UTMAPF|STAS|SYNCS|ELECT|UCGABAR|MUFU)" | sed 's/\..*//' | sort | uniq -c | sort -rn > gpurun_out/sass_mnemonics.txt; cat gpurun_out/sass_mnemonics.txt
du -sh gpurun_out
