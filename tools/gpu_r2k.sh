#!/bin/bash
# Round 2, session K: ncu --set full of the final cluster kernel (one three-product and one one-product launch) + hot spots
set -u
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"unet_mega" -s 1 -c 1 -f -o gpurun_out/prof_mega_p3 python tools/profile_loop.py > gpurun_out/ncu_mega_p3.log 2>&1; echo "ncu mega p3 exit $?"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"unet_mega" -s 12 -c 1 -f -o gpurun_out/prof_mega_p1 python tools/profile_loop.py > gpurun_out/ncu_mega_p1.log 2>&1; echo "ncu mega p1 exit $?"
for r in mega_p3 mega_p1; do
  python tools/ncu_summary.py gpurun_out/prof_$r.ncu-rep > gpurun_out/ncu_${r}_summary.csv 2> gpurun_out/ncu_${r}_traffic.txt; cat gpurun_out/ncu_${r}_traffic.txt
  python tools/ncu_hotspots.py gpurun_out/prof_$r.ncu-rep 40 > gpurun_out/ncu_${r}_hotspots.txt 2>&1
done
rm -f gpurun_out/prof_mega_p3.ncu-rep
head -3 gpurun_out/ncu_mega_p1_hotspots.txt
