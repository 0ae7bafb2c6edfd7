#!/bin/bash
# ncu captures of one guided loop: launch list (gpu__time_duration) + full set for the whole-forward cluster kernel, the
# guide and the final kernels. Numbers printed under ncu are never bench values.
set -u
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_loop.py --graph > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches exit $?"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"unet_mega" -s 2 -c 3 -f -o gpurun_out/prof_mega python tools/profile_loop.py > gpurun_out/ncu_mega.log 2>&1; echo "ncu mega exit $?"
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"guide_step|final_kernel" -c 6 -f -o gpurun_out/prof_guide python tools/profile_loop.py > gpurun_out/ncu_guide.log 2>&1; echo "ncu guide exit $?"
python tools/summarize_launches.py gpurun_out/launches.csv > gpurun_out/launches_summary.txt 2>&1; cat gpurun_out/launches_summary.txt
python tools/ncu_summary.py gpurun_out/prof_mega.ncu-rep > gpurun_out/ncu_mega_summary.csv 2> gpurun_out/ncu_mega_traffic.txt; cat gpurun_out/ncu_mega_traffic.txt
python tools/ncu_summary.py gpurun_out/prof_guide.ncu-rep > gpurun_out/ncu_guide_summary.csv 2> gpurun_out/ncu_guide_traffic.txt; cat gpurun_out/ncu_guide_traffic.txt
ls -la gpurun_out | head -40
