#!/bin/bash
# ncu captures only (launch list + full set for conv / guide / final kernels)
set -u
mkdir -p gpurun_out
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_loop.py > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches exit $?"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_kernel -s 0 -c 40 -f -o gpurun_out/prof_conv python tools/profile_loop.py > gpurun_out/ncu_conv.log 2>&1; echo "ncu conv exit $?"
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"guide_step|final_kernel" -c 4 -f -o gpurun_out/prof_guide python tools/profile_loop.py > gpurun_out/ncu_guide.log 2>&1; echo "ncu guide exit $?"
ls -la gpurun_out
