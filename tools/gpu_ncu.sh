#!/bin/bash
# ncu captures (launch list over one guided loop + full set for the conv / guide / final kernels) and SASS evidence
set -u
mkdir -p gpurun_out
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_loop.py > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches exit $?"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"conv5_tc" -s 0 -c 40 -f -o gpurun_out/prof_conv python tools/profile_loop.py > gpurun_out/ncu_conv.log 2>&1; echo "ncu conv exit $?"
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"guide_step|final_kernel|blc_to_tc" -c 6 -f -o gpurun_out/prof_guide python tools/profile_loop.py > gpurun_out/ncu_guide.log 2>&1; echo "ncu guide exit $?"
ls -la gpurun_out
