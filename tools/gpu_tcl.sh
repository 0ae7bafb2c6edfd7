#!/bin/bash
set -u
mkdir -p gpurun_out
for t in 5 20; do timeout 300 python tools/tc_timeline.py cfg5 $t > gpurun_out/tc_timeline_cfg5_t$t.txt 2>&1; cat gpurun_out/tc_timeline_cfg5_t$t.txt | tail -45; done
