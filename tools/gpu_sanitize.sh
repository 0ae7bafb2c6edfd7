#!/bin/bash
# compute-sanitizer over every hand-written kernel family (memcheck / racecheck / synccheck). Summaries -> gpurun_out/sanitize_*.txt
set -u
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck synccheck racecheck; do
  for what in ${SANITIZE_WHAT:-mega layers layers_big guide loop}; do
    echo "== $tool $what"
    timeout 900 $CS --tool $tool --print-limit 20 --error-exitcode 3 python tools/sanitize_driver.py $what > gpurun_out/sanitize_${tool}_${what}.txt 2>&1
    echo "exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Error|error:" gpurun_out/sanitize_${tool}_${what}.txt | head -6
  done
done
