#!/bin/bash
set -u
mkdir -p gpurun_out
for b in 16 64 128 136 200 256 400; do
  echo "== B=$b"; timeout 200 python tools/mega_timeline.py --batch $b 2>&1 | grep -E "mega=|mega_info"
done
echo "== cfg5 (H=128, B=512)"; timeout 300 python tools/mega_timeline.py --workload cfg5 2>&1 | grep -E "mega=|mega_info"
