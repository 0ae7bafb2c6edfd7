#!/bin/bash
# 8-GPU session (one box): NCCL shard-equivalence at 8 ranks (unguided + guided vs the single-GPU run on the same injected
# noise) and the 8-GPU bench lines of BASELINE configs 4 (100 trajectories per GPU) and 5 (H=128, 512 per GPU = 4096).
set -u
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt; wc -l gpurun_out/gpus.txt
timeout 900 python tests/test_gpu_multi.py > gpurun_out/shard_equivalence_n8.jsonl 2> gpurun_out/shard_equivalence_n8.err; echo "equivalence exit $?"; cat gpurun_out/shard_equivalence_n8.jsonl; tail -3 gpurun_out/shard_equivalence_n8.err
for wl in cfg4 cfg5; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 --workload $wl > gpurun_out/bench_${wl}_n8.json 2> gpurun_out/bench_${wl}_n8.err; echo "bench $wl n8 exit $?"; tail -2 gpurun_out/bench_${wl}_n8.err
done
python - <<'PY'
import json
for f in ("bench_cfg4_n8", "bench_cfg5_n8"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, d.get("n_gpus"), round(d["value"]), round(d["ms_per_step"], 3), round(d.get("e2e", {}).get("value", 0)), d["config"]["global_batch"])
    except Exception as e:
        print(f, "ERR", e)
PY
