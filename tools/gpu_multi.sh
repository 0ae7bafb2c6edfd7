#!/bin/bash
# 2-GPU session: NCCL shard-equivalence test and the bench at N=1 and N=2
set -u
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
timeout 600 python -m pytest tests/test_gpu_multi.py -q --tb=short -p no:cacheprovider 2>&1 | tail -15
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -2 gpurun_out/bench_n1.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -3 gpurun_out/bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err; tail -3 gpurun_out/bench_ref_n2.err
python - <<'PY'
import json
for f in ("bench_n1", "bench_n2", "bench_ref_n2"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, d.get("n_gpus"), d["value"], d["ms_per_step"], d.get("e2e", {}).get("value"))
    except Exception as e:
        print(f, "ERR", e)
PY
