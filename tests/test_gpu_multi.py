"""Batch-sharded sampling over NCCL on all GPUs of one box (2 to 8) equals the single-GPU run on the same injected noise
(SURVEY §8e): bit for bit, unguided and guided — the guide counts the (trajectory, evaluation) pairs whose normaliser clamp
was decided by other trajectories of the batch (the only cross-trajectory coupling of the path); with none of
those in the single-GPU run every sharding must reproduce it exactly. Skipped when fewer than 2 GPUs are visible.
`python tests/test_gpu_multi.py` runs it stand-alone and prints a one-line JSON summary (tools/gpu_multi.sh commits it to profiles/)."""
import os
import socket
import subprocess
import sys
import tempfile

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
rank, world, port, out = int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], sys.argv[5]
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world, device_id=dev)
import bench
from mpd_public_b200.parallel import sample_sharded
workload = sys.argv[6]
model, guide, ds, prob, sd, n_grid, _mk = bench.build_problem(workload, dev)
B, H, D = 8 * 13, bench.WORKLOADS[workload][1], prob.robot.state_dim   # 104 trajectories: 13 per rank at 8 ranks, 52 at 2
hard = ds.get_hard_conditions(torch.vstack((torch.as_tensor(prob.start), torch.as_tensor(prob.goal))).to(dev), normalize=True)
gen = torch.Generator().manual_seed(42)
noise = torch.randn((31, B, H, D), generator=gen).to(dev)          # the same GLOBAL noise on every rank
res = {}
for tag, g in (("unguided", None), ("guided", guide)):
    kw = dict(bench.sample_kwargs(g))
    res[tag] = sample_sharded(lambda n, nz: model.sample(hard, n, noise=nz, **kw), B, noise=noise)
    if rank == 0:
        if g is not None:
            g.batch_dependent_clamps(reset=True)
        res[tag + "_single"] = model.sample(hard, B, noise=noise, **kw)
        if g is not None:
            res["dependent"] = torch.tensor(g.batch_dependent_clamps(reset=True))
if rank == 0:
    torch.save({k: v.cpu() for k, v in res.items()}, out)
dist.barrier()
dist.destroy_process_group()
'''


def run_equivalence(world, workload="cfg2"):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    with tempfile.TemporaryDirectory() as d:
        path, out = os.path.join(d, "w.py"), os.path.join(d, "out.pt")
        open(path, "w").write(_WORKER)
        procs = [subprocess.Popen([sys.executable, path, ROOT, str(r), str(world), str(port), out, workload], stdout=subprocess.PIPE,
                                  stderr=subprocess.STDOUT, text=True) for r in range(world)]
        outs = [p.communicate(timeout=900)[0] for p in procs]
        for p, o in zip(procs, outs):
            assert p.returncode == 0, o[-3000:]
        res = torch.load(out)
    summary = {"world_size": world, "workload": workload, "batch": int(res["unguided"].shape[0]),
               "unguided_bit_identical": bool(torch.equal(res["unguided"], res["unguided_single"])),
               "guided_bit_identical": bool(torch.equal(res["guided"], res["guided_single"])),
               "guided_max_rel_diff": float((res["guided"] - res["guided_single"]).abs().max() / res["guided_single"].abs().max()),
               "batch_dependent_clamps_in_single_gpu_run": int(res["dependent"])}
    assert summary["unguided_bit_identical"]  # trajectories are independent
    if summary["batch_dependent_clamps_in_single_gpu_run"] == 0:
        assert summary["guided_bit_identical"], summary
    else:  # shards couple through LimitsNormalizer's batch-global clamp where a trajectory sits in its 1e-4 band
        assert summary["guided_max_rel_diff"] < 1e-3, summary
    return summary


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("workload", ["cfg2", "cfg4"])
def test_nccl_sharded_sampling_matches_single_gpu(workload):
    print(run_equivalence(min(torch.cuda.device_count(), 8), workload))


if __name__ == "__main__":
    import json
    n = min(torch.cuda.device_count(), 8)
    for wl in ("cfg2", "cfg4"):
        print(json.dumps(run_equivalence(n, wl)))
