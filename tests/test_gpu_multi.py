"""Batch-sharded sampling over NCCL on 2 GPUs of one box equals the single-GPU run on the same injected noise
(SURVEY §8e). Skipped when fewer than 2 GPUs are visible."""
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
model, guide, ds, prob, sd, n_grid = bench.build_problem("cfg2", dev)
B, H, D = 64, 64, prob.robot.state_dim
hard = ds.get_hard_conditions(torch.vstack((torch.as_tensor(prob.start), torch.as_tensor(prob.goal))).to(dev), normalize=True)
gen = torch.Generator().manual_seed(42)
noise = torch.randn((31, B, H, D), generator=gen).to(dev)          # the same GLOBAL noise on every rank
res = {}
for tag, g in (("unguided", None), ("guided", guide)):
    kw = dict(bench.sample_kwargs(g))
    res[tag] = sample_sharded(lambda n, nz: model.sample(hard, n, noise=nz, **kw), B, noise=noise)
    if rank == 0:
        res[tag + "_single"] = model.sample(hard, B, noise=noise, **kw)
if rank == 0:
    torch.save({k: v.cpu() for k, v in res.items()}, out)
dist.barrier()
dist.destroy_process_group()
'''


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_nccl_sharded_sampling_matches_single_gpu():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    with tempfile.TemporaryDirectory() as d:
        path, out = os.path.join(d, "w.py"), os.path.join(d, "out.pt")
        open(path, "w").write(_WORKER)
        procs = [subprocess.Popen([sys.executable, path, ROOT, str(r), "2", str(port), out], stdout=subprocess.PIPE,
                                  stderr=subprocess.STDOUT, text=True) for r in range(2)]
        outs = [p.communicate(timeout=600)[0] for p in procs]
        for p, o in zip(procs, outs):
            assert p.returncode == 0, o[-3000:]
        res = torch.load(out)
    assert res["unguided"].shape == (64, 64, 4)
    assert torch.equal(res["unguided"], res["unguided_single"])  # trajectories are independent: bit-identical shards
    # guided: shards couple only through LimitsNormalizer's batch-global clip branch
    err = float((res["guided"] - res["guided_single"]).abs().max() / res["guided_single"].abs().max())
    assert err < 1e-3, err
