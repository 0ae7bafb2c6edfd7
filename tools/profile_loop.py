"""One guided loop between cudaProfilerStart/Stop, for `ncu --profile-from-start off` (see profiles/README.md)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cfg4")
ap.add_argument("--loops", type=int, default=1)
ap.add_argument("--graph", action="store_true")
args = ap.parse_args()
dev = torch.device("cuda", 0)
model, guide, ds, prob, sd, n_grid, _mk = bench.build_problem(args.workload, dev)
mid, H, B, opt, wc, ws = bench.WORKLOADS[args.workload]
model.use_cuda_graph = args.graph
hard = ds.get_hard_conditions(torch.vstack((torch.as_tensor(prob.start), torch.as_tensor(prob.goal))).to(dev), normalize=True)
noise = torch.randn((31, B, H, prob.robot.state_dim), device=dev)
kw = bench.sample_kwargs(guide)
for _ in range(2):
    model.sample(hard, B, noise=noise, **kw)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(args.loops):
    model.sample(hard, B, noise=noise, **kw)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled", args.loops, "loop(s)")
