"""Host-side breakdown of one end-to-end run_inference() call (debug aid for bench.py's e2e number)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench

dev = torch.device("cuda", 0)
model, guide, ds, prob, sd, n_grid, _mk = bench.build_problem("cfg4", dev)
mid, H, B, opt, wc, ws = bench.WORKLOADS["cfg4"]
kw = bench.sample_kwargs(guide)
sg_host = torch.vstack((torch.as_tensor(prob.start), torch.as_tensor(prob.goal))).pin_memory()
out_host = torch.empty((B, H, prob.robot.state_dim), dtype=torch.float32).pin_memory()


def sync():
    torch.cuda.synchronize()
    return time.perf_counter()


for it in range(6):
    t0 = sync()
    sg = sg_host.to(dev, non_blocking=True)
    hc = ds.get_hard_conditions(sg, normalize=True)
    t1h = time.perf_counter(); t1 = sync()
    x = model.run_inference(None, hc, n_samples=B, horizon=H, return_chain=False, **kw)
    t2h = time.perf_counter(); t2 = sync()
    out_host.copy_(x, non_blocking=True)
    t3 = sync()
    if it >= 3:
        print(f"hard conds: host {1e3*(t1h-t0):.3f} ms, done {1e3*(t1-t0):.3f} | run_inference: host {1e3*(t2h-t1):.3f} ms, done {1e3*(t2-t1):.3f} | d2h {1e3*(t3-t2):.3f}")

# inside run_inference: cProfile of the host side
import cProfile
import pstats
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    x = model.run_inference(None, hc, n_samples=B, horizon=H, return_chain=False, **kw)
    torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
