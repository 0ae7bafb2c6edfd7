"""Per-phase clock64 timeline of the persistent tensor-core conv kernel, CTA 0 (debug aid): for every layer of one UNet
forward {setup done, all MMAs of the CTA's items issued, first item: accumulators ready / TMEM read / GroupNorm+Mish done,
CTA done} in us since the CTA's start, next to the items per CTA.  python tools/tc_timeline.py [workload] [t]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from mpd_public_b200 import _lib

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg5"
t_step = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = torch.device("cuda", 0)
model, guide, ds, prob, sd, n_grid, _mk = bench.build_problem(wl, dev)
mid, H, B, opt, wc, ws = bench.WORKLOADS[wl]
eng = model._engine(H)
lib = _lib.lib()
x = torch.randn((B, H, prob.robot.state_dim), device=dev)
eng.set_option("mega", 0)
for _ in range(3):
    eng.unet_forward_uniform(x, t_step)
_lib.check(lib.mpdb_engine_set_option(eng.handle, b"timeline", 1.0))
eng.unet_forward_uniform(x, t_step)
torch.cuda.synchronize()
n = lib.mpdb_engine_num_ops(eng.handle) - 1
buf = (C.c_int64 * (16 * n))()
_lib.check(lib.mpdb_engine_read_timeline(eng.handle, buf, n))
a = np.array(buf[:]).reshape(n, 16).astype(np.float64)
print(f"{wl} B={B} H={H} t={t_step}: precision {lib.mpdb_engine_step_precision(eng.handle, t_step)}")
print("op   setup  mma_all_issued  item0: acc_ready  tmem_read  gn_bar1  gn_bar2  gn_mish_done |  cta_done   (us since CTA 0 started)")
for i in range(n):
    if a[i, 0] == 0:
        continue
    d = (a[i] - a[i, 0]) / 1965.0
    print(f"{i:3d} {d[1]:6.2f} {d[3]:12.2f} {d[4]:16.2f} {d[5]:10.2f} {d[8]:8.2f} {d[9]:8.2f} {d[6]:12.2f}   | {d[7]:8.2f}")
