"""Per-phase clock64 timeline of the tensor-core conv kernels of one UNet forward (debug aid)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from mpd_public_b200 import _lib

dev = torch.device("cuda", 0)
model, guide, ds, prob, sd, n_grid, _mk = bench.build_problem("cfg4", dev)
B = 100
model.tensor_cores = "force"
eng = model._engine()
lib = _lib.lib()
x = torch.randn((B, 64, 14), device=dev)
t = torch.full((B,), 5, dtype=torch.long, device=dev)
for _ in range(3):
    model.model(x, t, None)
_lib.check(lib.mpdb_engine_set_option(eng.handle, b"timeline", 1.0))
model.model(x, t, None)
n = lib.mpdb_engine_num_ops(eng.handle) - 1
buf = (C.c_int64 * (16 * n))()
_lib.check(lib.mpdb_engine_read_timeline(eng.handle, buf, n))
a = np.array(buf[:]).reshape(n, 16)
names = ["setup", "loads_issued", "mma_issued", "acc_ready", "tmem_ld", "gn_mish", "end", "gn_bar1", "gn_bar2"]
print("op   " + " ".join(f"{k:>12s}" for k in names) + "   (cycles since kernel start of CTA 0,0; 1965 cycles = 1 us)")
for i in range(n):
    if a[i, 0] == 0:
        continue
    d = a[i, 1:10] - a[i, 0]
    print(f"{i:3d}  " + " ".join(f"{int(v):12d}" for v in d))
