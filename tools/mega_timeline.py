"""Per-layer clock64 timeline of the whole-forward cluster kernel (cluster 0, thread 0 of each CTA) + device time of the
UNet body with the cluster kernel and with per-layer kernels. Debug / profiling aid (profiles/README.md)."""
import argparse
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from mpd_public_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cfg4")
ap.add_argument("--batch", type=int, default=0)
ap.add_argument("--t", type=int, default=5, help="timestep of the forward: selects the precision (t = 5: one product, t = 20: 22-bit split)")
args = ap.parse_args()
dev = torch.device("cuda", 0)
model, guide, ds, prob, sd, n_grid, _mk = bench.build_problem(args.workload, dev)
mid, H, B, opt, wc, ws = bench.WORKLOADS[args.workload]
B = args.batch or B
D = prob.robot.state_dim
eng = model._engine()
lib = _lib.lib()
x = torch.randn((B, H, D), device=dev)
eng.unet_forward_uniform(x, args.t)  # loads the parameters into the engine
print("mega_info (in_use, G, layers, a_bytes, smem, why):", eng.mega_info(B))
print(f"t = {args.t}: MMA products per step = {lib.mpdb_engine_step_precision(eng.handle, args.t)}")


def body_ms(reps=50):
    ms, fl, n = C.c_float(), C.c_double(), C.c_int32()
    _lib.check(lib.mpdb_profile_unet_body(eng.handle, _lib.fptr(x), args.t, B, reps, C.byref(ms), C.byref(fl), C.byref(n),
                                          _lib.stream_ptr(dev)))
    return ms.value, fl.value, n.value


for mega in (1, 0):
    eng.set_option("mega", mega)
    ms, fl, n = body_ms()
    print(f"mega={mega}: UNet body {ms * 1e3:.1f} us per forward, {n} launches, {fl / ms / 1e9:.1f} TFLOP/s useful (stream-ordered launches)")
eng.set_option("mega", 1)
eng.set_option("mega_timeline", 1)
for _ in range(3):
    eng.unet_forward_uniform(x, args.t)
torch.cuda.synchronize()
nl = 48
ND = 16
buf = (C.c_int64 * (ND * 8 * nl))()
desc = (C.c_int32 * (4 * nl))()
n = lib.mpdb_engine_read_mega_timeline(eng.handle, buf, desc, nl)
a = np.array(buf[:ND * 8 * n]).reshape(n, 8, ND).astype(np.float64)
d = np.array(desc[:4 * n]).reshape(n, 4)
t0 = a[0, :, 0].min()
us = 1.0 / 1965.0
names = {0: "input", 1: "conv5", 2: "down", 3: "up"}
print("layer type    L   CO act | start(us)  wait->acc  acc->epi  epi->deliv  (rank 0)   | slowest rank: deliv-start")
for l in range(n):
    act = d[l, 3]
    r0 = a[l, 0]
    span = (a[l, :, 3] - a[l, :, 0]).max() * us
    print(f"{l:3d} {names[d[l,0]]:6s} {d[l,1]:4d} {d[l,2]:4d} {act:3d} | {(r0[0]-t0)*us:8.2f} {(r0[1]-r0[0])*us if r0[1] else 0:9.2f} "
          f"{(r0[2]-max(r0[1],r0[0]))*us:9.2f} {(r0[3]-r0[2])*us:10.2f}            | {span:8.2f}")
print("\nfine stamps, rank 0 (us after 'inputs landed'): mma_wake  w_ready  mma_issued | acc_done  tmem_ld  a_free_sent  gn_bar1  gn_bar2  epi_done  a_free_ok  stores  fenced")
for l in range(n):
    r0 = a[l, 0]
    rel = lambda k: (r0[k] - r0[0]) * us if r0[k] else float('nan')
    print(f"{l:3d} {names[d[l,0]]:6s} {d[l,1]:4d} {d[l,2]:4d} | {rel(8):7.2f} {rel(9):7.2f} {rel(10):7.2f} | {rel(1):7.2f} {rel(4):7.2f} {rel(5):7.2f} {rel(12):7.2f} {rel(13):7.2f} {rel(2):7.2f} {rel(6):7.2f} {rel(7):7.2f} {rel(3):7.2f}")
wa = sum((a[l,0,1]-a[l,0,0])*us for l in range(n) if a[l,0,1])
ae = sum((a[l,0,2]-max(a[l,0,1],a[l,0,0]))*us for l in range(n))
ed = sum((a[l,0,3]-a[l,0,2])*us for l in range(n))
print(f"phase sums rank 0 (us): wait->acc {wa:.1f}  acc->epi {ae:.1f}  epi->deliv {ed:.1f}  cluster {os.environ.get('MPDB_MEGA_DBG_CLUSTER', '0')}")
print(f"total (first start -> last delivered): {(a[n-1,:,3].max() - t0) * us:.1f} us")

# per-cluster entry / setup / exit on the GPU-wide timer (ns): slots 44..47 of the same buffer
full = np.array(buf[:]).reshape(48, 8, ND)
cl = full[44:48].reshape(-1)[:4 * 64].reshape(64, 4).astype(np.float64)
ids = np.nonzero(cl[:, 0] > 0)[0]
cl = cl[cl[:, 0] > 0]
if len(cl):
    print("per-cluster run (us):", " ".join(f"{i}:{(c[2]-c[1])/1e3:.1f}" for i, c in zip(ids, cl)))
    t0g = cl[:, 0].min()
    print(f"clusters: {len(cl)}; kernel entry spread {cl[:,0].max()-t0g:.0f} ns; setup done after entry: "
          f"{(cl[:,1]-cl[:,0]).min():.0f}..{(cl[:,1]-cl[:,0]).max():.0f} ns; exit after first entry: "
          f"{(cl[:,2]-t0g).min()/1e3:.1f}..{(cl[:,2]-t0g).max()/1e3:.1f} us; per-cluster run (setup->exit): "
          f"{((cl[:,2]-cl[:,1])/1e3).min():.1f}..{((cl[:,2]-cl[:,1])/1e3).max():.1f} us")
