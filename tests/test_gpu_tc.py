"""tcgen05 implicit-GEMM core (csrc/unet_tc.cu) against torch conv1d: fp16-split (3 MMAs) must reproduce the
fp32 convolution to ~1e-5 relative. Raw accumulators, no epilogue."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,CI,CO,L", [(5, 32, 32, 8), (23, 64, 64, 16), (7, 128, 256, 8), (3, 32, 32, 64),
                                       (4, 64, 32, 32), (2, 32, 64, 128), (100, 256, 256, 8)])
def test_tc_conv5_raw(B, CI, CO, L):
    from mpd_public_b200 import _lib
    lib = _lib.lib()
    g = torch.Generator().manual_seed(B * 1000 + CI + L)
    x = torch.randn((B, CI, L), generator=g)
    w = torch.randn((CO, CI, 5), generator=g) / (CI * 5) ** 0.5
    ref = F.conv1d(x.double(), w.double(), padding=2).float()
    xcm = torch.zeros((B, CI, L + 4))
    xcm[:, :, 2:L + 2] = x
    xcm, wd = xcm.cuda().contiguous(), w.cuda().contiguous()
    SPT = 132 // (L + 4)
    tiles = (B + SPT - 1) // SPT
    raw = torch.full((tiles, CO // 32, 128, 32), float("nan"), device="cuda")
    _lib.check(lib.mpdb_debug_tc_conv5(_lib.fptr(xcm), _lib.fptr(wd), _lib.fptr(raw), B, CI, CO, L,
                                       _lib.stream_ptr(torch.device("cuda"))))
    torch.cuda.synchronize()
    raw = raw.cpu()
    out = torch.empty((B, CO, L))
    for b in range(B):
        t, s = divmod(b, SPT)
        rows = raw[t, :, s * (L + 4): s * (L + 4) + L, :]          # [CO/32, L, 32]
        out[b] = rows.permute(0, 2, 1).reshape(CO, L)
    err = float((out - ref).abs().max() / ref.abs().max())
    assert err < 1e-5, err  # 22-bit fp16 split: measured ~1e-6
