"""Error of the first reverse step (t = T-1, eps amplified 4602x) against the oracle, exact fp32 path vs tensor-core path,
and the oracle's own fp32-vs-fp64 spread on the same inputs (precision policy evidence)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oracle import mpd_oracle as O
from tests.golden import cases as C
from tests.test_gpu_parity import cuda_model, oracle_model, rel

for ucase, batch in (("panda_opt1_h64", 100), ("pm2d_opt0_h64", 64)):
    model = cuda_model(ucase)
    om = oracle_model(ucase)
    d, h, opt, seed = C.UNET_CASES[ucase]
    n_iters = C.T_DIFF + C.N_EXTRA
    gen = torch.Generator().manual_seed(5)
    noise = torch.randn((n_iters + 1, batch, h, d), generator=gen)
    hard = {0: torch.linspace(-0.5, 0.5, d), h - 1: torch.linspace(0.4, -0.4, d)}
    ohc = {k: v[None].repeat(batch, 1) for k, v in hard.items()}
    hard_cuda = {k: v.cuda() for k, v in hard.items()}
    kw = dict(n_diffusion_steps_without_noise=C.N_EXTRA, noise_std_extra_schedule_fn=lambda _t: C.NOISE_STD)
    t = torch.full((batch,), C.T_DIFF - 1, dtype=torch.long)
    out = {}
    for tc in ("off", "force"):
        model.tensor_cores = tc
        chain = model.run_inference(None, hard_cuda, n_samples=batch, horizon=h, return_chain=True, noise=noise.cuda(), **kw).cpu()
        with torch.no_grad():
            ref = om.ddpm_step(chain[0].clone(), ohc, t, noise[1], None, C.N_GUIDE_STEPS, False, C.T_START_GUIDE, C.NOISE_STD)
        ref = O.apply_hard_conditioning(ref, ohc)
        dlt = (chain[1] - ref).abs() / ref.abs().max()
        out[tc] = chain
        print(f"[{ucase} B={batch}] tc={tc}: step t=T-1 rel err vs oracle(fp32) {rel(chain[1], ref):.3e}; elements > 1e-3: {int((dlt > 1e-3).sum())} of {dlt.numel()}; "
              f"final-sample difference vs exact-path loop: {rel(chain[-1], out['off'][-1]):.3e}")
    model.tensor_cores = "auto"
    # the oracle's own fp32 vs fp64 spread at this step
    om64 = O.OracleDiffusion({k: torch.as_tensor(v).double() for k, v in C.unet_weights(ucase).items()}, n_diffusion_steps=C.T_DIFF) if hasattr(O, "OracleDiffusion") else None
    try:
        with torch.no_grad():
            r32 = om.ddpm_step(out["off"][0].clone(), ohc, t, noise[1], None, C.N_GUIDE_STEPS, False, C.T_START_GUIDE, C.NOISE_STD)
            r64 = om64.ddpm_step(out["off"][0].clone().double(), {k: v.double() for k, v in ohc.items()}, t, noise[1].double(), None, C.N_GUIDE_STEPS, False, C.T_START_GUIDE, C.NOISE_STD)
        d2 = (r32.double() - r64).abs() / r64.abs().max()
        print(f"[{ucase}] oracle fp32 vs fp64 at t=T-1: rel {float(d2.max()):.3e}; elements > 1e-3: {int((d2 > 1e-3).sum())}")
    except Exception as e:
        print("fp64 oracle comparison unavailable:", repr(e)[:200])
