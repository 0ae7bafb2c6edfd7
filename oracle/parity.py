"""ORACLE-SIDE PARITY HARNESS — test infrastructure, not product code (same import rule as mpd_oracle.py: only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s CPU legs may import this).

`check_loop_per_step` is the north-star criterion made exact: every step of the CUDA path's own chain is re-done by the
oracle from the CUDA path's x_t with the same injected noise and must match x_{t-1} within `tol` (1e-3 relative,
`max|a-b| / max|b|`; t = T-1 has its own bound). The guided steps contain discontinuities — nearest-texel
lookup, hinge, nearest wall — where a 1e-6 difference in the mean flips a branch and moves an element by ~weight * |grad|.
Instead of tolerating such flips, the CUDA guide RECORDS its discrete decisions (mpdb_guide_record_decisions) and the oracle
(a) takes them over, which makes the compared function smooth, so the step is held to `tol` everywhere, and
(b) audits them: a recorded decision may differ from the oracle's own only where the oracle's deciding quantity sits on the
boundary (`mpd_oracle.audit_decisions`); anything else is a failure.
"""
from __future__ import annotations

import numpy as np
import torch

from . import mpd_oracle as O


def rel(a, b):
    a = a.detach().double().cpu().numpy() if torch.is_tensor(a) else np.asarray(a, dtype=np.float64)
    b = b.detach().double().cpu().numpy() if torch.is_tensor(b) else np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def run_recorded(model, guide, hard_cuda, noise_cuda, batch, H, n_diffusion_steps, n_extra, t_start_guide, n_guide_steps,
                 noise_std, scale_grad_by_std=False):
    """The guided loop through the public API with the guide recording its decisions (CUDA graphs are bypassed while
    recording). Returns (chain [S, B, H, D] on the CPU, decisions int32 [n_evals, B, n_costs, NI, n_spheres] on the CPU)."""
    n_guided = sum(1 for i in range(-n_extra, n_diffusion_steps) if i < t_start_guide)
    buf = guide.record_decisions(noise_cuda.device, H, batch, n_guided * n_guide_steps)
    try:
        chain = model.run_inference(None, hard_cuda, n_samples=batch, horizon=H, return_chain=True, noise=noise_cuda,
                                    guide=guide, n_guide_steps=n_guide_steps, t_start_guide=t_start_guide,
                                    scale_grad_by_std=scale_grad_by_std,
                                    noise_std_extra_schedule_fn=lambda _t: noise_std, n_diffusion_steps_without_noise=n_extra)
        torch.cuda.synchronize()
        n_rec = guide.decisions_recorded()
    finally:
        guide.stop_recording()
    assert n_rec == n_guided * n_guide_steps, (n_rec, n_guided, n_guide_steps)
    return chain.cpu(), buf[:n_rec].cpu()


def check_loop_per_step(om, spec, chain, noise, hard, decisions, n_diffusion_steps, n_extra, t_start_guide, n_guide_steps,
                        noise_std, tol=1e-3, tol_t_last=2e-2, scale_grad_by_std=False, audit=True, steps_subset=None):
    """chain: [S, B, H, D] (CPU) produced by the CUDA path from `noise` [S, B, H, D]; decisions: as returned by run_recorded
    (None for an unguided loop). Returns {"worst": worst relative step error for t < T-1, "t_last": the error at t = T-1,
    "audit": summed audit counters}. Raises AssertionError on any violation."""
    S, B, H, D = chain.shape
    ohc = {k: v[None].repeat(B, 1) for k, v in hard.items()}
    steps = list(reversed(range(-n_extra, n_diffusion_steps)))
    assert S == len(steps) + 1
    worst, t_last, e_idx = 0.0, None, 0
    totals = {"n": 0, "index_diff": 0, "hinge_diff": 0, "wall_diff": 0, "unexplained": 0}
    with torch.no_grad():
        for k, i in enumerate(steps):
            guided = decisions is not None and i < t_start_guide
            decs = None
            if guided:
                decs = decisions[e_idx:e_idx + n_guide_steps]
                e_idx += n_guide_steps
            if steps_subset is not None and i not in steps_subset:
                continue
            t = torch.full((B,), i, dtype=torch.long)
            oguide, pending = None, []
            if guided:
                calls = iter(range(n_guide_steps))

                def oguide(z, decs=decs, calls=calls, pending=pending):
                    j = next(calls)
                    pending.append((z.clone(), decs[j]))
                    return O.guide_manager_grad(spec, z, decisions=decs[j])
            ref = om.ddpm_step(chain[k].clone(), ohc, t, noise[k + 1], oguide, n_guide_steps, scale_grad_by_std, t_start_guide,
                               noise_std)
            ref = O.apply_hard_conditioning(ref, ohc)
            e = rel(chain[k + 1], ref)
            if guided and audit:
                # The two sides' iterates inside this step differ by about as much as their results do (relative e, measured just
                # above; the guide moves x by small steps): in joint space e * range / 2, at a sphere centre times the arm's reach.
                # Decisions may differ only within that distance of their boundary — 4x margin, 1e-5 m floor.
                reach = 1.2 if spec.robot.kind == "panda" else 1.0
                half_range = float(((spec.maxs - spec.mins) / 2)[:spec.robot.q_dim].max())
                pos_tol = 1e-5 + 4.0 * e * half_range * reach
                for j, (z, dj) in enumerate(pending):
                    a = O.audit_decisions(spec, z, dj, pos_tol=pos_tol)
                    for key in totals:
                        totals[key] += a[key]
                    assert a["unexplained"] == 0, (f"step i={i}, guide evaluation {j}: decisions of the CUDA guide that the oracle "
                                                   f"cannot explain by a boundary within {pos_tol:.2e} m", a)
            if i == n_diffusion_steps - 1:
                t_last = e
                assert e < tol_t_last, (i, e)
            else:
                assert e < tol, (i, e, "guided" if guided else "unguided")
                worst = max(worst, e)
    if decisions is not None:
        assert e_idx == decisions.shape[0], (e_idx, decisions.shape)
    return {"worst": worst, "t_last": t_last, "audit": totals}
