"""Host-side mirror of reference `mpd/models/diffusion_models/sample_functions.py:5-83`.

`ddpm_sample_fn` and `guide_gradient_steps` keep the reference signatures (inference.py:253,269-275 pass
them around by reference). With this package's model/guide the arithmetic runs in libmpdb200 kernels;
a foreign `guide` callable is still honoured through the generic torch path of `guide_gradient_steps`.
"""
from __future__ import annotations

import torch


def apply_hard_conditioning(x, conditions):
    """reference :5-8 — in place, returns its argument."""
    for t, val in conditions.items():
        x[:, t, :] = val.clone()
    return x


def extract(a, t, x_shape):
    """reference :11-14"""
    b, *_ = t.shape
    out = a.gather(-1, t)
    return out.reshape(b, *((1,) * (len(x_shape) - 1)))


@torch.no_grad()
def ddpm_sample_fn(model, x, hard_conds, context, t, guide=None, n_guide_steps=1, scale_grad_by_std=False,
                   t_start_guide=torch.inf, noise_std_extra_schedule_fn=None, debug=False, **kwargs):
    """One reverse step, reference :18-62. Returns (x_{t-1}, None)."""
    t_single = t[0]
    if t_single < 0:
        t = torch.zeros_like(t)

    model_mean, _, model_log_variance = model.p_mean_variance(x=x, hard_conds=hard_conds, context=context, t=t)
    x = model_mean

    model_log_variance = extract(model.posterior_log_variance_clipped, t, x.shape)
    model_var = torch.exp(model_log_variance)

    if guide is not None and t_single < t_start_guide:
        x = guide_gradient_steps(x, hard_conds=hard_conds, guide=guide, n_guide_steps=n_guide_steps,
                                 scale_grad_by_std=scale_grad_by_std, model_var=model_var, debug=False)

    # no noise when t == 0 (handled inside the kernel, as `noise[t == 0] = 0`)
    noise = torch.randn_like(x)
    if noise_std_extra_schedule_fn is None:
        noise_std = 1.0
    else:
        noise_std = noise_std_extra_schedule_fn(t_single)

    values = None
    x = model._engine(x.shape[1]).add_noise_(x if x.is_contiguous() else x.contiguous(), t, noise, float(noise_std))
    return x, values


def guide_gradient_steps(x, hard_conds=None, guide=None, n_guide_steps=1, scale_grad_by_std=False, model_var=None,
                         debug=False, **kwargs):
    """reference :65-83 — x <- x + guide(x) [* model_var], hard conditioning, n times; returns a new tensor."""
    hard_conds = hard_conds or {}
    if getattr(guide, "_mpdb_fusable", False):
        return guide.guide_steps(x, hard_conds, n_guide_steps, model_var if scale_grad_by_std else None)
    for _ in range(n_guide_steps):
        grad_scaled = guide(x)
        if scale_grad_by_std:
            grad_scaled = model_var * grad_scaled
        x = x + grad_scaled
        x = apply_hard_conditioning(x, hard_conds)
    return x
