"""Generates tests/golden/*.npz by running the REFERENCE's own code (from /root/reference, under the
name-only shim of oracle/ref_shim.py). Run in the build container: `python tests/golden/make_golden.py`.

The reference holds no tests or golden vectors for this path, so these files are the
pin: UNet / schedule / p_mean_variance / ddpm_sample_fn / guide manager / p_sample_loop outputs are
produced by reference code. The cost arithmetic under the guide (collision/GP/FK/SDF — sources absent)
is the oracle's restatement bound into the reference guide manager ("parity unpinned" part).
"""
from __future__ import annotations

import dataclasses
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import mpd_oracle as O  # noqa: E402
from oracle import ref_shim  # noqa: E402
from tests.golden import cases as C  # noqa: E402
from mpd_public_b200 import synthetic as S  # noqa: E402

torch.set_num_threads(8)


def save(name, **arrs):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **{k: np.asarray(v) for k, v in arrs.items()})
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.1f} KiB)")


def ref_model(case):
    d, h, opt, seed = C.UNET_CASES[case]
    return ref_shim.build_reference_model(C.unet_weights(case), d, h, 32, S.UNET_DIM_MULTS[opt], C.T_DIFF)


def gen_unet():
    out = {}
    for case in C.UNET_CASES:
        model = ref_model(case)
        x = torch.as_tensor(C.unet_input(case))
        t = torch.tensor(C.UNET_T, dtype=torch.long)
        with torch.no_grad():
            out[case] = model.model(x, t, None).numpy()
    save("unet_eps", **out)


def gen_schedule():
    ref = ref_shim.load()
    out = {}
    for sched, T in (("exponential", 25), ("cosine", 20)):
        class _M(torch.nn.Module):
            state_dim = 4
        m = ref.GaussianDiffusionModel(model=_M(), variance_schedule=sched, n_diffusion_steps=T, predict_epsilon=True)
        for k, v in m.state_dict().items():
            out[f"{sched}.{k}"] = v.numpy()
    save("schedule", **out)


def gen_normalizer():
    ref = ref_shim.load()
    prob = C.guide_problem("panda3d")
    nz = ref.LimitsNormalizer(torch.stack([torch.as_tensor(prob.mins), torch.as_tensor(prob.maxs)]))
    x_in = torch.as_tensor(C.guide_input("panda3d"))
    x_out = torch.as_tensor(C.guide_input("panda3d", out_of_range=True))
    save("normalizer", un_in=nz.unnormalize(x_in).numpy(), un_out=nz.unnormalize(x_out).numpy(),
         renorm=nz.normalize(nz.unnormalize(x_in)).numpy())


def gen_guide_and_steps():
    ref = ref_shim.load()
    for case, (model_id, ucase, cell, wc, ws, batch) in C.GUIDE_CASES.items():
        prob = C.guide_problem(case)
        spec = O.make_guide_spec(prob, wc, ws)
        guide = ref_shim.build_reference_guide(spec)
        model = ref_model(ucase)
        hard = O.hard_conditions(prob)
        out = {}
        # guide manager alone
        for tag, oor in (("in", False), ("oor", True)):
            x = torch.as_tensor(C.guide_input(case, out_of_range=oor))
            out[f"guide_grad_{tag}"] = guide(x).numpy()
        # guide_gradient_steps (5 iterations)
        x = torch.as_tensor(C.guide_input(case))
        hc = {k: v[None].repeat(batch, 1) for k, v in hard.items()}
        out["guide_steps5"] = ref.guide_gradient_steps(x.clone(), hard_conds=hc, guide=guide, n_guide_steps=5).numpy()
        # teacher-forced single steps
        for i in C.STEP_LIST:
            x = torch.as_tensor(C.step_input(case, i))
            t = torch.full((batch,), i, dtype=torch.long)
            with torch.no_grad():
                mean, _, _ = model.p_mean_variance(x=x.clone(), hard_conds=hc, context=None, t=torch.clamp(t, min=0))
            out[f"mean_{i}"] = mean.numpy()
            for gtag, g in (("noguide", None), ("guide", guide)):
                torch.manual_seed(1000 + i)
                xn, _ = ref.ddpm_sample_fn(model, x.clone(), hc, None, t, guide=g, n_guide_steps=C.N_GUIDE_STEPS,
                                           t_start_guide=C.T_START_GUIDE, noise_std_extra_schedule_fn=lambda _t: C.NOISE_STD)
                out[f"step_{gtag}_{i}"] = xn.numpy()
        # full loops (reference run_inference, RNG = global torch generator)
        for gtag, g in (("noguide", None), ("guide", guide)):
            torch.manual_seed(77)
            chain = model.run_inference(None, hard, n_samples=batch, horizon=prob.n_support_points, return_chain=True,
                                        sample_fn=ref.ddpm_sample_fn, guide=g, n_guide_steps=C.N_GUIDE_STEPS,
                                        t_start_guide=C.T_START_GUIDE, noise_std_extra_schedule_fn=lambda _t: C.NOISE_STD,
                                        n_diffusion_steps_without_noise=C.N_EXTRA)
            out[f"loop_chain_{gtag}"] = chain.numpy()
        save(f"guided_{case}", **out)


def gen_ddim():
    """`ddim_sample` of the reference (diffusion_model_base.py:184-259) through `conditional_sample(ddim=True)`: T // 5 steps,
    RNG = global torch generator (randn(shape), then one randn_like per step). Separate files so that adding them leaves the
    other fixtures untouched."""
    ref = ref_shim.load()
    for case, (model_id, ucase, cell, wc, ws, batch) in C.GUIDE_CASES.items():
        prob = C.guide_problem(case)
        spec = O.make_guide_spec(prob, wc, ws)
        guide = ref_shim.build_reference_guide(spec)
        model = ref_model(ucase)
        hard = O.hard_conditions(prob)
        hc = {k: v[None].repeat(batch, 1) for k, v in hard.items()}
        out = {}
        for gtag, g in (("noguide", None), ("guide", guide)):
            torch.manual_seed(78)
            x, chain = model.conditional_sample(hc, horizon=prob.n_support_points, batch_size=batch, ddim=True,
                                                return_chain=True, guide=g, n_guide_steps=C.N_GUIDE_STEPS,
                                                t_start_guide=C.T_START_GUIDE)
            out[f"chain_{gtag}"] = chain.numpy()
        save(f"ddim_{case}", **out)


def gen_scale_grad():
    """`scale_grad_by_std=True` (sample_functions.py:45-50, :77-78): the guide gradient multiplied by the posterior variance.
    One guided teacher-forced step per case through the reference's `ddpm_sample_fn`, and `guide_gradient_steps` with an
    explicit per-sample `model_var`. Separate files: the other fixtures stay as committed."""
    ref = ref_shim.load()
    for case, (model_id, ucase, cell, wc, ws, batch) in C.GUIDE_CASES.items():
        prob = C.guide_problem(case)
        spec = O.make_guide_spec(prob, wc, ws)
        guide = ref_shim.build_reference_guide(spec)
        model = ref_model(ucase)
        hard = O.hard_conditions(prob)
        hc = {k: v[None].repeat(batch, 1) for k, v in hard.items()}
        out = {}
        x = torch.as_tensor(C.guide_input(case))
        var = torch.linspace(0.2, 1.7, batch).reshape(batch, 1, 1)
        out["guide_steps2_var"] = ref.guide_gradient_steps(x.clone(), hard_conds=hc, guide=guide, n_guide_steps=2,
                                                           scale_grad_by_std=True, model_var=var).numpy()
        for i in (3, 0):
            x = torch.as_tensor(C.step_input(case, i))
            t = torch.full((batch,), i, dtype=torch.long)
            torch.manual_seed(1000 + i)
            xn, _ = ref.ddpm_sample_fn(model, x.clone(), hc, None, t, guide=guide, n_guide_steps=C.N_GUIDE_STEPS,
                                       scale_grad_by_std=True, t_start_guide=C.T_START_GUIDE,
                                       noise_std_extra_schedule_fn=lambda _t: C.NOISE_STD)
            out[f"step_scaled_{i}"] = xn.numpy()
        save(f"scale_grad_{case}", **out)


def gen_predict_x0():
    """`predict_epsilon=False` (diffusion_model_base.py:121-155): the model output is x0 itself. `p_mean_variance` at a few
    steps and one unguided `ddpm_sample_fn` step per case."""
    ref = ref_shim.load()
    out = {}
    for case, (model_id, ucase, cell, wc, ws, batch) in C.GUIDE_CASES.items():
        d, h, opt, seed = C.UNET_CASES[ucase]
        model = ref_shim.build_reference_model(C.unet_weights(ucase), d, h, 32, S.UNET_DIM_MULTS[opt], C.T_DIFF,
                                               predict_epsilon=False)
        prob = C.guide_problem(case)
        hc = {k: v[None].repeat(batch, 1) for k, v in O.hard_conditions(prob).items()}
        for i in (24, 12, 3, 0):
            x = torch.as_tensor(C.step_input(case, i))
            t = torch.full((batch,), i, dtype=torch.long)
            with torch.no_grad():
                mean, _, _ = model.p_mean_variance(x=x.clone(), hard_conds=hc, context=None, t=t)
            out[f"{case}.mean_{i}"] = mean.numpy()
            torch.manual_seed(1000 + i)
            xn, _ = ref.ddpm_sample_fn(model, x.clone(), hc, None, t)
            out[f"{case}.step_{i}"] = xn.numpy()
    save("predict_x0", **out)


def gen_pos_guide():
    """The position-only `GuideManagerTrajectories` (guides.py:15-146): gradient and velocity trajectory after each of three
    consecutive calls of the reference manager (x advanced by the returned gradient, as `guide_gradient_steps` does)."""
    for case, (model_id, ucase, cell, wc, ws, batch) in C.GUIDE_CASES.items():
        prob = C.guide_problem(case)
        # GP prior only: the reference manager differentiates each cost once without retain_graph / allow_unused, so it runs
        # with a single cost that reads positions and velocities (see ref_shim.build_reference_pos_guide)
        spec = dataclasses.replace(O.make_guide_spec(prob, wc, ws), grid_fields=[], border_limits=None, self_pairs=None)
        h = prob.n_support_points
        guide = ref_shim.build_reference_pos_guide(spec, prob.start, prob.goal, prob.dt, h - 1, batch)
        out = {"velocity_init": guide.velocity.detach().numpy()}
        x = torch.as_tensor(C.pos_guide_input(case))
        for call in range(3):
            g = guide(x)
            out[f"grad_{call}"] = g.detach().numpy()
            out[f"velocity_{call}"] = guide.velocity.detach().numpy()
            x = x + g.detach()
        save(f"pos_guide_{case}", **out)


def gen_state_dict_keys():
    import json
    out = {}
    for case in C.UNET_CASES:
        out[case] = {k: list(v.shape) for k, v in ref_model(case).state_dict().items()}
    path = os.path.join(HERE, "state_dict_keys.json")
    json.dump(out, open(path, "w"))
    print(f"wrote {path}")


if __name__ == "__main__":
    assert ref_shim.available(), "reference tree not present"
    if sys.argv[1:] == ["ddim"]:  # add the DDIM fixtures only (the others are left as committed)
        gen_ddim()
        sys.exit(0)
    if sys.argv[1:] == ["scale_grad"]:
        gen_scale_grad()
        sys.exit(0)
    if sys.argv[1:] == ["pos_guide"]:
        gen_pos_guide()
        sys.exit(0)
    if sys.argv[1:] == ["predict_x0"]:
        gen_predict_x0()
        sys.exit(0)
    gen_state_dict_keys()
    gen_schedule()
    gen_unet()
    gen_normalizer()
    gen_guide_and_steps()
    gen_ddim()
    gen_scale_grad()
    gen_predict_x0()
    gen_pos_guide()
