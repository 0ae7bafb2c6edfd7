"""Shared definitions of the golden cases: inputs are regenerated from seeds on both sides
(generator in the build container, tests anywhere); only reference OUTPUTS are stored in the .npz files.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from mpd_public_b200 import synthetic as S  # noqa: E402

GOLDEN_DIR = os.path.dirname(os.path.abspath(__file__))

# name -> (state_dim, horizon, dim_mults option, weight seed)
UNET_CASES = {
    "panda_opt1_h64": (14, 64, 1, 0),
    "pm2d_opt0_h64": (4, 64, 0, 0),
    "pm2d_opt1_h64": (4, 64, 1, 5),
    "panda_opt1_h128": (14, 128, 1, 0),
}
UNET_T = [0, 7, 24]  # one t per sample (B = 3): exercises per-sample time conditioning

T_DIFF = 25
N_EXTRA = 5
T_START_GUIDE = 7      # ceil(0.25 * 25), reference inference.py:238
N_GUIDE_STEPS = 5
NOISE_STD = 0.5        # reference inference.py:243


def unet_weights(case):
    d, h, opt, seed = UNET_CASES[case]
    return S.make_unet_state_dict(seed, d, 32, S.UNET_DIM_MULTS[opt])


def unet_input(case, batch=3, seed=11):
    d, h, opt, _ = UNET_CASES[case]
    rng = np.random.default_rng([seed, d, h])
    return rng.standard_normal((batch, h, d)).astype(np.float32)


# guided cases: name -> (model id, unet case, grid cell, weight_collision, weight_smoothness, batch)
GUIDE_CASES = {
    "simple2d": ("EnvSimple2D-RobotPointMass", "pm2d_opt0_h64", 0.01, 3e-2, 1e-2, 4),
    "panda3d": ("EnvSpheres3D-RobotPanda", "panda_opt1_h64", 0.04, 1e-2, 1e-7, 2),
}
# steps i (as in p_sample_loop) at which the teacher-forced ddpm_sample_fn is recorded
STEP_LIST = [24, 23, 12, 6, 3, 0, -2]


def guide_problem(case):
    model_id, ucase, cell, wc, ws, batch = GUIDE_CASES[case]
    d, h, opt, _ = UNET_CASES[ucase]
    return S.make_problem_by_id(model_id, n_support_points=h, cell=cell)


def guide_input(case, seed=21, out_of_range=False):
    """Normalised trajectories near the straight start->goal line (so costs are active)."""
    model_id, ucase, cell, wc, ws, batch = GUIDE_CASES[case]
    prob = guide_problem(case)
    d, h = prob.robot.state_dim, prob.n_support_points
    rng = np.random.default_rng([seed, d, h, int(out_of_range)])
    s = np.concatenate([prob.start, np.zeros(prob.robot.q_dim)])
    g = np.concatenate([prob.goal, np.zeros(prob.robot.q_dim)])
    lam = np.linspace(0, 1, h)[None, :, None]
    x = (1 - lam) * s[None, None, :] + lam * g[None, None, :]
    x = 2 * (x - prob.mins) / (prob.maxs - prob.mins) - 1
    x = x + 0.15 * rng.standard_normal((batch, h, d))
    if not out_of_range:
        x = np.clip(x, -0.999, 0.999)
    else:
        x[0, 5, 0] = 1.3  # triggers the batch-global clip of LimitsNormalizer.unnormalize
    return x.astype(np.float32)


def pos_guide_input(case, seed=41):
    """Normalised POSITION-only trajectories near the straight start->goal line (GuideManagerTrajectories, guides.py:15-146);
    one entry out of range so that the normaliser's batch-global clip branch is taken."""
    model_id, ucase, cell, wc, ws, batch = GUIDE_CASES[case]
    prob = guide_problem(case)
    q, h = prob.robot.q_dim, prob.n_support_points
    rng = np.random.default_rng([seed, q, h])
    lam = np.linspace(0, 1, h)[None, :, None]
    x = (1 - lam) * prob.start[None, None, :] + lam * prob.goal[None, None, :]
    x = 2 * (x - prob.mins[:q]) / (prob.maxs[:q] - prob.mins[:q]) - 1
    x = x + 0.12 * rng.standard_normal((batch, h, q))
    x[1, 9, 0] = 1.2
    return x.astype(np.float32)


def step_input(case, i, seed=31):
    """x_t for the teacher-forced step at loop index i (scaled roughly like the marginal at t)."""
    model_id, ucase, cell, wc, ws, batch = GUIDE_CASES[case]
    d, h, opt, _ = UNET_CASES[ucase]
    rng = np.random.default_rng([seed, d, h, i + 100])
    if i >= 12:
        return rng.standard_normal((batch, h, d)).astype(np.float32)
    return (guide_input(case) + 0.05 * rng.standard_normal((batch, h, d))).astype(np.float32)


def step_noise(shape, i):
    """The noise `torch.randn_like` returns inside ddpm_sample_fn after torch.manual_seed(1000 + i)."""
    torch.manual_seed(1000 + i)
    return torch.randn(shape)


def load(name):
    return np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
