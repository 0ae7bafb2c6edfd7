"""Real checkpoint / dataset ingestion — the files a trained mpd-public run leaves behind:

    <model_dir>/args.yaml                                   training arguments (inference.py:103)
    <model_dir>/checkpoints/ema_model_current_state_dict.pth  or  model_current_state_dict.pth (inference.py:145-148)
    <dataset_dir>/**/trajs-free.pt                          the trajectories the normaliser limits come from
                                                             (trajectories.py:82-110, normalization.py:82-85)

None of them can be downloaded here (README.md:68-73), so this module is exercised by a round-trip test on a
synthetic directory with the same layout.
"""
from __future__ import annotations

import os

import torch
import yaml

from .diffusion_model import GaussianDiffusionModel
from .normalization import DatasetNormalizer, LimitsNormalizer
from .synthetic import UNET_DIM_MULTS
from .temporal_unet import TemporalUnet


def load_params_from_yaml(path):
    """reference mpd/utils/loading.py:4"""
    with open(path, "r") as f:
        return yaml.load(f, Loader=yaml.Loader)


def load_diffusion_model(model_dir, state_dim, n_support_points, device="cuda", verbose=False):
    """Builds GaussianDiffusionModel(TemporalUnet) from `args.yaml` exactly as inference.py:127-149 does and loads the
    (EMA) state dict. Returns (model.eval() on `device`, args)."""
    args = load_params_from_yaml(os.path.join(model_dir, "args.yaml"))
    diffusion_configs = dict(variance_schedule=args["variance_schedule"], n_diffusion_steps=args["n_diffusion_steps"],
                             predict_epsilon=args["predict_epsilon"])
    unet_configs = dict(state_dim=state_dim, n_support_points=n_support_points, unet_input_dim=args["unet_input_dim"],
                        dim_mults=UNET_DIM_MULTS[args["unet_dim_mults_option"]])
    if args.get("diffusion_model_class", "GaussianDiffusionModel") != "GaussianDiffusionModel":
        raise NotImplementedError(f"diffusion_model_class {args['diffusion_model_class']}")
    import contextlib
    import io
    with contextlib.nullcontext() if verbose else contextlib.redirect_stdout(io.StringIO()):
        unet = TemporalUnet(**unet_configs)
    model = GaussianDiffusionModel(model=unet, **diffusion_configs, **unet_configs)
    name = "ema_model_current_state_dict.pth" if args.get("use_ema", True) else "model_current_state_dict.pth"
    sd = torch.load(os.path.join(model_dir, "checkpoints", name), map_location="cpu")
    model.load_state_dict(sd)  # strict: same keys and shapes as the reference module
    return model.to(device).eval(), args


def load_trajectory_limits(dataset_dir, q_dim, include_velocity=True, device="cpu"):
    """Walks `dataset_dir` for trajs-free.pt files (trajectories.py:86-110) and returns (DatasetNormalizer over the
    'traj' field, n_support_points, state_dim)."""
    trajs = []
    for cur, _sub, files in os.walk(dataset_dir, topdown=True):
        if "trajs-free.pt" in files:
            trajs.append(torch.load(os.path.join(cur, "trajs-free.pt"), map_location=device))
    if not trajs:
        raise FileNotFoundError(f"no trajs-free.pt under {dataset_dir}")
    t = torch.cat(trajs)
    if not include_velocity:
        t = t[..., :q_dim]
    b, h, d = t.shape
    return DatasetNormalizer({"traj": t}, LimitsNormalizer), h, d
