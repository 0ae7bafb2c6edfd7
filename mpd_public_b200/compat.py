"""Alias layer: makes the reference's OWN import lines resolve to this framework, so that the body of
`scripts/inference/inference.py` (reference lines 76-282: seed, dataset, model, start / goal, costs, guide, `run_inference`,
the prior-then-guide post-loop) runs with no edit.

    import mpd_public_b200.compat as compat
    compat.install()                                   # registers mpd.*, mp_baselines.*, torch_robotics.*, experiment_launcher
    from mpd.models import TemporalUnet, UNET_DIM_MULTS                                            # inference.py:16
    from mpd.models.diffusion_models.guides import GuideManagerTrajectoriesWithVelocity           # :17
    from mpd.models.diffusion_models.sample_functions import guide_gradient_steps, ddpm_sample_fn # :18
    from mpd.trainer import get_dataset, get_model                                                 # :19
    from mp_baselines.planners.costs.cost_functions import CostCollision, CostComposite, CostGPTrajectory  # :15
    from torch_robotics.torch_utils.seed import fix_random_seed                                    # :22 ...

What stands behind each name:
  * `mpd.models.*`, the guide managers, the sample functions, the three cost classes: this package's implementations
    (CUDA kernels underneath), same signatures;
  * `mpd.trainer.get_model`: the reference's lookup `getattr(mpd.models, model_class)(**kwargs).to(device)`
    (train_loaders.py:12-25); `get_dataset` returns `(train_subset, train_dataloader, val_subset, val_dataloader)` with
    `train_subset.dataset` = this package's `TrajectoryDataset` (env / robot / task / normaliser; built from the on-disk
    dataset directory when there is one — `ingest.load_trajectory_limits` — else from the seeded synthetic problem of that
    environment id, since the reference's datasets are downloads, README.md:68-73);
  * `torch_robotics.*`: the handful of helpers inference.py touches (`fix_random_seed`, `TimerCUDA`, `get_torch_device`,
    `freeze_torch_model_params`, the three trajectory metrics) restated in a few lines each — torch_robotics itself is not in
    the reference tree;
  * `experiment_launcher.single_experiment_yaml / run_experiment`: call-through decorators (the launcher is out of scope).
Nothing here is on the timed path. `install()` refuses to shadow a real installation of any of these packages unless
`force=True`.
"""
from __future__ import annotations

import importlib.util
import random
import sys
import time
import types

import numpy as np
import torch

_INSTALLED = {}

DEFAULT_DATASET_DIR = None  # set to a directory holding <env-robot id>/**/trajs-free.pt to use real normaliser limits


def fix_random_seed(seed):
    """torch_robotics.torch_utils.seed.fix_random_seed (call site inference.py:78)"""
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)


def get_torch_device(device="cuda"):
    """torch_robotics.torch_utils.torch_utils.get_torch_device (inference.py:80)"""
    if "cuda" in str(device) and torch.cuda.is_available():
        return torch.device(device)
    if "cuda" in str(device):
        raise RuntimeError("mpd_public_b200 has no CPU path: a CUDA device is required")
    return torch.device(device)


def freeze_torch_model_params(model):
    """torch_robotics.torch_utils.torch_utils.freeze_torch_model_params (inference.py:152)"""
    for p in model.parameters():
        p.requires_grad = False
    model.is_frozen = True


class TimerCUDA:
    """torch_robotics.torch_utils.torch_timer.TimerCUDA (inference.py:248,265): wall time between two device syncs"""

    def __enter__(self):
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        self.start = time.perf_counter()
        self.elapsed = 0.0
        return self

    def __exit__(self, *exc):
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        self.elapsed = time.perf_counter() - self.start
        return False


class _Subset:
    """torch.utils.data.Subset stand-in: inference.py only reads `.dataset` (:114)"""

    def __init__(self, dataset):
        self.dataset = dataset

    def __len__(self):
        return 0


def get_dataset(dataset_class=None, dataset_subdir=None, batch_size=2, val_set_size=0.05, results_dir=None, save_indices=False,
                tensor_args=None, include_velocity=True, use_extra_objects=True, obstacle_cutoff_margin=0.05, **kwargs):
    """mpd.trainer.get_dataset as inference.py:107-113 calls it. `dataset_subdir` is the training run's environment / robot id
    (e.g. 'EnvSpheres3D-RobotPanda'); `n_support_points` may come through **kwargs (args.yaml)."""
    from . import synthetic as S
    from .planning import TrajectoryDataset
    if dataset_class not in (None, "TrajectoryDataset"):
        raise NotImplementedError(f"dataset_class {dataset_class}")
    if dataset_subdir not in S.MODEL_IDS:
        raise KeyError(f"unknown environment / robot id {dataset_subdir!r}; known: {sorted(S.MODEL_IDS)}")
    device = (tensor_args or {}).get("device", "cuda")
    h = int(kwargs.get("n_support_points", 64))
    prob = S.make_problem_by_id(dataset_subdir, h)
    ds = TrajectoryDataset(prob, device, include_velocity=include_velocity, use_extra_objects=use_extra_objects,
                           obstacle_cutoff_margin=obstacle_cutoff_margin)
    if DEFAULT_DATASET_DIR is not None:  # real limits when the dataset is on disk
        import os
        from .ingest import load_trajectory_limits
        d = os.path.join(DEFAULT_DATASET_DIR, dataset_subdir)
        if os.path.isdir(d):
            ds.normalizer, _h, _d = load_trajectory_limits(d, prob.robot.q_dim, include_velocity, device)
    sub = _Subset(ds)
    return sub, None, sub, None


def get_model(model_class=None, checkpoint_path=None, freeze_loaded_model=False, tensor_args=None, **kwargs):
    """mpd.trainer.get_model (train_loaders.py:12-25): `getattr(mpd.models, model_class)(**kwargs).to(device)`"""
    import mpd_public_b200 as M
    if checkpoint_path is not None:
        model = torch.load(checkpoint_path)
        if freeze_loaded_model:
            freeze_torch_model_params(model)
        return model
    return getattr(M, model_class)(**kwargs).to((tensor_args or {}).get("device", "cuda"))


def single_experiment_yaml(fn):
    """experiment_launcher.single_experiment_yaml: call-through (the launcher / YAML bookkeeping is out of scope)"""
    return fn


def run_experiment(fn, *args, **kwargs):
    return fn(*args, **kwargs)


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__dict__["__mpdb_alias__"] = True
    return m


def install(force=False):
    """Registers the alias modules in sys.modules (idempotent). Raises if a real `mpd` / `mp_baselines` / `torch_robotics`
    installation is importable and force is False — shadowing the reference silently would be the wrong default."""
    import mpd_public_b200 as M
    from . import ingest, planning
    if _INSTALLED:
        return dict(_INSTALLED)
    for top in ("mpd", "mp_baselines", "torch_robotics", "experiment_launcher"):
        present = sys.modules.get(top)
        if present is None:
            try:
                present = importlib.util.find_spec(top)
            except (ImportError, ValueError):
                present = None
        if present is not None and not getattr(present, "__mpdb_alias__", False) and not force:
            raise RuntimeError(f"a real '{top}' package is importable; call compat.install(force=True) to shadow it")
    models = _module("mpd.models", TemporalUnet=M.TemporalUnet, UNET_DIM_MULTS=M.UNET_DIM_MULTS,
                     GaussianDiffusionModel=M.GaussianDiffusionModel)
    guides = _module("mpd.models.diffusion_models.guides", GuideManagerTrajectoriesWithVelocity=M.GuideManagerTrajectoriesWithVelocity,
                     GuideManagerTrajectories=M.GuideManagerTrajectories)
    sfn = _module("mpd.models.diffusion_models.sample_functions", guide_gradient_steps=M.guide_gradient_steps,
                  ddpm_sample_fn=M.ddpm_sample_fn, apply_hard_conditioning=M.apply_hard_conditioning, extract=M.extract)
    dmb = _module("mpd.models.diffusion_models.diffusion_model_base", GaussianDiffusionModel=M.GaussianDiffusionModel,
                  make_timesteps=M.make_timesteps)
    dm = _module("mpd.models.diffusion_models", guides=guides, sample_functions=sfn, diffusion_model_base=dmb,
                 GaussianDiffusionModel=M.GaussianDiffusionModel, TemporalUnet=M.TemporalUnet)
    models.diffusion_models = dm
    trainer = _module("mpd.trainer", get_dataset=get_dataset, get_model=get_model)
    loading = _module("mpd.utils.loading", load_params_from_yaml=ingest.load_params_from_yaml)
    utils = _module("mpd.utils", loading=loading)
    norm = _module("mpd.datasets.normalization", LimitsNormalizer=M.LimitsNormalizer, DatasetNormalizer=M.DatasetNormalizer)
    datasets = _module("mpd.datasets", normalization=norm, TrajectoryDataset=planning.TrajectoryDataset)
    mpd = _module("mpd", models=models, trainer=trainer, utils=utils, datasets=datasets)
    costs = _module("mp_baselines.planners.costs.cost_functions", CostCollision=M.CostCollision, CostComposite=M.CostComposite,
                    CostGPTrajectory=M.CostGPTrajectory)
    mpb_costs = _module("mp_baselines.planners.costs", cost_functions=costs)
    mpb_pl = _module("mp_baselines.planners", costs=mpb_costs)
    mpb = _module("mp_baselines", planners=mpb_pl)
    tr_seed = _module("torch_robotics.torch_utils.seed", fix_random_seed=fix_random_seed)
    tr_timer = _module("torch_robotics.torch_utils.torch_timer", TimerCUDA=TimerCUDA)
    tr_tu = _module("torch_robotics.torch_utils.torch_utils", get_torch_device=get_torch_device,
                    freeze_torch_model_params=freeze_torch_model_params,
                    to_numpy=lambda x, **k: x.detach().cpu().numpy(),
                    to_torch=lambda x, **k: x.to(**k) if torch.is_tensor(x) else torch.tensor(x, **k))
    tr_utils = _module("torch_robotics.torch_utils", seed=tr_seed, torch_timer=tr_timer, torch_utils=tr_tu)
    tr_metrics = _module("torch_robotics.trajectory.metrics", compute_smoothness=planning.compute_smoothness,
                         compute_path_length=planning.compute_path_length,
                         compute_variance_waypoints=planning.compute_variance_waypoints)
    tr_traj = _module("torch_robotics.trajectory", metrics=tr_metrics)
    tr = _module("torch_robotics", torch_utils=tr_utils, trajectory=tr_traj)
    el = _module("experiment_launcher", single_experiment_yaml=single_experiment_yaml, run_experiment=run_experiment)
    mods = {m.__name__: m for m in (mpd, models, dm, guides, sfn, dmb, trainer, utils, loading, datasets, norm, mpb, mpb_pl, mpb_costs,
                                    costs, tr, tr_utils, tr_seed, tr_timer, tr_tu, tr_traj, tr_metrics, el)}
    for name, m in mods.items():
        sys.modules[name] = m
    _INSTALLED.update(mods)
    return dict(mods)


def uninstall():
    for name in list(_INSTALLED):
        if sys.modules.get(name) is _INSTALLED[name]:
            del sys.modules[name]
    _INSTALLED.clear()
