"""ORACLE — test infrastructure. Name-only shim that lets the reference's own diffusion core import.

Works only where `/root/reference` exists (the build container); it never travels to the GPU box and
nothing in `-m gpu` tests, `smoke()` or `bench.py` uses it. `tests/golden/make_golden.py` uses it to
generate the committed golden vectors, and `tests/test_golden_reproducible.py` (skipped when the
reference is absent) re-runs the generators and compares with the committed fixtures.

The shim supplies *names* for modules that are missing here (matplotlib, torch_robotics,
mp_baselines — SURVEY.md Appendix D); every number still comes from reference code. The single
exception is `interpolate_points_v1`, whose source is absent: it is bound to the oracle's restatement.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("MPD_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "mpd"))


def _mod(name, **attrs):
    m = sys.modules.get(name)
    if m is None:
        m = types.ModuleType(name)
        m.__path__ = []  # behave like a package
        sys.modules[name] = m
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


_loaded = None


def load():
    """Returns a namespace with the reference's TemporalUnet, GaussianDiffusionModel, ddpm_sample_fn,
    guide_gradient_steps, GuideManagerTrajectoriesWithVelocity, LimitsNormalizer, UNET_DIM_MULTS."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("reference tree not present")
    from oracle import mpd_oracle

    try:
        import matplotlib  # noqa: F401
    except Exception:
        _mod("matplotlib")
        _mod("matplotlib.pyplot")
    _mod("torch_robotics")
    _mod("torch_robotics.torch_utils")
    _mod("torch_robotics.torch_utils.torch_timer", TimerCUDA=object)
    _mod("torch_robotics.torch_utils.torch_utils",
         to_numpy=lambda x: x.detach().cpu().numpy(),
         to_torch=lambda x, **k: x.to(**k) if torch.is_tensor(x) else torch.tensor(x, **k))
    _mod("torch_robotics.torch_planning_objectives")
    _mod("torch_robotics.torch_planning_objectives.fields")
    _mod("torch_robotics.torch_planning_objectives.fields.distance_fields",
         interpolate_points_v1=lambda x, num_interpolated_points=128: mpd_oracle.interpolate_points(x, num_interpolated_points))
    _mod("mp_baselines")
    _mod("mp_baselines.planners")
    _mod("mp_baselines.planners.costs")
    _mod("mp_baselines.planners.costs.cost_functions", CostGPTrajectory=object)
    _mod("mp_baselines.planners.costs.factors")
    _mod("mp_baselines.planners.costs.factors.mp_priors_multi", MultiMPPrior=object)

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        from mpd.models import TemporalUnet, UNET_DIM_MULTS, GaussianDiffusionModel
        from mpd.models.diffusion_models.sample_functions import ddpm_sample_fn, guide_gradient_steps
        from mpd.models.diffusion_models.guides import GuideManagerTrajectoriesWithVelocity
    spec = importlib.util.spec_from_file_location(
        "_ref_normalization", os.path.join(REFERENCE_ROOT, "mpd", "datasets", "normalization.py"))
    norm = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(norm)

    ns = types.SimpleNamespace(
        TemporalUnet=TemporalUnet, UNET_DIM_MULTS=UNET_DIM_MULTS, GaussianDiffusionModel=GaussianDiffusionModel,
        ddpm_sample_fn=ddpm_sample_fn, guide_gradient_steps=guide_gradient_steps,
        GuideManagerTrajectoriesWithVelocity=GuideManagerTrajectoriesWithVelocity,
        LimitsNormalizer=norm.LimitsNormalizer)
    _loaded = ns
    return ns


def build_reference_model(unet_sd_numpy, state_dim, n_support_points, unet_input_dim=32, dim_mults=(1, 2, 4, 8),
                          n_diffusion_steps=25, variance_schedule="exponential", predict_epsilon=True):
    """Reference GaussianDiffusionModel(TemporalUnet) with the given weights loaded strictly."""
    import contextlib
    import io
    ref = load()
    with contextlib.redirect_stdout(io.StringIO()):
        unet = ref.TemporalUnet(n_support_points=n_support_points, state_dim=state_dim,
                                unet_input_dim=unet_input_dim, dim_mults=dim_mults)
    unet.load_state_dict({k: torch.as_tensor(v) for k, v in unet_sd_numpy.items()}, strict=True)
    model = ref.GaussianDiffusionModel(model=unet, variance_schedule=variance_schedule,
                                       n_diffusion_steps=n_diffusion_steps, predict_epsilon=predict_epsilon)
    model.eval()
    return model


class RefDatasetStub:
    """The two attributes of TrajectoryDataset the guide touches (guides.py:180)."""

    def __init__(self, mins, maxs):
        ref = load()
        self.normalizer = ref.LimitsNormalizer(torch.stack([torch.as_tensor(mins), torch.as_tensor(maxs)]))

    def unnormalize_trajectories(self, x):
        return self.normalizer.unnormalize(x)

    def normalize_trajectories(self, x):
        return self.normalizer.normalize(x)


class RefCostStub:
    """Callable with CostComposite's call signature (guides.py:190), arithmetic = the oracle's restatement."""

    def __init__(self, guide_spec):
        self.spec = guide_spec

    def __call__(self, trajs, x_interpolated=None, return_invidual_costs_and_weights=False, **kw):
        from oracle import mpd_oracle
        costs, weights = mpd_oracle.composite_costs(self.spec, trajs, x_interpolated)
        if return_invidual_costs_and_weights:
            return costs, weights
        return sum(w * c for c, w in zip(costs, weights))


def build_reference_guide(guide_spec):
    ref = load()
    return ref.GuideManagerTrajectoriesWithVelocity(
        RefDatasetStub(guide_spec.mins, guide_spec.maxs), RefCostStub(guide_spec),
        clip_grad=guide_spec.clip_grad, max_grad_norm=guide_spec.max_grad_norm,
        interpolate_trajectories_for_collision=guide_spec.interpolate,
        num_interpolated_points_for_collision=guide_spec.n_interp)


def build_reference_pos_guide(guide_spec, start_pos, goal_pos, dt, num_steps, n_samples):
    """The reference's position-only `GuideManagerTrajectories` (guides.py:15-146) with name-only stand-ins for what is
    absent: `MultiMPPrior.const_vel_trajectory` -> the oracle's restatement (source absent: unpinned), `robot.get_velocity`
    -> the velocity half of the state, the cost -> RefCostStub. Every line of the manager itself (velocity trajectory,
    separate clipping of the position / velocity gradients, end-row zeroing, weighting, velocity update) is reference code.
    The manager calls `torch.autograd.grad(cost.sum(), [x_pos, velocity])` once per cost without `retain_graph` /
    `allow_unused` (guides.py:88), so the reference itself runs only with ONE cost that reads both halves of the state
    (a composite of several costs raises inside autograd): build it with a GP-only spec."""
    load()
    from oracle import mpd_oracle
    gm = sys.modules["mpd.models.diffusion_models.guides"]
    q = guide_spec.robot.q_dim

    class _Prior:
        @staticmethod
        def const_vel_trajectory(start, goal, dt_, n, q_dim, set_initial_final_vel_to_zero=False, tensor_args=None):
            return mpd_oracle.const_vel_trajectory(start, goal, dt_, n, q_dim, set_initial_final_vel_to_zero)

    class _Robot:
        q_dim = q

        def __init__(self):
            self.dt = dt

        def get_velocity(self, x):
            return x[..., q:]

    gm.MultiMPPrior = _Prior
    return gm.GuideManagerTrajectories(
        RefDatasetStub(guide_spec.mins[:q], guide_spec.maxs[:q]), RefCostStub(guide_spec), clip_grad=guide_spec.clip_grad,
        max_grad_norm=guide_spec.max_grad_norm, interpolate_trajectories_for_collision=guide_spec.interpolate,
        num_interpolated_points_for_collision=guide_spec.n_interp, start_state_pos=torch.as_tensor(start_pos),
        goal_state_pos=torch.as_tensor(goal_pos), num_steps=num_steps, robot=_Robot(), n_samples=n_samples,
        tensor_args=dict(device="cpu", dtype=torch.float32))

