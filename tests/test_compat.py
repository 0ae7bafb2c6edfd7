"""The alias layer (mpd_public_b200/compat.py): the reference's own import lines resolve to this framework and the calls
`scripts/inference/inference.py` makes bind to our signatures.

CPU part: when the reference tree is present (build container), inference.py is PARSED (never copied): every name it imports
from mpd / mp_baselines / torch_robotics / experiment_launcher for lines 76-326 must resolve through the aliases, and every
keyword it passes to those callables (and to `model.run_inference` / `dataset.get_hard_conditions` / `task.*`) must be
accepted. A committed list of the imported names keeps the check alive where the reference tree is absent (GPU box).
GPU part: examples/inference_like_reference.py — the same sequence of calls as inference.py:76-326, written against the
reference's module paths — runs for both robots and all three planner algorithms."""
import ast
import importlib
import inspect
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/scripts/inference/inference.py"

# (module, name) pairs inference.py imports for the sampling path (lines 13-26). Rendering / simulation imports
# (isaac_gym_envs, RobotPanda, interpolate_traj_via_points, PlanningVisualizer) are out of scope.
IMPORTED = [
    ("experiment_launcher", "single_experiment_yaml"), ("experiment_launcher", "run_experiment"),
    ("mp_baselines.planners.costs.cost_functions", "CostCollision"), ("mp_baselines.planners.costs.cost_functions", "CostComposite"),
    ("mp_baselines.planners.costs.cost_functions", "CostGPTrajectory"),
    ("mpd.models", "TemporalUnet"), ("mpd.models", "UNET_DIM_MULTS"),
    ("mpd.models.diffusion_models.guides", "GuideManagerTrajectoriesWithVelocity"),
    ("mpd.models.diffusion_models.sample_functions", "guide_gradient_steps"),
    ("mpd.models.diffusion_models.sample_functions", "ddpm_sample_fn"),
    ("mpd.trainer", "get_dataset"), ("mpd.trainer", "get_model"), ("mpd.utils.loading", "load_params_from_yaml"),
    ("torch_robotics.torch_utils.seed", "fix_random_seed"), ("torch_robotics.torch_utils.torch_timer", "TimerCUDA"),
    ("torch_robotics.torch_utils.torch_utils", "get_torch_device"), ("torch_robotics.torch_utils.torch_utils", "freeze_torch_model_params"),
    ("torch_robotics.trajectory.metrics", "compute_smoothness"), ("torch_robotics.trajectory.metrics", "compute_path_length"),
    ("torch_robotics.trajectory.metrics", "compute_variance_waypoints"),
]
OUT_OF_SCOPE_MODULES = ("torch_robotics.isaac_gym_envs", "torch_robotics.robots", "torch_robotics.trajectory.utils",
                        "torch_robotics.visualizers")


@pytest.fixture()
def aliases():
    import mpd_public_b200.compat as compat
    compat.install()
    yield compat
    compat.uninstall()


def test_reference_import_lines_resolve(aliases):
    for mod, name in IMPORTED:
        m = importlib.import_module(mod)
        assert hasattr(m, name), (mod, name)
    import mpd.models
    assert mpd.models.GaussianDiffusionModel is importlib.import_module("mpd_public_b200").GaussianDiffusionModel  # get_model's lookup


@pytest.mark.skipif(not os.path.exists(REF), reason="reference tree not present on this box")
def test_inference_py_binds_to_our_signatures(aliases):
    import mpd_public_b200 as M
    tree = ast.parse(open(REF).read())
    # 1. the import list above is exactly what the file imports from those packages (minus rendering / simulation)
    found = []
    for node in ast.walk(tree):
        if isinstance(node, ast.ImportFrom) and node.module and node.module.split(".")[0] in ("mpd", "mp_baselines", "torch_robotics", "experiment_launcher"):
            if node.module.startswith(OUT_OF_SCOPE_MODULES):
                continue
            found += [(node.module, a.name) for a in node.names]
    assert sorted(found) == sorted(IMPORTED), sorted(set(found) ^ set(IMPORTED))
    # 2. keyword arguments of the calls on the path (lines 76-326) are accepted by our callables
    targets = {
        "CostCollision": M.CostCollision, "CostGPTrajectory": M.CostGPTrajectory, "CostComposite": M.CostComposite,
        "GuideManagerTrajectoriesWithVelocity": M.GuideManagerTrajectoriesWithVelocity, "TemporalUnet": M.TemporalUnet,
        "get_model": aliases.get_model, "get_dataset": aliases.get_dataset, "guide_gradient_steps": M.guide_gradient_steps,
        "run_inference": M.GaussianDiffusionModel.run_inference, "warmup": M.GaussianDiffusionModel.warmup,
        "get_hard_conditions": M.TrajectoryDataset.get_hard_conditions, "random_coll_free_q": M.PlanningTask.random_coll_free_q,
        "get_trajs_collision_and_free": M.PlanningTask.get_trajs_collision_and_free,
        "unnormalize_trajectories": M.TrajectoryDataset.unnormalize_trajectories,
        "compute_success_free_trajs": M.PlanningTask.compute_success_free_trajs,
        "compute_fraction_free_trajs": M.PlanningTask.compute_fraction_free_trajs,
        "compute_collision_intensity_trajs": M.PlanningTask.compute_collision_intensity_trajs,
        "get_collision_fields": M.PlanningTask.get_collision_fields,
        "get_collision_fields_extra_objects": M.PlanningTask.get_collision_fields_extra_objects,
        "compute_smoothness": M.compute_smoothness, "compute_path_length": M.compute_path_length,
        "compute_variance_waypoints": M.compute_variance_waypoints, "fix_random_seed": aliases.fix_random_seed,
        "get_torch_device": aliases.get_torch_device, "freeze_torch_model_params": aliases.freeze_torch_model_params,
    }
    checked = 0
    for node in ast.walk(tree):
        if not isinstance(node, ast.Call) or not (76 <= node.lineno <= 326):
            continue
        name = node.func.id if isinstance(node.func, ast.Name) else node.func.attr if isinstance(node.func, ast.Attribute) else None
        fn = targets.get(name)
        if fn is None:
            continue
        sig = inspect.signature(fn)
        has_var_kw = any(p.kind == p.VAR_KEYWORD for p in sig.parameters.values())
        for kw in node.keywords:
            if kw.arg is None:  # **expansion
                continue
            assert kw.arg in sig.parameters or has_var_kw, f"inference.py:{node.lineno} {name}(... {kw.arg}=...) is not accepted"
        n_pos = len(node.args) + (1 if isinstance(node.func, ast.Attribute) and name not in ("compute_smoothness",) and inspect.isfunction(fn) and "self" in sig.parameters else 0)
        positional = [p for p in sig.parameters.values() if p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD)]
        assert n_pos <= len(positional) or any(p.kind == p.VAR_POSITIONAL for p in sig.parameters.values()), (node.lineno, name)
        checked += 1
    assert checked >= 20, checked
    # 3. the diffusion keyword arguments inference.py routes through run_inference reach ddpm_sample_fn / the fused loop
    fused = inspect.signature(M.GaussianDiffusionModel._p_sample_loop_fused).parameters
    for kw in ("guide", "n_guide_steps", "t_start_guide", "noise_std_extra_schedule_fn"):
        assert kw in fused and kw in inspect.signature(M.ddpm_sample_fn).parameters


def test_install_refuses_to_shadow_a_real_package(tmp_path, monkeypatch):
    import mpd_public_b200.compat as compat
    compat.uninstall()
    (tmp_path / "mp_baselines").mkdir()
    (tmp_path / "mp_baselines" / "__init__.py").write_text("REAL = True\n")
    monkeypatch.syspath_prepend(str(tmp_path))
    importlib.invalidate_caches()
    with pytest.raises(RuntimeError, match="real 'mp_baselines'"):
        compat.install()
    mods = compat.install(force=True)
    assert "mp_baselines" in mods
    compat.uninstall()
    sys.modules.pop("mp_baselines", None)


@pytest.mark.gpu
@pytest.mark.parametrize("model_id,planner_alg", [("EnvSpheres3D-RobotPanda", "mpd"), ("EnvSpheres3D-RobotPanda", "diffusion_prior_then_guide"),
                                                  ("EnvSimple2D-RobotPointMass", "mpd"), ("EnvDense2D-RobotPointMass", "diffusion_prior")])
def test_inference_script_runs_against_the_aliases(model_id, planner_alg):
    sys.path.insert(0, os.path.join(ROOT, "examples"))
    import inference_like_reference as script
    n = 12
    res = script.experiment(model_id=model_id, planner_alg=planner_alg, n_samples=n, seed=7, verbose=False)
    t_start_guide, n_extra, n_guide = 7, 5, 5
    n_entries = 25 + n_extra + 1 + ((t_start_guide + n_extra) * n_guide if planner_alg == "diffusion_prior_then_guide" else 0)
    D = 14 if "Panda" in model_id else 4
    assert res["trajs_iters_shape"] == (n_entries, n, 64, D)
    final = res["trajs_final_normalized"]
    assert torch.isfinite(final).all()
    for k, v in res["hard_conds"].items():
        assert torch.equal(final[:, k, :], v.expand(n, D))
    assert 0.0 <= res["fraction_free"] <= 1.0 and 0.0 <= res["collision_intensity"] <= 1.0 and res["t_total"] > 0
