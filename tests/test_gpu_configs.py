"""BASELINE.json configs 2, 3 and (one shard of) 5 as parity cases (the bench line is config 4):
  2  EnvDense2D-RobotPointMass, H=64, B=100, guidance on
  3  EnvNarrowPassageDense2D-RobotPointMass, H=64, B=512, collision / smoothness weight sweep
     (w_coll in {1e-2, 3e-2, 1e-1} x w_smooth in {1e-7, 1e-4, 1e-2})
The guide gradient and guide_gradient_steps are compared with the oracle on the same seeded inputs (the oracle looks up
the texels the CUDA grid builder produced); one guided reverse step of the full model is compared per configuration, and
the whole guided loop is checked through size-independent properties (finite, hard conditions exact, chain end == sample).
B = 512 does not fit one wave of clusters, so config 3 also exercises the per-layer tensor-core kernels inside the loop."""
import numpy as np
import pytest
import torch

from mpd_public_b200 import synthetic as S
from oracle import mpd_oracle as O
from tests.golden import cases as C
from tests.test_gpu_parity import TOL_KERNEL, TOL_STEP, cuda_model, oracle_model, rel

pytestmark = pytest.mark.gpu

CONFIGS = {
    "cfg2_dense2d": ("EnvDense2D-RobotPointMass", 100, [(3e-2, 1e-2)]),
    "cfg3_narrow2d": ("EnvNarrowPassageDense2D-RobotPointMass", 512,
                      [(wc, ws) for wc in (1e-2, 3e-2, 1e-1) for ws in (1e-7, 1e-4, 1e-2)]),
}
UCASE = "pm2d_opt0_h64"
H = 64

_built = {}


def build(name, wc, ws):
    import mpd_public_b200 as M
    key = (name, wc, ws)
    if key not in _built:
        model_id, batch, _ = CONFIGS[name]
        if name not in _built:
            prob = S.make_problem_by_id(model_id, n_support_points=H, cell=0.01)
            _built[name] = (prob, M.TrajectoryDataset(prob, "cuda"))
        prob, ds = _built[name]
        robot, task = ds.robot, ds.task
        robot.dt = prob.dt
        costs, weights = [], []
        for f in task.get_collision_fields():
            costs.append(M.CostCollision(robot, H, field=f, sigma_coll=1.0))
            weights.append(wc)
        costs.append(M.CostGPTrajectory(robot, H, prob.dt, sigma_gp=1.0))
        weights.append(ws)
        guide = M.GuideManagerTrajectoriesWithVelocity(ds, M.CostComposite(robot, H, costs, weights_cost_l=weights),
                                                       clip_grad=True, interpolate_trajectories_for_collision=True)
        texels = [f.texels.cpu() for f in task.get_collision_fields() if hasattr(f, "texels")]
        _built[key] = (guide, ds, prob, O.make_guide_spec(prob, wc, ws, texels_list=texels))
    return _built[key]


def near_line_input(prob, batch, seed, out_of_range=False):
    """Normalised trajectories around the straight start -> goal line (costs active), a few beyond [-1, 1] if asked."""
    d = prob.robot.state_dim
    rng = np.random.default_rng([seed, batch, d])
    s = np.concatenate([prob.start, np.zeros(prob.robot.q_dim)])
    g = np.concatenate([prob.goal, np.zeros(prob.robot.q_dim)])
    lam = np.linspace(0, 1, H)[None, :, None]
    x = (1 - lam) * s[None, None, :] + lam * g[None, None, :]
    x = 2 * (x - prob.mins) / (prob.maxs - prob.mins) - 1
    x = x + 0.15 * rng.standard_normal((batch, H, d))
    if out_of_range:
        x[batch // 2, 7, 1] = -1.4  # one element in the whole batch flips LimitsNormalizer's global clip
    else:
        x = np.clip(x, -0.999, 0.999)
    return torch.as_tensor(x.astype(np.float32))


@pytest.mark.parametrize("name", list(CONFIGS))
def test_guide_gradient_over_the_weight_sweep(name):
    model_id, batch, sweep = CONFIGS[name]
    for k, (wc, ws) in enumerate(sweep):
        guide, ds, prob, spec = build(name, wc, ws)
        x = near_line_input(prob, batch, seed=40 + k, out_of_range=(k % 2 == 1))
        ref, parts = O.guide_manager_grad(spec, x, return_parts=True)
        got = guide(x.cuda())
        assert float(ref.abs().max()) > 0
        assert rel(got, ref) < TOL_KERNEL, (wc, ws, rel(got, ref))
        assert float(got[:, 0].abs().max()) == 0 and float(got[:, -1].abs().max()) == 0  # guides.py:202-203


@pytest.mark.parametrize("name", list(CONFIGS))
def test_guide_gradient_steps(name):
    import mpd_public_b200 as M
    model_id, batch, sweep = CONFIGS[name]
    wc, ws = sweep[len(sweep) // 2]
    guide, ds, prob, spec = build(name, wc, ws)
    hard = O.hard_conditions(prob)
    ohc = {k: v[None].repeat(batch, 1) for k, v in hard.items()}
    x = near_line_input(prob, batch, seed=7, out_of_range=True)
    ref = O.OracleDiffusion.guide_gradient_steps(None, x.clone(), ohc, lambda z: O.guide_manager_grad(spec, z), 5)
    got = M.guide_gradient_steps(x.cuda(), hard_conds={k: v.cuda() for k, v in ohc.items()}, guide=guide, n_guide_steps=5)
    assert rel(got, ref) < TOL_KERNEL


@pytest.mark.parametrize("name", list(CONFIGS))
def test_guided_loop_properties_and_one_step_parity(name):
    model_id, batch, sweep = CONFIGS[name]
    wc, ws = sweep[0]
    guide, ds, prob, spec = build(name, wc, ws)
    model = cuda_model(UCASE)
    model.tensor_cores = "auto"
    om = oracle_model(UCASE)
    D = prob.robot.state_dim
    hard = O.hard_conditions(prob)
    hard_cuda = {k: v.cuda() for k, v in hard.items()}
    n_iters = C.T_DIFF + C.N_EXTRA
    gen = torch.Generator().manual_seed(batch)
    noise = torch.randn((n_iters + 1, batch, H, D), generator=gen)
    kw = dict(guide=guide, n_guide_steps=C.N_GUIDE_STEPS, t_start_guide=C.T_START_GUIDE,
              noise_std_extra_schedule_fn=lambda _t: C.NOISE_STD, n_diffusion_steps_without_noise=C.N_EXTRA)
    in_use = model._engine().mega_info(batch)[0]
    assert in_use == (batch <= 104), "config 2 runs the cluster kernel, config 3 (B = 512) the per-layer kernels"
    chain = model.run_inference(None, hard_cuda, n_samples=batch, horizon=H, return_chain=True, noise=noise.cuda(), **kw).cpu()
    final = model.run_inference(None, hard_cuda, n_samples=batch, horizon=H, return_chain=False, noise=noise.cuda(), **kw).cpu()
    assert chain.shape == (n_iters + 1, batch, H, D) and torch.isfinite(chain).all()
    assert torch.equal(final, chain[-1])
    for k, v in hard.items():
        assert torch.equal(chain[:, :, k, :], v.expand(n_iters + 1, batch, D))
    # every step of the chain, teacher-forced, against the oracle at 1e-3 — guided steps with the guide's recorded decisions
    # taken over and audited (oracle/parity.py); the recorded run must reproduce the graph run bit for bit
    from oracle import parity as P
    chain_rec, dec = P.run_recorded(model, guide, hard_cuda, noise.cuda(), batch, H, C.T_DIFF, C.N_EXTRA, C.T_START_GUIDE,
                                    C.N_GUIDE_STEPS, C.NOISE_STD)
    assert torch.equal(chain_rec, chain)
    res = P.check_loop_per_step(om, spec, chain, noise, hard, dec, C.T_DIFF, C.N_EXTRA, C.T_START_GUIDE, C.N_GUIDE_STEPS,
                                C.NOISE_STD, tol=TOL_STEP)
    print(f"[{name} B={batch}] worst per-step rel err {res['worst']:.3e}, t = T-1 {res['t_last']:.3e}, audit {res['audit']}")


def test_config5_shard_shape_panda_h128_b512():
    """One GPU's shard of config 5 (EnvSpheres3D-RobotPanda, H = 128, 512 trajectories): guide gradient against the oracle
    at this horizon (128 interpolation points = the support points themselves), and the guided loop through properties."""
    import mpd_public_b200 as M
    h, batch = 128, 512
    prob = S.make_problem_by_id("EnvSpheres3D-RobotPanda", n_support_points=h, cell=0.04)
    ds = M.TrajectoryDataset(prob, "cuda")
    robot, task = ds.robot, ds.task
    robot.dt = prob.dt
    costs = [M.CostCollision(robot, h, field=f, sigma_coll=1.0) for f in task.get_collision_fields()]
    weights = [1e-2] * len(costs)
    costs.append(M.CostGPTrajectory(robot, h, prob.dt, sigma_gp=1.0))
    weights.append(1e-7)
    guide = M.GuideManagerTrajectoriesWithVelocity(ds, M.CostComposite(robot, h, costs, weights_cost_l=weights), clip_grad=True,
                                                   interpolate_trajectories_for_collision=True)
    texels = [f.texels.cpu() for f in task.get_collision_fields() if hasattr(f, "texels")]
    spec = O.make_guide_spec(prob, 1e-2, 1e-7, texels_list=texels)
    d = prob.robot.state_dim
    rng = np.random.default_rng(5)
    s0 = np.concatenate([prob.start, np.zeros(prob.robot.q_dim)])
    g0 = np.concatenate([prob.goal, np.zeros(prob.robot.q_dim)])
    lam = np.linspace(0, 1, h)[None, :, None]
    x = (1 - lam) * s0[None, None, :] + lam * g0[None, None, :]
    x = 2 * (x - prob.mins) / (prob.maxs - prob.mins) - 1
    x = torch.as_tensor(np.clip(x + 0.15 * rng.standard_normal((64, h, d)), -0.999, 0.999).astype(np.float32))
    ref = O.guide_manager_grad(spec, x)
    assert float(ref.abs().max()) > 0
    assert rel(guide(x.cuda()), ref) < TOL_KERNEL

    model = cuda_model("panda_opt1_h128")
    model.tensor_cores = "auto"
    hard = O.hard_conditions(prob)
    hard_cuda = {k: v.cuda() for k, v in hard.items()}
    n_iters = C.T_DIFF + C.N_EXTRA
    noise = torch.randn((n_iters + 1, batch, h, d), generator=torch.Generator().manual_seed(9))
    kw = dict(guide=guide, n_guide_steps=C.N_GUIDE_STEPS, t_start_guide=C.T_START_GUIDE,
              noise_std_extra_schedule_fn=lambda _t: C.NOISE_STD, n_diffusion_steps_without_noise=C.N_EXTRA)
    chain = model.run_inference(None, hard_cuda, n_samples=batch, horizon=h, return_chain=True, noise=noise.cuda(), **kw)
    assert chain.shape == (n_iters + 1, batch, h, d) and torch.isfinite(chain).all()
    for k, v in hard_cuda.items():
        assert torch.equal(chain[:, :, k, :], v.expand(n_iters + 1, batch, d))
    # a trajectory's result does not depend on its shard: the first 64 trajectories alone give the same samples (the clip
    # flag is batch-global, so equality holds when neither run trips it; both report it through the same flag path)
    # 22-bit split on every step for this comparison: a one-product step rounds activations to fp16, which turns the last-bit
    # differences between the two code paths into differences of the size of its own rounding error (tests/test_gpu_mega.py)
    model.tensor_cores = "force"
    try:
        sub = model.run_inference(None, hard_cuda, n_samples=64, horizon=h, return_chain=False, noise=noise[:, :64].contiguous().cuda(),
                                  **dict(kw, guide=None))
        full = model.run_inference(None, hard_cuda, n_samples=batch, horizon=h, return_chain=False, noise=noise.cuda(), **dict(kw, guide=None))
    finally:
        model.tensor_cores = "auto"
    assert rel(sub, full[:64]) < 5e-5  # B = 64 runs the cluster kernel, B = 512 the per-layer kernels: same products, partial sums
    # combined in a different order (30 free-running steps)


def C_pos_input(prob, batch, q, seed=23):
    rng = np.random.default_rng(seed)
    lam = np.linspace(0, 1, H)[None, :, None]
    x = (1 - lam) * prob.start[None, None, :] + lam * prob.goal[None, None, :]
    x = 2 * (x - prob.mins[:q]) / (prob.maxs[:q] - prob.mins[:q]) - 1
    x = (x + 0.12 * rng.standard_normal((batch, H, q))).astype(np.float32)
    x[0, 11, 1] = -1.3  # trips the normaliser's batch-global clip
    return x


@pytest.mark.parametrize("model_id,batch,wc,ws", [("EnvSimple2D-RobotPointMass", 9, 3e-2, 1e-2), ("EnvSpheres3D-RobotPanda", 5, 1e-2, 1e-4)])
def test_position_only_guide_manager(model_id, batch, wc, ws):
    """GuideManagerTrajectories (reference guides.py:15-146): position-only state + the manager's own velocity
    trajectory, position / velocity gradients clipped separately, velocity updated by every call. Three consecutive calls
    against the oracle's restatement (gradient and velocity after each)."""
    import mpd_public_b200 as M
    prob = S.make_problem_by_id(model_id, n_support_points=H, cell=0.02)
    ds = M.TrajectoryDataset(prob, "cuda", include_velocity=False)
    robot, task = ds.robot, ds.task
    robot.dt = prob.dt
    q = prob.robot.q_dim
    assert ds.state_dim == q
    costs = [M.CostCollision(robot, H, field=f, sigma_coll=1.0) for f in task.get_collision_fields()]
    weights = [wc] * len(costs)
    costs.append(M.CostGPTrajectory(robot, H, prob.dt, sigma_gp=1.0))
    weights.append(ws)
    guide = M.GuideManagerTrajectories(ds, M.CostComposite(robot, H, costs, weights_cost_l=weights), clip_grad=True,
                                       interpolate_trajectories_for_collision=True, start_state_pos=torch.as_tensor(prob.start),
                                       goal_state_pos=torch.as_tensor(prob.goal), num_steps=H - 1, robot=robot, n_samples=batch,
                                       tensor_args=dict(device="cuda", dtype=torch.float32))
    texels = [f.texels.cpu() for f in task.get_collision_fields() if hasattr(f, "texels")]
    spec = O.make_guide_spec(prob, wc, ws, texels_list=texels)
    vel = O.const_vel_trajectory(prob.start, prob.goal, prob.dt, H - 1, q, set_initial_final_vel_to_zero=True)[:, q:]
    vel = vel[None].repeat(batch, 1, 1)
    assert rel(guide.velocity, vel) < 1e-6
    rng = np.random.default_rng(17)
    lam = np.linspace(0, 1, H)[None, :, None]
    x = (1 - lam) * prob.start[None, None, :] + lam * prob.goal[None, None, :]
    x = 2 * (x - prob.mins[:q]) / (prob.maxs[:q] - prob.mins[:q]) - 1
    x = torch.as_tensor((x + 0.12 * rng.standard_normal((batch, H, q))).astype(np.float32))
    x[1, 9, 0] = 1.2  # the batch-global clip of the normaliser looks at the positions
    xc = x.cuda()
    for call in range(3):
        ref, vel = O.guide_manager_pos_grad(spec, x, vel)
        got = guide(xc)
        assert got.shape == (batch, H, q)
        assert float(ref.abs().max()) > 0
        assert rel(got, ref) < TOL_KERNEL, (call, rel(got, ref))
        assert rel(guide.velocity, vel) < TOL_KERNEL, (call, rel(guide.velocity, vel))
        assert float(got[:, 0].abs().max()) == 0 and float(got[:, -1].abs().max()) == 0
        x = x + ref
        xc = xc + got
    # use_velocity_from_finite_difference=True (guides.py:77-79): velocities = central differences of the positions, the costs
    # reach the positions through them as well, one clipped position gradient per cost, the velocity trajectory is left alone
    guide_fd = M.GuideManagerTrajectories(ds, guide.cost, clip_grad=True, interpolate_trajectories_for_collision=True,
                                          use_velocity_from_finite_difference=True, start_state_pos=torch.as_tensor(prob.start),
                                          goal_state_pos=torch.as_tensor(prob.goal), num_steps=H - 1, robot=robot, n_samples=batch,
                                          tensor_args=dict(device="cuda", dtype=torch.float32))
    vel_before = guide_fd.velocity.clone()
    x = torch.as_tensor(C_pos_input(prob, batch, q))
    for call in range(2):
        ref = O.guide_manager_pos_grad_fd(spec, x)
        got = guide_fd(x.cuda())
        assert float(ref.abs().max()) > 0
        assert rel(got, ref) < TOL_KERNEL, (call, rel(got, ref))
        assert float(got[:, 0].abs().max()) == 0 and float(got[:, -1].abs().max()) == 0
        x = x + ref
    assert torch.equal(guide_fd.velocity, vel_before)


def test_position_only_model_runs_the_reverse_loop_with_its_guide():
    """A position-only diffusion model (state_dim = q_dim) sampled with GuideManagerTrajectories: the stateful guide takes the
    step-by-step path (ddpm_sample_fn per step on the CUDA entry points). Properties only: finite, hard conditions exact,
    the guide's velocity trajectory moved."""
    import mpd_public_b200 as M
    prob = S.make_problem_by_id("EnvSimple2D-RobotPointMass", n_support_points=H, cell=0.02)
    ds = M.TrajectoryDataset(prob, "cuda", include_velocity=False)
    robot, task = ds.robot, ds.task
    robot.dt = prob.dt
    q, batch = prob.robot.q_dim, 6
    sd = S.make_unet_state_dict(3, q, 32, S.UNET_DIM_MULTS[0])
    unet = M.TemporalUnet(n_support_points=H, state_dim=q, unet_input_dim=32, dim_mults=S.UNET_DIM_MULTS[0])
    model = M.GaussianDiffusionModel(model=unet, n_diffusion_steps=C.T_DIFF, predict_epsilon=True)
    model.load_state_dict({"model." + k: torch.as_tensor(v) for k, v in sd.items()}, strict=False)
    model = model.to("cuda").eval()
    costs = [M.CostCollision(robot, H, field=f, sigma_coll=1.0) for f in task.get_collision_fields()]
    weights = [3e-2] * len(costs)
    costs.append(M.CostGPTrajectory(robot, H, prob.dt, sigma_gp=1.0))
    weights.append(1e-2)
    guide = M.GuideManagerTrajectories(ds, M.CostComposite(robot, H, costs, weights_cost_l=weights), clip_grad=True,
                                       interpolate_trajectories_for_collision=True, start_state_pos=torch.as_tensor(prob.start),
                                       goal_state_pos=torch.as_tensor(prob.goal), num_steps=H - 1, robot=robot, n_samples=batch,
                                       tensor_args=dict(device="cuda", dtype=torch.float32))
    vel0 = guide.velocity.clone()
    hard = ds.get_hard_conditions(torch.vstack((torch.as_tensor(prob.start), torch.as_tensor(prob.goal))).cuda(), normalize=True)
    assert hard[0].shape == (q,)
    torch.manual_seed(4)
    chain = model.run_inference(None, hard, n_samples=batch, horizon=H, return_chain=True, sample_fn=M.ddpm_sample_fn,
                                guide=guide, n_guide_steps=2, t_start_guide=C.T_START_GUIDE,
                                noise_std_extra_schedule_fn=lambda _t: C.NOISE_STD, n_diffusion_steps_without_noise=2)
    assert chain.shape == (C.T_DIFF + 2 + 1, batch, H, q) and torch.isfinite(chain).all()
    for k, v in hard.items():
        assert torch.equal(chain[:, :, k, :], v.expand(chain.shape[0], batch, q))
    assert not torch.equal(guide.velocity, vel0)



def test_ingested_checkpoint_runs_the_guided_loop_on_the_gpu(tmp_path):
    """SURVEY §8f.3 on the GPU: a model directory with the reference's on-disk layout (args.yaml + checkpoints/
    ema_model_current_state_dict.pth holding the full GaussianDiffusionModel state dict — schedule buffers + `model.*` keys with
    the reference's names, order and shapes, tests/golden/state_dict_keys.json) and a dataset directory (trajs-free.pt) are
    ingested the way inference.py:103-149 does it; the guided loop of the ingested model equals, bit for bit, the loop of a
    model built directly from the same tensors, and the normaliser limits come out of the trajectory file."""
    import json
    import os
    import yaml
    from mpd_public_b200 import ingest
    d = str(tmp_path)
    os.makedirs(os.path.join(d, "model", "checkpoints"))
    os.makedirs(os.path.join(d, "data", "3"))
    yaml.dump(dict(variance_schedule="exponential", n_diffusion_steps=C.T_DIFF, predict_epsilon=True, unet_input_dim=32,
                   unet_dim_mults_option=0, diffusion_model_class="GaussianDiffusionModel", use_ema=True),
              open(os.path.join(d, "model", "args.yaml"), "w"))
    direct = cuda_model(UCASE)
    sd = {k: v.detach().cpu() for k, v in direct.state_dict().items()}
    ref_layout = json.load(open(os.path.join(C.GOLDEN_DIR, "state_dict_keys.json")))[UCASE]
    assert list(ref_layout.keys()) == list(sd.keys())  # the file written below is what the reference's torch.save would hold
    assert all(list(sd[k].shape) == list(shape) for k, shape in ref_layout.items())
    torch.save(sd, os.path.join(d, "model", "checkpoints", "ema_model_current_state_dict.pth"))
    model, args = ingest.load_diffusion_model(os.path.join(d, "model"), state_dim=4, n_support_points=H, device="cuda")
    assert args["n_diffusion_steps"] == C.T_DIFF and not model.training and next(model.parameters()).is_cuda

    trajs = torch.rand((40, H, 4), generator=torch.Generator().manual_seed(5)) * 2 - 1
    torch.save(trajs, os.path.join(d, "data", "3", "trajs-free.pt"))
    nz, h, dim = ingest.load_trajectory_limits(os.path.join(d, "data"), q_dim=2)
    assert (h, dim) == (H, 4)
    lim = nz.normalizers["traj"]
    assert torch.equal(lim.mins, trajs.reshape(-1, 4).min(0).values) and torch.equal(lim.maxs, trajs.reshape(-1, 4).max(0).values)

    wc, ws = CONFIGS["cfg2_dense2d"][2][0]
    guide, ds, prob, _spec = build("cfg2_dense2d", wc, ws)
    hard = ds.get_hard_conditions(torch.vstack((torch.as_tensor(prob.start), torch.as_tensor(prob.goal))).cuda(), normalize=True)
    B = 16
    noise = torch.randn((C.T_DIFF + 5 + 1, B, H, 4), generator=torch.Generator().manual_seed(9)).cuda()
    kw = dict(guide=guide, n_guide_steps=5, t_start_guide=7, noise_std_extra_schedule_fn=lambda _t: 0.5,
              n_diffusion_steps_without_noise=5)
    a = model.run_inference(None, hard, n_samples=B, horizon=H, return_chain=False, noise=noise, **kw)
    b = direct.run_inference(None, hard, n_samples=B, horizon=H, return_chain=False, noise=noise, **kw)
    assert torch.isfinite(a).all() and torch.equal(a, b)


@pytest.mark.parametrize("n_interp", [64, 100, 200, 512])
def test_guide_gradient_for_other_interpolation_densities(n_interp):
    """`num_interpolated_points_for_collision` other than the default 128 (guides.py:160-167): equal to the horizon (ratio 1, one
    tap per support row), a non-integer ratio, more rows than one pass of the row kernel holds (200 > 128), and a density whose
    interpolation adjoint has more than 8 taps per support row (512 / 64: the kernel's on-the-fly fallback instead of its tap
    table). Guide gradient and five gradient steps against the oracle on the config-2 problem."""
    import mpd_public_b200 as M
    name = "cfg2_dense2d"
    model_id, batch, sweep = CONFIGS[name]
    wc, ws = sweep[0]
    guide, ds, prob, spec0 = build(name, wc, ws)
    texels = [f.texels.cpu() for f in ds.task.get_collision_fields() if hasattr(f, "texels")]
    spec = O.make_guide_spec(prob, wc, ws, texels_list=texels, n_interp=n_interp)
    old = guide.num_interpolated_points_for_collision
    try:
        guide.num_interpolated_points_for_collision = n_interp  # re-read on every call, as the reference does
        x = near_line_input(prob, 32, seed=90 + n_interp, out_of_range=True)
        ref = O.guide_manager_grad(spec, x)
        got = guide(x.cuda())
        assert float(ref.abs().max()) > 0
        assert rel(got, ref) < TOL_KERNEL, (n_interp, rel(got, ref))
        hard = O.hard_conditions(prob)
        ohc = {k: v[None].repeat(32, 1) for k, v in hard.items()}
        ref5 = O.OracleDiffusion.guide_gradient_steps(None, x.clone(), ohc, lambda z: O.guide_manager_grad(spec, z), 5)
        got5 = M.guide_gradient_steps(x.cuda(), hard_conds={k: v.cuda() for k, v in ohc.items()}, guide=guide, n_guide_steps=5)
        assert rel(got5, ref5) < TOL_KERNEL, (n_interp, rel(got5, ref5))
    finally:
        guide.num_interpolated_points_for_collision = old
