"""Parity on exactly what is benched: the per-step oracle criterion on

  * BASELINE config 4 AS BENCHED — EnvSpheres3D-RobotPanda, H = 64, B = 100, the full 0.01 m grid (201^3 texels),
    dim_mults option 1, the whole-forward cluster kernel, the per-timestep precision policy, CUDA-graph replay, the guide
    evaluations of a step fused in one launch, the 13th cluster ragged (100 = 12 * 8 + 4);
  * one shard of config 5 through the per-layer kernels — H = 128 and B = 128 > 104, so the cluster kernel is off.

Every step of the CUDA chain is re-done by the oracle from the CUDA x_t with the same noise and must match within 1e-3
relative; guided steps are NOT given any allowance for branch flips: the guide records its discrete decisions, the oracle
takes them over and audits them (oracle/parity.py). Also here: the guide paths that must agree bit for bit (fused
evaluations vs one launch per evaluation), the batch-global clamp in its undecidable band, the self-collision cost,
per-field margins / sigma_coll / lattices, the Panda kinematics against known answers, a horizon other than the UNet's
`n_support_points`, and the precision policy itself.
"""
import math

import numpy as np
import pytest
import torch

from mpd_public_b200 import synthetic as S
from oracle import mpd_oracle as O
from oracle import parity as P
from tests.golden import cases as C
from tests.test_gpu_parity import TOL_KERNEL, TOL_STEP, TOL_T_LAST, cuda_model, oracle_model, rel

pytestmark = pytest.mark.gpu

_panda = {}


def panda_setup(h, wc=1e-2, ws=1e-7, **task_kw):
    """dataset / guide / oracle spec of EnvSpheres3D-RobotPanda on the full 0.01 m grid, built the way inference.py:195-236 does"""
    import mpd_public_b200 as M
    key = (h, wc, ws, tuple(sorted(task_kw.items())))
    if key not in _panda:
        prob = S.make_problem_by_id("EnvSpheres3D-RobotPanda", h)
        if ("ds", h) not in _panda:
            _panda[("ds", h)] = M.TrajectoryDataset(prob, "cuda")
        ds = _panda[("ds", h)]
        robot = ds.robot
        robot.dt = prob.dt
        fields = ds.task.get_collision_fields()
        costs = [M.CostCollision(robot, h, field=f, sigma_coll=1.0) for f in fields]
        weights = [wc] * len(costs)
        costs.append(M.CostGPTrajectory(robot, h, prob.dt, sigma_gp=1.0))
        weights.append(ws)
        guide = M.GuideManagerTrajectoriesWithVelocity(ds, M.CostComposite(robot, h, costs, weights_cost_l=weights), clip_grad=True,
                                                       interpolate_trajectories_for_collision=True)
        if ("tex", h) not in _panda:
            _panda[("tex", h)] = [f.texels.cpu() for f in fields if hasattr(f, "texels")]
        spec = O.make_guide_spec(prob, wc, ws, texels_list=_panda[("tex", h)])
        _panda[key] = (prob, ds, guide, spec)
    return _panda[key]


def _loop_parity(ucase, h, batch, expect_mega, seed):
    prob, ds, guide, spec = panda_setup(h)
    model = cuda_model(ucase)
    model.tensor_cores = "auto"
    model.use_cuda_graph = True
    om = oracle_model(ucase)
    D = prob.robot.state_dim
    eng = model._engine(h)
    in_use, G, n_layers, a_bytes, smem, why = eng.mega_info(batch)
    assert in_use == expect_mega, (in_use, why)
    hard = O.hard_conditions(prob)
    hard_cuda = {k: v.cuda() for k, v in hard.items()}
    n_iters = C.T_DIFF + C.N_EXTRA
    noise = torch.randn((n_iters + 1, batch, h, D), generator=torch.Generator().manual_seed(seed))
    kw = dict(guide=guide, n_guide_steps=C.N_GUIDE_STEPS, t_start_guide=C.T_START_GUIDE,
              noise_std_extra_schedule_fn=lambda _t: C.NOISE_STD, n_diffusion_steps_without_noise=C.N_EXTRA)
    chain = model.run_inference(None, hard_cuda, n_samples=batch, horizon=h, return_chain=True, noise=noise.cuda(), **kw).cpu()
    assert chain.shape == (n_iters + 1, batch, h, D) and torch.isfinite(chain).all()
    # the throughput entry the bench times returns the same plans
    assert torch.equal(model.sample(hard_cuda, batch, horizon=h, noise=noise.cuda(), **kw).cpu(), chain[-1])
    chain_rec, dec = P.run_recorded(model, guide, hard_cuda, noise.cuda(), batch, h, C.T_DIFF, C.N_EXTRA, C.T_START_GUIDE,
                                    C.N_GUIDE_STEPS, C.NOISE_STD)
    assert torch.equal(chain_rec, chain), "graph replay with fused guide launches vs recorded direct launches"
    assert dec.shape[2] == 3, "objects grid + workspace boundary + self-collision"
    res = P.check_loop_per_step(om, spec, chain, noise, hard, dec, C.T_DIFF, C.N_EXTRA, C.T_START_GUIDE, C.N_GUIDE_STEPS,
                                C.NOISE_STD, tol=TOL_STEP, tol_t_last=TOL_T_LAST)
    print(f"[{ucase} H={h} B={batch} mega={in_use}] worst per-step rel err (t < T-1) {res['worst']:.3e}, t = T-1 {res['t_last']:.3e}, "
          f"decision audit {res['audit']}")
    return res


def test_config4_as_benched_per_step_parity():
    res = _loop_parity("panda_opt1_h64", 64, 100, expect_mega=True, seed=41)
    assert res["audit"]["n"] > 0


def test_config5_shard_per_layer_kernels_per_step_parity():
    _loop_parity("panda_opt1_h128", 128, 128, expect_mega=False, seed=43)


def test_precision_policy_eps_error_and_step_selection():
    """The per-timestep precision policy (engine.cu step_prec): which steps issue one fp16 product, what that costs in eps,
    and that tensor_cores = 'force' keeps the 22-bit split everywhere. The eps error times the step's amplification
    c1[t] * sqrt(1/abar_t - 1) is what reaches the posterior mean: it must stay under 4e-4 (the size of the reference's own
    fp32 rounding at t = T-1, well inside the 1e-3 per-step bar that test_benched_config_per_step_parity enforces)."""
    model = cuda_model("panda_opt1_h64")
    om = oracle_model("panda_opt1_h64")
    eng = model._engine(64)
    x = torch.randn((100, 64, 14), generator=torch.Generator().manual_seed(3))
    sched = O.make_schedule(C.T_DIFF)
    amp = (sched["posterior_mean_coef1"] * sched["sqrt_recipm1_alphas_cumprod"]).numpy()
    try:
        for t in (0, 3, 7, 12, 15, 18, 19, 20, 24):
            with torch.no_grad():
                ref = O.unet_forward(om.sd, x, torch.full((100,), t))
            model.tensor_cores = "auto"
            model._engine(64)
            e_auto = rel(eng.unet_forward_uniform(x.cuda(), t), ref)
            model.tensor_cores = "force"
            model._engine(64)
            e_force = rel(eng.unet_forward_uniform(x.cuda(), t), ref)
            one_product = amp[t] <= 0.21
            print(f"t={t:2d} amplification {amp[t]:.3g}: eps rel err auto {e_auto:.2e} (one product: {one_product}), force {e_force:.2e}, "
                  f"-> mean error ~{e_auto * amp[t]:.1e}")
            assert e_force < 2e-5
            if one_product:
                assert e_auto * amp[t] < 4e-4, (t, e_auto, amp[t])
                assert e_auto > 2e-5, "the one-product step should differ measurably from the 22-bit split"
            else:
                assert e_auto < 2e-5
    finally:
        model.tensor_cores = "auto"
        model._engine(64)


def test_guide_fused_evaluations_equal_one_launch_per_evaluation():
    """guide_gradient_steps inside the loop: the n evaluations of a step in ONE launch (trajectory resident in shared memory,
    clip flag resolved per CTA) vs one launch per evaluation — bit-identical chains, including a batch that trips the clamp."""
    prob, ds, guide, spec = panda_setup(64)
    model = cuda_model("panda_opt1_h64")
    model.tensor_cores = "auto"
    eng = model._engine(64)
    B, H, D = 37, 64, 14
    hard_cuda = {k: v.cuda() for k, v in O.hard_conditions(prob).items()}
    n_iters = C.T_DIFF + C.N_EXTRA
    noise = torch.randn((n_iters + 1, B, H, D), generator=torch.Generator().manual_seed(12)).cuda()
    kw = dict(guide=guide, n_guide_steps=C.N_GUIDE_STEPS, t_start_guide=C.T_START_GUIDE, n_diffusion_steps_without_noise=C.N_EXTRA,
              noise_std_extra_schedule_fn=lambda _t: 2.5)  # large extra noise: iterates leave [-1, 1], the clamp branch is live
    out = {}
    try:
        for fuse in (1, 0):
            eng.set_option("fuse_guide", fuse)
            guide.batch_dependent_clamps(reset=True)
            out[fuse] = model.run_inference(None, hard_cuda, n_samples=B, horizon=H, return_chain=True, noise=noise, **kw)
        assert torch.isfinite(out[1]).all()
        assert torch.equal(out[0], out[1])
    finally:
        eng.set_option("fuse_guide", 1)


def test_batch_global_clamp_in_the_undecidable_band():
    """LimitsNormalizer.unnormalize clamps the whole batch iff any element leaves [-1 - 1e-4, 1 + 1e-4]
    (normalization.py:160-162). A trajectory with an element in (1, 1 + 1e-4] and none beyond cannot decide alone; in the
    fused launch it waits for the grid. guide_gradient_steps (5 evaluations) on inputs built to sit in that band — with and
    without another trajectory tripping the flag — against the oracle, which evaluates the branch on the whole batch."""
    import mpd_public_b200 as M
    case = "simple2d"
    from tests.test_gpu_parity import cuda_guide, oracle_guide_spec
    guide, ds, prob = cuda_guide(case)
    spec = oracle_guide_spec(case, ds)
    hard = O.hard_conditions(prob)
    B = 6
    ohc = {k: v[None].repeat(B, 1) for k, v in hard.items()}
    hc = {k: v.cuda() for k, v in ohc.items()}
    base = torch.as_tensor(C.guide_input(case))
    base = torch.cat([base, base[:2] * 0.9])[:B].clone()
    model = cuda_model("pm2d_opt0_h64")
    eng = model._engine(64)
    for trip in (False, True):
        x = base.clone()
        x[1, 20, 2] = 1.00005          # velocity column, no cost pulls it back quickly: stays in the band
        x[2, 30, 3] = -1.00007
        if trip:
            x[4, 10, 2] = 1.4          # far outside: the batch's flag is set, the band elements above get clamped to +-1
        ref = O.OracleDiffusion.guide_gradient_steps(None, x.clone(), ohc, lambda z: O.guide_manager_grad(spec, z), 5)
        got = M.guide_gradient_steps(x.cuda(), hard_conds=hc, guide=guide, n_guide_steps=5)   # five evaluations in ONE launch
        assert rel(got, ref) < TOL_KERNEL, (trip, rel(got, ref))
        got_seq, _ = guide.guide_steps(x.cuda(), hc, 5, return_chain=True)                   # one launch per evaluation
        assert torch.equal(got, got_seq), "fused launch (flag resolved per CTA / grid wait) vs one launch per evaluation"
    guide.batch_dependent_clamps(reset=True)
    x = base.clone()
    x[1, 20, 2] = 1.00005
    x[4, 10, 2] = 1.4
    M.guide_gradient_steps(x.cuda(), hard_conds=hc, guide=guide, n_guide_steps=1)
    assert guide.batch_dependent_clamps(reset=True) >= 1, "trajectory 1's clamp was decided by trajectory 4"


def test_self_collision_cost_and_per_field_parameters():
    """The robot self-collision field (a18 of the coverage table), a collision cost with sigma_coll != 1, a
    per-field cutoff margin and two grid fields on DIFFERENT lattices, against the oracle."""
    import mpd_public_b200 as M
    prob = S.make_problem_by_id("EnvSpheres3D-RobotPanda", 64, cell=0.04)
    ds = M.TrajectoryDataset(prob, "cuda")
    robot, H = ds.robot, 64
    robot.dt = prob.dt
    env = prob.env
    f_obj = ds.task.get_collision_fields()[0]
    # a second object field on a coarser lattice over a smaller box, with its own margin
    lim2 = np.array([[-0.8, -0.8, -0.8], [0.8, 0.8, 0.8]])
    shape2 = tuple(int(round((lim2[1][k] - lim2[0][k]) / 0.1)) + 1 for k in range(3))
    extra = np.array([[0.3, 0.2, 0.5, 0.25], [-0.35, -0.3, 0.6, 0.2]])
    f_extra = M.GridSDFField.from_primitives(lim2, 0.1, shape2, extra, np.zeros((0, 6)), "cuda", cutoff_margin=0.12)
    f_border = M.WorkspaceBoundaryField(env.limits)
    f_self = M.SelfCollisionField(robot, cutoff_margin=0.35)   # wide margin: the pairwise hinges are active on the test input
    costs = [M.CostCollision(robot, H, field=f_obj, sigma_coll=1.0), M.CostCollision(robot, H, field=f_extra, sigma_coll=0.5),
             M.CostCollision(robot, H, field=f_border, sigma_coll=1.0), M.CostCollision(robot, H, field=f_self, sigma_coll=2.0),
             M.CostGPTrajectory(robot, H, prob.dt, sigma_gp=1.0)]
    weights = [1e-2, 2e-2, 1e-2, 3e-2, 1e-7]
    guide = M.GuideManagerTrajectoriesWithVelocity(ds, M.CostComposite(robot, H, costs, weights_cost_l=weights), clip_grad=True,
                                                   interpolate_trajectories_for_collision=True)
    # oracle: the same four collision costs, one by one (per-cost margin / sigma / weight), through the reference manager logic
    g_obj = O.GridSDF(env.limits, env.cell, f_obj.texels.cpu(), env.grid_shape)
    g_extra = O.GridSDF(lim2, 0.1, f_extra.texels.cpu(), shape2)
    x = torch.as_tensor(C.guide_input("panda3d"))
    mins, maxs = torch.as_tensor(prob.mins), torch.as_tensor(prob.maxs)
    pairs = O.default_self_pairs(prob.robot)
    assert pairs == f_self.pairs
    xn = x.clone()
    with torch.enable_grad():
        xn.requires_grad_(True)
        xu = O.limits_unnormalize(xn, mins, maxs)
        xi = O.interpolate_points(xu, 128)
        cen = O.sphere_centers(prob.robot, xi[..., :7])
        r = prob.robot.sphere_radius
        cost_l = [O.collision_cost(g_obj, cen, r, 0.05, 1.0), O.collision_cost(g_extra, cen, r, 0.12, 0.5),
                  O.collision_cost(lambda p: O.border_sdf(p, env.limits), cen, r, 0.05, 1.0),
                  O.self_collision_cost(cen, r, pairs, 0.35, 2.0)[0], O.gp_cost(xu, 7, prob.dt, 1.0)]
        grad, parts = 0, []
        for c, w in zip(cost_l, weights):
            gc = torch.autograd.grad([c.sum()], [xu], retain_graph=True)[0]
            gc = O.clip_grad_by_norm(gc, 1.0)
            gc[..., 0, :] = 0.0
            gc[..., -1, :] = 0.0
            parts.append(gc)
            grad = grad + w * gc
        ref = -1.0 * grad
    assert all(float(p.abs().max()) > 0 for p in parts), [float(p.abs().max()) for p in parts]
    got = guide(x.cuda())
    assert rel(got, ref) < TOL_KERNEL, rel(got, ref)
    # the configuration is re-read on every call, as in the reference: changing a weight takes effect
    guide.cost.weights_cost_l[3] = 0.0
    got0 = guide(x.cuda())
    ref0 = ref + weights[3] * parts[3]
    assert rel(got0, ref0) < TOL_KERNEL


def test_panda_forward_kinematics_known_answers():
    """SURVEY App. E constants through the CUDA chain (mpdb_debug_fk) against answers that do not come from this repository's
    tables: the flange of the zero configuration sits at (0.088, 0, 0.926) (0.333 + 0.316 + 0.384 - 0.107 up, 0.0825 - 0.0825
    + 0.088 out: the public Franka Emika Panda home pose), link 1..7 origins of q = 0 follow from the same sums, and rotating
    joint 1 by 90 degrees rotates every point about the z axis."""
    import ctypes as Ct
    from mpd_public_b200 import _lib
    prob, ds, guide, spec = panda_setup(64)
    handle = guide._handle(torch.device("cuda"), 64)
    q = torch.zeros((3, 7))
    q[1, 0] = math.pi / 2
    q[2] = torch.tensor([0.3, -0.5, 0.2, -1.9, 0.4, 1.7, -0.6])
    cen = torch.empty((3, 8, 3), device="cuda")
    _lib.check(_lib.lib().mpdb_debug_fk(handle, _lib.fptr(q.cuda()), _lib.fptr(cen), 3, _lib.stream_ptr(torch.device("cuda"))))
    cen = cen.cpu()
    zero = torch.tensor([[0, 0, 0.333], [0, 0, 0.333], [0, 0, 0.649], [0.0825, 0, 0.649], [0, 0, 1.033], [0, 0, 1.033],
                         [0.088, 0, 1.033], [0.088, 0, 0.926]])
    assert float((cen[0] - zero).abs().max()) < 1e-6, cen[0]
    rotz = torch.tensor([[0.0, -1.0, 0.0], [1.0, 0.0, 0.0], [0.0, 0.0, 1.0]])
    assert float((cen[1] - zero @ rotz.T).abs().max()) < 1e-6
    # a generic configuration: the oracle's own chain (its constants are written out in oracle/mpd_oracle.py, not imported)
    ref = O.sphere_centers(prob.robot, q[2].double()[None])[0]
    assert float((cen[2].double() - ref).abs().max()) < 2e-6
    # link lengths are invariants of the chain whatever q is
    d13 = float((cen[2, 2] - cen[2, 0]).norm())
    d35 = float((cen[2, 4] - cen[2, 3]).norm())
    assert abs(d13 - 0.316) < 1e-6 and abs(d35 - math.hypot(0.0825, 0.384)) < 1e-6


def test_horizon_other_than_n_support_points():
    """The reference's UNet is fully convolutional: `conditional_sample(horizon=...)` works for any horizon divisible by 8
    (temporal_unet.py:118-171). Here every horizon gets its own device plan; a request must never run on a plan built for
    another horizon (that wrote past the caller's buffers)."""
    model = cuda_model("pm2d_opt0_h64")
    om = oracle_model("pm2d_opt0_h64")
    B, H2, D = 5, 32, 4
    x = torch.randn((B, H2, D), generator=torch.Generator().manual_seed(8))
    t = torch.tensor([0, 5, 11, 17, 24])
    with torch.no_grad():
        ref = O.unet_forward(om.sd, x, t)
    assert rel(model.model(x.cuda(), t.cuda(), None), ref) < TOL_KERNEL
    hard = {0: torch.linspace(-0.3, 0.3, D).cuda(), H2 - 1: torch.linspace(0.2, -0.2, D).cuda()}
    n_iters = C.T_DIFF
    noise = torch.randn((n_iters + 1, B, H2, D), generator=torch.Generator().manual_seed(9))
    chain = model.run_inference(None, hard, n_samples=B, horizon=H2, return_chain=True, noise=noise.cuda()).cpu()
    assert chain.shape == (n_iters + 1, B, H2, D) and torch.isfinite(chain).all()
    res = P.check_loop_per_step(om, None, chain, noise, {k: v.cpu() for k, v in hard.items()}, None, C.T_DIFF, 0, float("inf"), 1, 1.0)
    assert res["worst"] < TOL_STEP
    # the engine of one horizon refuses tensors of another instead of reading past them
    with pytest.raises(RuntimeError):
        model._engine(64).unet_forward(x.cuda(), t.cuda())
    with pytest.raises(RuntimeError):
        model._engine(64).sample_loop(noise.cuda(), hard, None, 0, float("inf"), 0, False, [1.0] * n_iters, False, False)


def test_prior_then_guide_post_loop_chain():
    """run_prior_then_guidance (inference.py:263-282): N x guide_gradient_steps(n_guide_steps=1) after an unguided loop, every
    iterate kept. One call with a chain output vs the reference's loop of single calls (same kernels: bit-identical)."""
    import mpd_public_b200 as M
    from tests.test_gpu_parity import cuda_guide
    guide, ds, prob = cuda_guide("panda3d")
    batch = 2
    hard = O.hard_conditions(prob)
    hc = {k: v.cuda()[None].repeat(batch, 1) for k, v in hard.items()}
    x = torch.as_tensor(C.guide_input("panda3d")).cuda()
    n = (C.T_START_GUIDE + C.N_EXTRA) * C.N_GUIDE_STEPS
    trajs, post = x, []
    for _ in range(n):
        trajs = M.guide_gradient_steps(trajs, hard_conds=hc, guide=guide, n_guide_steps=1, unnormalize_data=False)
        post.append(trajs)
    ref_chain = torch.stack(post, dim=0)
    out, chain = guide.guide_steps(x, hc, n, return_chain=True)
    assert chain.shape == ref_chain.shape and torch.equal(chain, ref_chain) and torch.equal(out, ref_chain[-1])


@pytest.mark.parametrize("shape", [(7,), (100, 64, 14), (512, 128, 14), (3, 1000, 1001)])
def test_device_normal_reproduces_torch_cuda_generator(shape):
    """csrc/rng.cu: all the loop's draws in ONE launch, bit-identical to consecutive `torch.randn(shape, device='cuda')` /
    `randn_like` calls (the reference's generator consumption, diffusion_model_base.py:165 + sample_functions.py:51), and the
    torch generator ends up exactly where those calls would have left it."""
    from mpd_public_b200.diffusion_model import _DeviceNormal
    dev = torch.device("cuda", 0)
    assert _DeviceNormal.supported(dev)
    n_draws = 5 if np.prod(shape) > 1e6 else 31
    for seed in (0, 123456789, 2 ** 63 + 11):
        torch.manual_seed(seed)
        torch.randn(33, device=dev)                       # a draw before: the offset does not start at zero
        ref = torch.stack([torch.randn(shape, device=dev) for _ in range(n_draws)])
        tail_ref = torch.randn(5, device=dev)
        torch.manual_seed(seed)
        torch.randn(33, device=dev)
        rng = _DeviceNormal(dev)
        buf = torch.empty((n_draws, *shape), device=dev)
        rng.advance(buf[0].numel(), n_draws)
        rng.fill(buf)
        tail = torch.randn(5, device=dev)
        assert torch.equal(buf, ref), (shape, seed, float((buf - ref).abs().max()))
        assert torch.equal(tail, tail_ref), "the generator must advance exactly as the eager draws advance it"


def test_run_inference_same_seed_with_and_without_device_rng():
    """Identical seeds => identical samples whether the noise comes from torch's own normal_() launches or from the one-launch
    generator, eager or graph-replayed (north-star: 'outputs match the reference on identical seeds')."""
    model = cuda_model("pm2d_opt0_h64")
    model.tensor_cores = "auto"
    B, H, D = 9, 64, 4
    hard = {0: torch.linspace(-0.5, 0.5, D).cuda(), H - 1: torch.linspace(0.4, -0.4, D).cuda()}
    kw = dict(n_diffusion_steps_without_noise=C.N_EXTRA, noise_std_extra_schedule_fn=lambda _t: C.NOISE_STD)
    outs = {}
    try:
        for dev_rng in (False, True):
            for graphed in (False, True):
                model.device_rng, model.graph_rng = dev_rng, graphed
                torch.manual_seed(2024)
                outs[(dev_rng, graphed)] = (model.run_inference(None, hard, n_samples=B, horizon=H, return_chain=True, **kw),
                                            torch.randn(3, device="cuda"))
    finally:
        model.device_rng, model.graph_rng = True, True
    base = outs[(False, False)]
    for key, (chain, tail) in outs.items():
        assert torch.equal(chain, base[0]), key
        assert torch.equal(tail, base[1]), key
