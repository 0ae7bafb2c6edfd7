"""Whole-forward cluster kernel (csrc/unet_mega.cu) — the UNet as ONE launch of 8-CTA clusters — against the
per-layer tensor-core kernels (same fp16-split products, partial sums combined in a different order: agreement at the
fp16-split rounding level, bar 2e-5) and against the oracle (same bar). Batch sizes cover full clusters, a ragged last cluster and B < 8."""
import pytest
import torch

from oracle import mpd_oracle as O
from tests.golden import cases as C
from tests.test_gpu_parity import cuda_model, oracle_model, rel

pytestmark = pytest.mark.gpu

TOL_TC = 2e-5  # measured ~1.5e-6


@pytest.mark.parametrize("case,batch", [("panda_opt1_h64", 8), ("panda_opt1_h64", 100), ("panda_opt1_h64", 3),
                                        ("pm2d_opt0_h64", 37), ("pm2d_opt1_h64", 16), ("panda_opt1_h128", 9)])
def test_mega_forward_matches_per_layer_and_oracle(case, batch):
    model = cuda_model(case)
    model.tensor_cores = "force"
    eng = model._engine()
    om = oracle_model(case)
    d, h, opt, seed = C.UNET_CASES[case]
    g = torch.Generator().manual_seed(1000 + batch)
    x = torch.randn((batch, h, d), generator=g)
    model.tensor_cores = "force"
    try:
        in_use, G, n_layers, a_bytes, smem, why = eng.mega_info(batch)
        assert in_use, f"whole-forward kernel not in use for {case}: {why}"
        assert smem <= 227 * 1024 and G in (1, 2, 4, 8)
        for t in (0, 7, 23):
            with torch.no_grad():
                ref = O.unet_forward(om.sd, x, torch.full((batch,), t))
            eng.set_option("mega", 1)
            out_mega = eng.unet_forward_uniform(x.cuda(), t)
            eng.set_option("mega", 0)
            out_layers = eng.unet_forward_uniform(x.cuda(), t)
            torch.cuda.synchronize()
            assert torch.isfinite(out_mega).all()
            print(f"[{case} B={batch} t={t}] mega vs oracle {rel(out_mega, ref):.2e}, per-layer vs oracle "
                  f"{rel(out_layers, ref):.2e}, mega vs per-layer {rel(out_mega, out_layers):.2e}")
            assert rel(out_mega, ref) < TOL_TC, (t, rel(out_mega, ref))
            assert rel(out_layers, ref) < TOL_TC
            assert rel(out_mega, out_layers) < TOL_TC
    finally:
        eng.set_option("mega", 1)
        model.tensor_cores = "auto"


def test_mega_is_batch_invariant():
    """A trajectory's result does not depend on its cluster or its neighbours (bitwise)."""
    model = cuda_model("panda_opt1_h64")
    model.tensor_cores = "force"
    eng = model._engine()
    g = torch.Generator().manual_seed(5)
    x = torch.randn((29, 64, 14), generator=g).cuda()
    model.tensor_cores = "force"
    try:
        a = eng.unet_forward_uniform(x, 3)
        b = eng.unet_forward_uniform(x[11:18].contiguous(), 3)
        assert torch.equal(a[11:18], b)
    finally:
        model.tensor_cores = "auto"


def test_mega_one_wave_policy():
    """Batches that need more clusters than fit in one wave fall back to the per-layer kernels (and say why); option
    mega = 2 forces the cluster kernel, which must still agree."""
    model = cuda_model("panda_opt1_h64")
    model.tensor_cores = "force"
    eng = model._engine()
    g = torch.Generator().manual_seed(9)
    x = torch.randn((256, 64, 14), generator=g).cuda()
    try:
        in_use, G, n_layers, a_bytes, smem, why = eng.mega_info(256)
        assert not in_use and "one wave" in why, (in_use, why)
        ref = eng.unet_forward_uniform(x, 4)
        eng.set_option("mega", 2)
        assert eng.mega_info(256)[0]
        out = eng.unet_forward_uniform(x, 4)
        assert rel(out, ref) < TOL_TC
    finally:
        eng.set_option("mega", 1)
        model.tensor_cores = "auto"
    assert eng.mega_info(100)[0]


def test_mega_repeated_runs_are_bit_identical():
    """Hand-off protocol check (DSMEM stores -> cluster-scope release/acquire -> tensor-core reads): a stale or torn
    operand read would change some output; 300 back-to-back forwards over two batch shapes must be bit-identical."""
    model = cuda_model("panda_opt1_h64")
    model.tensor_cores = "force"
    eng = model._engine()
    try:
        for batch, t in ((100, 5), (37, 17)):
            g = torch.Generator().manual_seed(batch)
            x = torch.randn((batch, 64, 14), generator=g).cuda()
            ref = eng.unet_forward_uniform(x, t)
            acc = torch.zeros((), device="cuda", dtype=torch.int64)
            for _ in range(150):
                out = eng.unet_forward_uniform(x, t)
                acc += (out != ref).sum()
            assert int(acc) == 0
    finally:
        model.tensor_cores = "auto"


def test_fused_projection_and_ddpm_update_match_the_separate_kernel():
    """final_conv.1 + posterior mean + noise + hard conditions inside the cluster kernel's last epilogue (option
    fuse_final) against the separate final_kernel launch: same chain up to the rounding of the 32-term projection. Run with
    the 22-bit split on every step ('force'): a one-product step rounds activations to fp16, which turns a last-bit difference
    of its input into a difference of the size of its own rounding error (~1e-3 * amplification), so alternative code paths
    agree to ~1e-4 there instead of ~1e-6 (checked below with that bound)."""
    model = cuda_model("panda_opt1_h64")
    model.tensor_cores = "force"
    eng = model._engine()
    B, H, D = 21, 64, 14
    n_iters = C.T_DIFF + C.N_EXTRA
    g = torch.Generator().manual_seed(77)
    noise = torch.randn((n_iters + 1, B, H, D), generator=g).cuda()
    hard = {0: torch.linspace(-0.5, 0.5, D).cuda(), H - 1: torch.linspace(0.4, -0.4, D).cuda()}
    kw = dict(n_diffusion_steps_without_noise=C.N_EXTRA, noise_std_extra_schedule_fn=lambda _t: C.NOISE_STD)
    chains = {}
    try:
        for fuse in (1, 0):
            eng.set_option("fuse_final", fuse)
            chains[fuse] = model.run_inference(None, hard, n_samples=B, horizon=H, return_chain=True, noise=noise, **kw)
    finally:
        eng.set_option("fuse_final", 1)
        model.tensor_cores = "auto"
    assert torch.isfinite(chains[1]).all()
    for k, v in hard.items():
        assert torch.equal(chains[1][:, :, k, :], v.expand(n_iters + 1, B, D))
    # free-running chains: the two code paths differ by the summation order of the projection (~1e-7 per step), which the first
    # reverse step amplifies by up to 1095 and the following ones carry along; measured 1.5e-5 .. 2.3e-5 depending on the build
    assert rel(chains[1], chains[0]) < 5e-5, rel(chains[1], chains[0])
    model.tensor_cores = "auto"
    try:
        for fuse in (1, 0):
            eng.set_option("fuse_final", fuse)
            chains[fuse] = model.run_inference(None, hard, n_samples=B, horizon=H, return_chain=True, noise=noise, **kw)
    finally:
        eng.set_option("fuse_final", 1)
    assert rel(chains[1], chains[0]) < 2e-3, rel(chains[1], chains[0])


def test_api_path_replayed_as_one_graph_matches_eager_draws():
    """run_inference() drawing its own noise: the 31 normal_() draws + the loop are replayed as one torch CUDA graph.
    Same seed => same samples as the eager path (graph-safe Philox offsets), the generator advances identically, new
    start/goal values reach the static buffers, and an engine option change forces a re-capture."""
    model = cuda_model("panda_opt1_h64")
    model.tensor_cores = "auto"
    eng = model._engine()
    B, H, D = 17, 64, 14
    kw = dict(n_diffusion_steps_without_noise=C.N_EXTRA, noise_std_extra_schedule_fn=lambda _t: C.NOISE_STD)
    hard_a = {0: torch.linspace(-0.5, 0.5, D).cuda(), H - 1: torch.linspace(0.4, -0.4, D).cuda()}
    hard_b = {0: torch.linspace(0.3, -0.2, D).cuda(), H - 1: torch.linspace(-0.1, 0.6, D).cuda()}

    def run(hard, graphed, chain=False):
        model.graph_rng = graphed
        torch.manual_seed(123)
        out = model.run_inference(None, hard, n_samples=B, horizon=H, return_chain=chain, **kw)
        tail = torch.randn(4, device="cuda")  # where the generator stands after the call
        return out, tail

    try:
        eager, tail_e = run(hard_a, False)
        graphed, tail_g = run(hard_a, True)
        assert model.__dict__.get("graph_rng", True), model.__dict__.get("_graph_rng_error")
        assert torch.equal(eager, graphed)
        assert torch.equal(tail_e, tail_g), "the graph replay must consume the generator exactly like the eager draws"
        again, _ = run(hard_a, True)
        assert torch.equal(again, graphed)
        other_e, _ = run(hard_b, False)
        other_g, _ = run(hard_b, True)      # same graph, new start / goal copied into the static buffers
        assert torch.equal(other_e, other_g) and not torch.equal(other_g, graphed)
        assert torch.equal(other_g[:, 0, :], hard_b[0].expand(B, D))
        chain_e, _ = run(hard_a, False, chain=True)
        chain_g, _ = run(hard_a, True, chain=True)
        assert torch.equal(chain_e, chain_g) and torch.equal(chain_g[-1], graphed)
        eng.set_option("fuse_final", 0)     # bumps the engine generation: the cached graph must not be reused
        refused, _ = run(hard_a, True)
        assert rel(refused, graphed) < 2e-3  # another code path for the projection; one-product steps: see the test above
    finally:
        eng.set_option("fuse_final", 1)
        model.graph_rng = True


@pytest.mark.parametrize("case", ["panda3d", "simple2d"])
def test_guide_evaluations_of_a_step_in_one_launch_are_bit_identical(case):
    """Option fuse_guide: the n_guide_steps evaluations of a step as ONE launch (trajectory kept in shared memory, the
    batch-global clip flag through a grid barrier) against one launch per evaluation: the same chain, bit for bit."""
    from tests.test_gpu_parity import cuda_guide
    model_id, ucase, cell, wc, ws, batch = C.GUIDE_CASES[case]
    model = cuda_model(ucase)
    model.tensor_cores = "auto"
    eng = model._engine()
    guide, ds, prob = cuda_guide(case)
    hard = {k: v.cuda() for k, v in O.hard_conditions(prob).items()}
    H, D = prob.n_support_points, prob.robot.state_dim
    n_iters = C.T_DIFF + C.N_EXTRA
    g = torch.Generator().manual_seed(31)
    noise = (1.3 * torch.randn((n_iters + 1, batch, H, D), generator=g)).cuda()  # some values beyond [-1, 1]: exercises the clip flag
    kw = dict(guide=guide, n_guide_steps=C.N_GUIDE_STEPS, t_start_guide=C.T_START_GUIDE,
              noise_std_extra_schedule_fn=lambda _t: C.NOISE_STD, n_diffusion_steps_without_noise=C.N_EXTRA)
    chains = {}
    try:
        for fuse in (1, 0):
            eng.set_option("fuse_guide", fuse)
            chains[fuse] = model.run_inference(None, hard, n_samples=batch, horizon=H, return_chain=True, noise=noise, **kw)
    finally:
        eng.set_option("fuse_guide", 1)
    assert torch.isfinite(chains[1]).all()
    assert torch.equal(chains[1], chains[0])
