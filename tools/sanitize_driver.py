"""Small driver for compute-sanitizer (tools/gpu_sanitize.sh): one launch of every hand-written kernel family at a small batch —
the whole-forward cluster kernel (both precisions), the per-layer path (persistent conv5_tc_kernel, cluster-fused rtb_tc_kernel,
final_kernel, exact fp32 conv_kernel), the guide kernel (single evaluation, fused evaluations, position-only), the DDIM loop,
the one-launch normal generator."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import mpd_public_b200 as M

what = sys.argv[1] if len(sys.argv) > 1 else "all"
dev = torch.device("cuda", 0)
model, guide, ds, prob, sd, n_grid, _mk = bench.build_problem("cfg4", dev)
model.use_cuda_graph = False
H, D, B = 64, prob.robot.state_dim, 11
eng = model._engine(H)
x = torch.randn((B, H, D), device=dev)
hard = ds.get_hard_conditions(torch.vstack((torch.as_tensor(prob.start), torch.as_tensor(prob.goal))).to(dev), normalize=True)
hc = {k: v.unsqueeze(0).repeat(B, 1) for k, v in hard.items()}
if what in ("all", "mega"):
    for t in (5, 20):
        eng.unet_forward_uniform(x, t)
if what in ("all", "layers"):
    eng.set_option("mega", 0)
    for t in (5, 20):
        eng.unet_forward_uniform(x, t)
    model.tensor_cores = "off"
    model._engine(H)
    eng.unet_forward_uniform(x, 7)
    model.tensor_cores = "auto"
    model._engine(H)
    eng.set_option("mega", 1)
if what in ("all", "layers_big"):
    # several work items per CTA: the software-pipelined epilogue of the persistent per-layer kernel (scratch alternating with
    # the item parity, no end-of-item barrier), the four-slot parameter table, and final_kernel at one thread per row
    xb = torch.randn((300, H, D), device=dev)
    eng.set_option("mega", 0)
    for t in (5, 20):
        eng.unet_forward_uniform(xb, t)
    model.p_mean_variance(xb, {}, None, torch.full((300,), 5, device=dev, dtype=torch.long))
    eng.set_option("mega", 1)
if what in ("all", "guide"):
    xg = x.clamp(-1, 1) * 0.7
    guide(xg)
    M.guide_gradient_steps(xg, hard_conds=hc, guide=guide, n_guide_steps=3)
    guide.guide_steps(xg, hc, 2, return_chain=True)
if what in ("all", "loop"):
    noise = torch.randn((8, B, H, D), device=dev)
    model2 = model
    model2.run_inference(None, hard, n_samples=B, horizon=H, return_chain=True, guide=guide, n_guide_steps=2, t_start_guide=30,
                         noise_std_extra_schedule_fn=lambda _t: 0.5, n_diffusion_steps_without_noise=1,
                         noise=torch.randn((27, B, H, D), device=dev)) if False else None
    model2.ddim_sample((B, H, D), hc, guide=guide, t_start_guide=7)
    torch.manual_seed(1)
    model2.run_inference(None, hard, n_samples=B, horizon=H, return_chain=False, guide=guide, n_guide_steps=2, t_start_guide=3,
                         noise_std_extra_schedule_fn=lambda _t: 0.5, n_diffusion_steps_without_noise=1)
torch.cuda.synchronize()
print("sanitize driver done:", what)
