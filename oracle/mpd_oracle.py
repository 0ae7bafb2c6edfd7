"""ORACLE — test infrastructure, not product code.

CPU restatement (plain PyTorch, fp32 by default, fp64 on request) of the guided
`p_sample_loop` hot path of jacarvalho/mpd-public. Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s CPU-baseline / `--impl reference` legs may import this file; the product
(`mpd_public_b200/`) never does and fails loudly when its CUDA library is missing.

Pinning status
--------------
* **Reference-pinned** (checked against outputs of the reference's own code, generated in the build
  container by `tests/golden/make_golden.py` under the name-only shim of `oracle/ref_shim.py`):
  `unet_forward` (reference temporal_unet.py:118-171, layers.py:229-355), `make_schedule`
  (diffusion_model_base.py:66-104, helpers.py:40-46), `p_mean_variance` (:143-155),
  `ddpm_step` (sample_functions.py:18-62), `guide_gradient_steps` (:65-83), `p_sample_loop`
  (diffusion_model_base.py:158-182), `ddim_sample` (:184-259), `guide_manager_grad` (guides.py:173-236: per-cost autograd,
  clip-by-norm with +1e-6, endpoint zeroing, weighting, negation), `limits_unnormalize`
  (normalization.py:156-167, including the batch-global clip branch).
* **Parity unpinned** (sources absent from /root/reference: `mp_baselines@8a50c3c`,
  `torch_robotics@d704c78`, `storm@54543cf`, see .SUBMODULES.json): `interpolate_points`,
  `panda_sphere_centers`, `GridSDF`, `collision_cost_*`, `gp_cost`, `const_vel_trajectory`. `guide_manager_pos_grad`
  (the position-only manager's own lines, guides.py:60-118) IS pinned, with the GP prior as its single cost: the reference
  manager differentiates every cost once without `retain_graph` / `allow_unused`, so it cannot run a composite of several
  costs itself; `ddim_sample`, `scale_grad_by_std` and `predict_epsilon=False` are pinned too. These follow the frozen spec of
  SURVEY.md Appendix C/E; each choice is a named switch below. The reference holds no tests or
  golden vectors for any of this path.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch
import torch.nn.functional as F

# ------------------------------------------------------------------------------------------------
# Switches for the unpinned arithmetic
# ------------------------------------------------------------------------------------------------
SWITCH_C3_COLLISION_ON_INTERPOLATED = True   # C3: collision costs see the interpolated trajectory
SWITCH_C5_NEAREST_CELL = True                # C5: nearest-cell lookup, gradient from the stored texel
SWITCH_C4_SELF_SUM_OVER_PAIRS = True         # C4: self-collision cost = sum over listed sphere pairs of the hinge (not a min)
SWITCH_FD_CENTRAL = True                     # robot.get_velocity(x_pos) of a position-only trajectory = central difference, zero end rows

# Panda kinematic chain — SURVEY Appendix E (public Franka URDF values), written out HERE and not imported from the product
# (`mpd_public_b200.synthetic` holds its own copy; tests/test_host_cpu.py compares the two and checks both against the known
# flange pose of the zero configuration), so a wrong joint origin cannot pass on both sides unnoticed.
PANDA_JOINT_XYZ = (
    (0.0, 0.0, 0.333),       # J1
    (0.0, 0.0, 0.0),         # J2
    (0.0, -0.316, 0.0),      # J3
    (0.0825, 0.0, 0.0),      # J4
    (-0.0825, 0.384, 0.0),   # J5
    (0.0, 0.0, 0.0),         # J6
    (0.088, 0.0, 0.0),       # J7
)
PANDA_JOINT_ROLL = (0.0, -math.pi / 2, math.pi / 2, math.pi / 2, -math.pi / 2, math.pi / 2, math.pi / 2)
PANDA_FLANGE_XYZ = (0.0, 0.0, 0.107)


# ------------------------------------------------------------------------------------------------
# TemporalUnet  (reference temporal_unet.py:118-171)
# ------------------------------------------------------------------------------------------------
def group_norm_n_groups(n_channels, target_n_groups=8):
    """reference layers.py:389-395"""
    if n_channels < target_n_groups:
        return 1
    for n_groups in range(target_n_groups, target_n_groups + 10):
        if n_channels % n_groups == 0:
            return n_groups
    return 1


def sinusoidal_pos_emb(t, dim=32):
    """reference layers.py:243-255 (always evaluated in fp32 there; here in t's float dtype)."""
    half = dim // 2
    e = math.log(10000) / (half - 1)
    freq = torch.exp(torch.arange(half, device=t.device) * -e)
    emb = t[:, None] * freq[None, :]
    return torch.cat((emb.sin(), emb.cos()), dim=-1)


def time_mlp(sd, t, dtype):
    """reference layers.py:229-240: Sinusoidal(32) -> Linear(32,128) -> Mish -> Linear(128,32)."""
    emb = sinusoidal_pos_emb(t).to(dtype)
    h = F.mish(F.linear(emb, sd["time_mlp.encoder.1.weight"], sd["time_mlp.encoder.1.bias"]))
    return F.linear(h, sd["time_mlp.encoder.3.weight"], sd["time_mlp.encoder.3.bias"])


def conv1d_block(sd, p, x):
    """reference layers.py:276-293: Conv1d(k, pad k//2) -> GroupNorm -> Mish."""
    w = sd[p + ".block.0.weight"]
    y = F.conv1d(x, w, sd[p + ".block.0.bias"], padding=w.shape[-1] // 2)
    y = F.group_norm(y, group_norm_n_groups(w.shape[0]), sd[p + ".block.2.weight"], sd[p + ".block.2.bias"], eps=1e-5)
    return F.mish(y)


def residual_temporal_block(sd, p, x, c_emb, capture=None):
    """reference layers.py:343-355."""
    cond = F.linear(F.mish(c_emb), sd[p + ".cond_mlp.1.weight"], sd[p + ".cond_mlp.1.bias"])
    h = conv1d_block(sd, p + ".blocks.0", x) + cond[:, :, None]
    if capture is not None:
        capture[p + ".blocks.0"] = h
    h = conv1d_block(sd, p + ".blocks.1", h)
    if (p + ".residual_conv.weight") in sd:
        res = F.conv1d(x, sd[p + ".residual_conv.weight"], sd[p + ".residual_conv.bias"])
    else:
        res = x
    out = h + res
    if capture is not None:
        capture[p] = out
    return out


def unet_forward(sd, x, t, n_levels=None, capture=None):
    """ε = TemporalUnet(x [B,H,D], t [B]) with conditioning_type=None, self_attention=False.

    `sd`: TemporalUnet state-dict (no 'model.' prefix) of torch tensors.
    """
    dtype = x.dtype
    if n_levels is None:
        n_levels = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("downs."))
    c_emb = time_mlp(sd, t, dtype)
    x = x.transpose(1, 2)  # b h c -> b c h  (temporal_unet.py:138)
    skips = []
    for i in range(n_levels):
        x = residual_temporal_block(sd, f"downs.{i}.0", x, c_emb, capture)
        x = residual_temporal_block(sd, f"downs.{i}.1", x, c_emb, capture)
        skips.append(x)
        if i < n_levels - 1:
            x = F.conv1d(x, sd[f"downs.{i}.4.conv.weight"], sd[f"downs.{i}.4.conv.bias"], stride=2, padding=1)
            if capture is not None:
                capture[f"downs.{i}.4"] = x
    x = residual_temporal_block(sd, "mid_block1", x, c_emb, capture)
    x = residual_temporal_block(sd, "mid_block2", x, c_emb, capture)
    for i in range(n_levels - 1):
        x = torch.cat((x, skips.pop()), dim=1)
        x = residual_temporal_block(sd, f"ups.{i}.0", x, c_emb, capture)
        x = residual_temporal_block(sd, f"ups.{i}.1", x, c_emb, capture)
        x = F.conv_transpose1d(x, sd[f"ups.{i}.4.conv.weight"], sd[f"ups.{i}.4.conv.bias"], stride=2, padding=1)
        if capture is not None:
            capture[f"ups.{i}.4"] = x
    x = conv1d_block(sd, "final_conv.0", x)
    if capture is not None:
        capture["final_conv.0"] = x
    x = F.conv1d(x, sd["final_conv.1.weight"], sd["final_conv.1.bias"])
    return x.transpose(1, 2)


# ------------------------------------------------------------------------------------------------
# Schedule (reference diffusion_model_base.py:66-104, helpers.py:26-46)
# ------------------------------------------------------------------------------------------------
def make_schedule(n_diffusion_steps, variance_schedule="exponential"):
    T = n_diffusion_steps
    if variance_schedule == "exponential":
        x = torch.linspace(0, T, T)
        beta_start, beta_end = torch.tensor(1e-4), torch.tensor(1.0)
        a = 1 / T * torch.log(beta_end / beta_start)
        betas = beta_start * torch.exp(a * x)
    elif variance_schedule == "cosine":
        steps = T + 1
        xs = np.linspace(0, steps, steps)
        ac = np.cos(((xs / steps) + 0.008) / (1 + 0.008) * np.pi * 0.5) ** 2
        ac = ac / ac[0]
        betas = torch.tensor(np.clip(1 - (ac[1:] / ac[:-1]), a_min=0, a_max=0.999), dtype=torch.float32)
    else:
        raise NotImplementedError
    alphas = 1.0 - betas
    ac = torch.cumprod(alphas, dim=0)
    ac_prev = torch.cat([torch.ones(1), ac[:-1]])
    pv = betas * (1.0 - ac_prev) / (1.0 - ac)
    return {
        "betas": betas,
        "alphas_cumprod": ac,
        "alphas_cumprod_prev": ac_prev,
        "sqrt_alphas_cumprod": torch.sqrt(ac),
        "sqrt_one_minus_alphas_cumprod": torch.sqrt(1.0 - ac),
        "log_one_minus_alphas_cumprod": torch.log(1.0 - ac),
        "sqrt_recip_alphas_cumprod": torch.sqrt(1.0 / ac),
        "sqrt_recipm1_alphas_cumprod": torch.sqrt(1.0 / ac - 1),
        "posterior_variance": pv,
        "posterior_log_variance_clipped": torch.log(torch.clamp(pv, min=1e-20)),
        "posterior_mean_coef1": betas * np.sqrt(ac_prev) / (1.0 - ac),
        "posterior_mean_coef2": (1.0 - ac_prev) * np.sqrt(alphas) / (1.0 - ac),
    }


# ------------------------------------------------------------------------------------------------
# Normaliser (reference normalization.py:149-167)
# ------------------------------------------------------------------------------------------------
def limits_normalize(x, mins, maxs):
    x = (x - mins) / (maxs - mins)
    return 2 * x - 1


def limits_unnormalize(x, mins, maxs, eps=1e-4):
    if x.max() > 1 + eps or x.min() < -1 - eps:   # batch-global branch
        x = torch.clip(x, -1, 1)
    x = (x + 1) / 2.0
    return x * (maxs - mins) + mins


# ------------------------------------------------------------------------------------------------
# Restated (unpinned) planning arithmetic — SURVEY Appendix C / E
# ------------------------------------------------------------------------------------------------
def interpolate_points(x, n):
    """C1: linear resampling along the horizon, align_corners=True, all D channels."""
    return F.interpolate(x.transpose(-2, -1), size=n, mode="linear", align_corners=True).transpose(-2, -1)


def _rx(roll, dtype):
    c, s = math.cos(roll), math.sin(roll)
    return torch.tensor([[1, 0, 0], [0, c, -s], [0, s, c]], dtype=dtype)


def panda_frames(q, joint_xyz, joint_roll, flange_xyz):
    """Origins [..., 8, 3] and rotations [..., 8, 3, 3] of link 1..7 + flange (App. E)."""
    dtype = q.dtype
    lead = q.shape[:-1]
    R = torch.eye(3, dtype=dtype).expand(*lead, 3, 3)
    o = torch.zeros(*lead, 3, dtype=dtype)
    origins, rots = [], []
    for i in range(7):
        o = o + (R @ torch.as_tensor(joint_xyz[i], dtype=dtype))
        c, s = torch.cos(q[..., i]), torch.sin(q[..., i])
        z, one = torch.zeros_like(c), torch.ones_like(c)
        Rz = torch.stack([torch.stack([c, -s, z], -1), torch.stack([s, c, z], -1), torch.stack([z, z, one], -1)], -2)
        R = R @ _rx(float(joint_roll[i]), dtype) @ Rz
        origins.append(o); rots.append(R)
    o = o + (R @ torch.as_tensor(flange_xyz, dtype=dtype))
    origins.append(o); rots.append(R)
    return torch.stack(origins, -2), torch.stack(rots, -3)


def sphere_centers(robot, q):
    """C2/E1: world positions of the robot's collision spheres, [..., S, ws_dim]."""
    if robot.kind == "pointmass":
        return q[..., None, :]
    o, R = panda_frames(q, PANDA_JOINT_XYZ, PANDA_JOINT_ROLL, PANDA_FLANGE_XYZ)
    f = torch.as_tensor(robot.sphere_frame.astype(np.int64) - 1)
    off = torch.as_tensor(robot.sphere_offset, dtype=q.dtype)
    return o[..., f, :] + torch.einsum("...sij,sj->...si", R[..., f, :, :], off)


def sdf_and_grad_analytic(p, spheres, boxes):
    """C5: min over primitives of the signed distance and its closed-form gradient at p[N, dim]."""
    dim = p.shape[-1]
    best = torch.full(p.shape[:-1], float("inf"), dtype=p.dtype)
    grad = torch.zeros_like(p)
    for s in np.asarray(spheres, dtype=np.float64).reshape(-1, dim + 1):
        c = torch.as_tensor(s[:dim], dtype=p.dtype)
        d = p - c
        n = torch.sqrt((d * d).sum(-1))
        val = n - float(s[dim])
        g = d / torch.clamp(n, min=1e-12)[..., None]
        take = val < best
        best = torch.where(take, val, best)
        grad = torch.where(take[..., None], g, grad)
    for b in np.asarray(boxes, dtype=np.float64).reshape(-1, 2 * dim):
        c = torch.as_tensor(b[:dim], dtype=p.dtype)
        h = torch.as_tensor(b[dim:], dtype=p.dtype)
        d = p - c
        qv = d.abs() - h
        qpos = torch.clamp(qv, min=0)
        outside = torch.sqrt((qpos * qpos).sum(-1))
        qmax, amax = qv.max(dim=-1)
        val = outside + torch.clamp(qmax, max=0)
        sgn = torch.where(d >= 0, torch.ones_like(d), -torch.ones_like(d))
        g_out = sgn * qpos / torch.clamp(outside, min=1e-12)[..., None]
        g_in = sgn * F.one_hot(amax, dim).to(p.dtype)
        g = torch.where((outside > 0)[..., None], g_out, g_in)
        take = val < best
        best = torch.where(take, val, best)
        grad = torch.where(take[..., None], g, grad)
    return best, grad


class GridSDF:
    """C5: voxel grid over the env limits; each node stores {sdf, d/dx, d/dy[, d/dz]}.

    `texels`: float tensor [prod(shape), 1+dim] (x slowest ... last axis fastest).
    """

    def __init__(self, limits, cell, texels, shape):
        self.lo = torch.as_tensor(np.asarray(limits)[0], dtype=torch.float32)
        self.cell = float(cell)
        self.shape = tuple(int(s) for s in shape)
        self.texels = texels
        self.dim = len(self.shape)

    @classmethod
    def build(cls, limits, cell, shape, spheres, boxes, dtype=torch.float32):
        axes = [torch.as_tensor(limits[0][d], dtype=dtype) + torch.arange(shape[d], dtype=dtype) * torch.as_tensor(cell, dtype=dtype)
                for d in range(len(shape))]
        pts = torch.stack(torch.meshgrid(*axes, indexing="ij"), -1).reshape(-1, len(shape))
        vals, grads = [], []
        for chunk in torch.split(pts, 1 << 20):
            v, g = sdf_and_grad_analytic(chunk, spheres, boxes)
            vals.append(v); grads.append(g)
        tex = torch.cat([torch.cat(vals)[:, None], torch.cat(grads)], dim=1).to(torch.float32)
        return cls(limits, cell, tex, shape)

    def cell_coords(self, p):
        """(p - lo) / cell, the quantity whose rounding picks the node"""
        return (p.detach() - self.lo.to(p.dtype)) / torch.as_tensor(self.cell, dtype=p.dtype)

    def flat_index(self, p):
        idx = torch.round(self.cell_coords(p)).long()
        flat = torch.zeros(p.shape[:-1], dtype=torch.long)
        for d in range(self.dim):
            flat = flat * self.shape[d] + idx[..., d].clamp(0, self.shape[d] - 1)
        return flat

    def __call__(self, p, flat=None):
        """sdf(p) with value from the nearest node and gradient = stored texel (surrogate). `flat`: node indices decided
        elsewhere (parity instrumentation: the CUDA kernel's own choice), else the nearest node."""
        tex = self.texels.to(p.dtype)
        if flat is None:
            flat = self.flat_index(p)
        t = tex[flat]
        return t[..., 0] + ((p - p.detach()) * t[..., 1:]).sum(-1)


def border_sdf(p, limits):
    """C4: distance to the nearest workspace wall, positive inside."""
    lo = torch.as_tensor(np.asarray(limits)[0], dtype=p.dtype)
    hi = torch.as_tensor(np.asarray(limits)[1], dtype=p.dtype)
    return torch.minimum(p - lo, hi - p).min(dim=-1).values


def border_sdf_forced(p, limits, axis, low):
    """border_sdf with the wall decided elsewhere: axis [...], low [...] bool (low wall = p - lo, else hi - p)"""
    lo = torch.as_tensor(np.asarray(limits)[0], dtype=p.dtype)
    hi = torch.as_tensor(np.asarray(limits)[1], dtype=p.dtype)
    pa = torch.gather(p, -1, axis[..., None])[..., 0]
    return torch.where(low, pa - lo[axis], hi[axis] - pa)


def collision_cost(sdf_fn, centers, radii, cutoff_margin, sigma_coll=1.0, active=None):
    """C4: sum over horizon rows and link spheres of relu(r + margin - sdf(p)) / sigma_coll^2. `active` (bool, same shape as
    the hinge argument): hinge branches decided elsewhere (parity instrumentation) instead of by the sign."""
    d = sdf_fn(centers)
    r = torch.as_tensor(radii, dtype=centers.dtype)
    v = r + cutoff_margin - d
    h = torch.relu(v) if active is None else torch.where(active, v, torch.zeros_like(v))
    return h.sum(dim=(-1, -2)) / (sigma_coll ** 2)


def self_collision_cost(centers, radii, pairs, margin, sigma_coll=1.0, active=None):
    """C4 (restated, unpinned): sum over rows and listed sphere pairs (a, b) of relu(margin - (|c_a - c_b| - r_a - r_b)).
    `active`: [..., rows, n_pairs] bool, hinge branches decided elsewhere. Returns (cost, hinge argument [..., rows, n_pairs])."""
    r = torch.as_tensor(radii, dtype=centers.dtype)
    a = torch.as_tensor([p[0] for p in pairs], dtype=torch.long)
    b = torch.as_tensor([p[1] for p in pairs], dtype=torch.long)
    diff = centers[..., a, :] - centers[..., b, :]
    n = torch.sqrt((diff * diff).sum(-1))
    v = margin - (n - r[a] - r[b])
    h = torch.relu(v) if active is None else torch.where(active, v, torch.zeros_like(v))
    return h.sum(dim=(-1, -2)) / (sigma_coll ** 2), v


def gp_cost(x, q_dim, dt, sigma_gp=1.0):
    """C6: constant-velocity GP prior, sum_h e_h^T Q^-1 e_h, e_h = x_{h+1} - Phi x_h."""
    p, v = x[..., :q_dim], x[..., q_dim:2 * q_dim]
    ep = p[..., 1:, :] - p[..., :-1, :] - dt * v[..., :-1, :]
    ev = v[..., 1:, :] - v[..., :-1, :]
    s = 1.0 / (sigma_gp ** 2)
    a, b, c = 12.0 / dt ** 3 * s, -6.0 / dt ** 2 * s, 4.0 / dt * s
    return (a * ep * ep + 2 * b * ep * ev + c * ev * ev).sum(dim=(-1, -2))


@dataclass
class GuideSpec:
    """Everything `GuideManagerTrajectoriesWithVelocity` + `CostComposite` hold (inference.py:195-236)."""
    robot: object                 # synthetic.RobotSpec
    mins: torch.Tensor            # [D]
    maxs: torch.Tensor
    grid_fields: list             # list[GridSDF]
    border_limits: object         # [2, dim] or None
    cutoff_margin: float
    dt: float
    weight_collision: float
    weight_smoothness: float
    sigma_gp: float = 1.0
    clip_grad: bool = True
    max_grad_norm: float = 1.0
    n_interp: int = 128
    interpolate: bool = True
    self_pairs: object = None     # list of (a, b) sphere pairs of the self-collision field, or None
    self_margin: float = 0.05
    sigma_coll: float = 1.0


def n_collision_costs(spec: GuideSpec):
    return len(spec.grid_fields) + (spec.border_limits is not None) + (spec.self_pairs is not None and spec.robot.kind == "panda")


def decode_decisions(spec: GuideSpec, dec):
    """int32 [B, n_costs, NI, S] as recorded by the CUDA guide (include/mpdb200.h, mpdb_guide_record_decisions) ->
    list per cost of dicts: grid {flat, active}; border {axis, low, active}; self {active [B, NI, n_pairs]}."""
    dec = torch.as_tensor(dec).long()
    out, k = [], 0
    for _ in spec.grid_fields:
        out.append({"flat": dec[:, k] >> 1, "active": (dec[:, k] & 1).bool()})
        k += 1
    if spec.border_limits is not None:
        out.append({"axis": dec[:, k] >> 2, "low": ((dec[:, k] >> 1) & 1).bool(), "active": (dec[:, k] & 1).bool()})
        k += 1
    if spec.self_pairs is not None and spec.robot.kind == "panda":
        a = torch.as_tensor([p[0] for p in spec.self_pairs], dtype=torch.long)
        b = torch.as_tensor([p[1] for p in spec.self_pairs], dtype=torch.long)
        m = dec[:, k]                                                   # [B, NI, S] partner masks
        act_ab = ((m[..., a] >> b) & 1).bool()                          # [B, NI, n_pairs]: a sees b active
        act_ba = ((m[..., b] >> a) & 1).bool()
        out.append({"active": act_ab, "active_sym": act_ba})
        k += 1
    return out


def composite_costs(spec: GuideSpec, x, x_interp, decisions=None, report=None):
    """C3: CostComposite(trajs, x_interpolated=..., return_invidual_costs_and_weights=True).
    `decisions` (decode_decisions output): discrete choices taken from the CUDA kernel instead of made here.
    `report` (list): receives per cost the quantities whose sign / rounding makes those choices."""
    q = spec.robot.q_dim
    xs = x_interp if SWITCH_C3_COLLISION_ON_INTERPOLATED else x
    centers = sphere_centers(spec.robot, xs[..., :q])
    r = torch.as_tensor(spec.robot.sphere_radius, dtype=centers.dtype)
    costs, weights = [], []
    k = 0
    for g in spec.grid_fields:
        d = decisions[k] if decisions is not None else None
        fn = (lambda p, g=g, d=d: g(p, flat=d["flat"])) if d is not None else g
        costs.append(collision_cost(fn, centers, spec.robot.sphere_radius, spec.cutoff_margin, spec.sigma_coll,
                                    d["active"] if d is not None else None))
        if report is not None:
            with torch.no_grad():
                cc = g.cell_coords(centers)
                report.append({"kind": "grid", "flat": g.flat_index(centers), "round_dist": 0.5 - (cc - torch.round(cc)).abs().amax(-1),
                               "hinge": r + spec.cutoff_margin - g(centers), "cell_coords": cc})
        weights.append(spec.weight_collision)
        k += 1
    if spec.border_limits is not None:
        d = decisions[k] if decisions is not None else None
        fn = (lambda p, d=d: border_sdf_forced(p, spec.border_limits, d["axis"], d["low"])) if d is not None else \
            (lambda p: border_sdf(p, spec.border_limits))
        costs.append(collision_cost(fn, centers, spec.robot.sphere_radius, spec.cutoff_margin, spec.sigma_coll,
                                    d["active"] if d is not None else None))
        if report is not None:
            with torch.no_grad():
                lo = torch.as_tensor(np.asarray(spec.border_limits)[0], dtype=centers.dtype)
                hi = torch.as_tensor(np.asarray(spec.border_limits)[1], dtype=centers.dtype)
                walls = torch.cat((centers - lo, hi - centers), dim=-1)      # [.., 2 * dim]: low walls then high walls
                report.append({"kind": "border", "walls": walls, "hinge": r + spec.cutoff_margin - border_sdf(centers, spec.border_limits)})
        weights.append(spec.weight_collision)
        k += 1
    if spec.self_pairs is not None and spec.robot.kind == "panda":
        d = decisions[k] if decisions is not None else None
        c, v = self_collision_cost(centers, spec.robot.sphere_radius, spec.self_pairs, spec.self_margin, spec.sigma_coll,
                                   d["active"] if d is not None else None)
        costs.append(c)
        if report is not None:
            report.append({"kind": "self", "hinge": v.detach()})
        weights.append(spec.weight_collision)
        k += 1
    costs.append(gp_cost(x, q, spec.dt, spec.sigma_gp))
    weights.append(spec.weight_smoothness)
    return costs, weights


def clip_grad_by_norm(grad, max_grad_norm):
    """reference guides.py:224-230."""
    n = torch.linalg.norm(grad + 1e-6, dim=-1, keepdims=True)
    return torch.clip(n, 0.0, max_grad_norm) / n * grad


def guide_manager_grad(spec: GuideSpec, x_normalized, return_parts=False, decisions=None, report=None):
    """reference guides.py:173-211 — returns -sum_c w_c * zero_ends(clip(d cost_c / d x_unnormalized)).
    `decisions`: int32 [B, n_costs, NI, S] recorded by the CUDA guide for this very evaluation — its texel / wall / hinge
    choices are taken over, so that the comparison is exact away from AND at the discontinuities of the cost (the choices
    themselves are audited separately, `audit_decisions`)."""
    x = x_normalized.clone()
    with torch.enable_grad():
        x.requires_grad_(True)
        x = limits_unnormalize(x, spec.mins.to(x.dtype), spec.maxs.to(x.dtype))
        x_interp = interpolate_points(x, spec.n_interp) if spec.interpolate else x
        cost_l, w_l = composite_costs(spec, x, x_interp, decode_decisions(spec, decisions) if decisions is not None else None,
                                      report)
        grad = 0
        parts = []
        for cost, w in zip(cost_l, w_l):
            g = torch.autograd.grad([cost.sum()], [x], retain_graph=True)[0]
            if spec.clip_grad:
                g = clip_grad_by_norm(g, spec.max_grad_norm)
            g[..., 0, :] = 0.0
            g[..., -1, :] = 0.0
            parts.append(g)
            grad = grad + w * g
    grad = -1.0 * grad
    return (grad, parts) if return_parts else grad


def audit_decisions(spec: GuideSpec, x_normalized, decisions, pos_tol=1e-5):
    """Are the CUDA guide's discrete decisions for this evaluation the oracle's own, except where the oracle's deciding
    quantity sits within `pos_tol` (metres: how far the two sides' sphere centres may differ, given how far their inputs
    differ) of the decision boundary? Returns a dict of counts; `unexplained` must be 0:
      * a texel index may differ only where some coordinate of (p - lo) / cell is within pos_tol / cell of a rounding boundary,
      * a wall may differ only where the two walls' distances agree within 2 pos_tol,
      * a hinge branch (grid, border, self pair) may differ only where |hinge argument| <= 2 pos_tol (sdf is 1-Lipschitz)
        (or, for a grid field, where the texel itself differs for the reason above)."""
    hinge_tol = 2.0 * pos_tol
    rep = []
    guide_manager_grad(spec, x_normalized, report=rep)
    dec = decode_decisions(spec, decisions)
    out = {"n": 0, "index_diff": 0, "hinge_diff": 0, "wall_diff": 0, "unexplained": 0}
    grid_cells = [g.cell for g in spec.grid_fields]
    for k, (d, rp) in enumerate(zip(dec, rep)):
        k_cell = grid_cells[k] if k < len(grid_cells) else 1.0
        if rp["kind"] == "grid":
            idx_diff = d["flat"] != rp["flat"]
            near = rp["round_dist"] <= pos_tol / k_cell
            out["index_diff"] += int(idx_diff.sum())
            out["unexplained"] += int((idx_diff & ~near).sum())
            own = rp["hinge"] > 0
            hd = (d["active"] != own) & ~idx_diff
            out["hinge_diff"] += int(hd.sum())
            out["unexplained"] += int((hd & (rp["hinge"].abs() > hinge_tol)).sum())
            out["n"] += d["flat"].numel()
        elif rp["kind"] == "border":
            walls = rp["walls"]
            dim = walls.shape[-1] // 2
            own = walls.argmin(-1)                                            # first minimum, low walls first
            theirs = torch.where(d["low"], d["axis"], d["axis"] + dim)
            wd = own != theirs
            tie = (torch.gather(walls, -1, theirs[..., None])[..., 0] - walls.amin(-1)).abs() <= hinge_tol
            out["wall_diff"] += int(wd.sum())
            out["unexplained"] += int((wd & ~tie).sum())
            hd = d["active"] != (rp["hinge"] > 0)
            out["hinge_diff"] += int(hd.sum())
            out["unexplained"] += int((hd & (rp["hinge"].abs() > hinge_tol)).sum())
            out["n"] += own.numel()
        else:
            own = rp["hinge"] > 0
            for act in (d["active"], d["active_sym"]):
                hd = act != own
                out["hinge_diff"] += int(hd.sum())
                out["unexplained"] += int((hd & (rp["hinge"].abs() > hinge_tol)).sum())
            out["n"] += own.numel()
    return out


def finite_difference_velocity(x_pos, dt):
    """robot.get_velocity of a position-only trajectory (reference guides.py:78; torch_robotics source absent -> restated,
    PARITY UNPINNED, switch FD_CENTRAL): central difference, zero at both ends."""
    v = torch.zeros_like(x_pos)
    v[..., 1:-1, :] = (x_pos[..., 2:, :] - x_pos[..., :-2, :]) / (2 * dt)
    return v


def const_vel_trajectory(start_pos, goal_pos, dt, num_steps, q_dim, set_initial_final_vel_to_zero=False, dtype=torch.float32):
    """`MultiMPPrior.const_vel_trajectory` (mp_baselines@8a50c3c; source absent, call site reference guides.py:46-53) —
    PARITY UNPINNED. num_steps + 1 states: positions on the straight line, velocity (goal - start) / (num_steps * dt)."""
    start = torch.as_tensor(start_pos, dtype=dtype)[:q_dim]
    goal = torch.as_tensor(goal_pos, dtype=dtype)[:q_dim]
    traj = torch.zeros(num_steps + 1, 2 * q_dim, dtype=dtype)
    mean_vel = (goal - start) / (num_steps * dt)
    for i in range(num_steps + 1):
        traj[i, :q_dim] = start * (num_steps - i) * 1.0 / num_steps + goal * i * 1.0 / num_steps
    traj[:, q_dim:] = mean_vel[None]
    if set_initial_final_vel_to_zero:
        traj[0, q_dim:] = 0.0
        traj[-1, q_dim:] = 0.0
    return traj


def guide_manager_pos_grad_fd(spec: GuideSpec, x_pos_normalized):
    """reference guides.py:60-118 with use_velocity_from_finite_difference=True: the velocity half of the state is
    robot.get_velocity(x_pos) (restated: finite_difference_velocity), one gradient w.r.t. the positions per cost (the costs
    reach them through the velocities as well), clipped, end rows zeroed, weighted; returns -grad_pos."""
    q = spec.robot.q_dim
    x_pos = x_pos_normalized.clone()
    with torch.enable_grad():
        x_pos.requires_grad_(True)
        x_pos = limits_unnormalize(x_pos, spec.mins[:q].to(x_pos.dtype), spec.maxs[:q].to(x_pos.dtype))
        x_interp = interpolate_points(x_pos, spec.n_interp) if spec.interpolate else x_pos
        x_pos_vel = torch.cat((x_pos, finite_difference_velocity(x_pos, spec.dt)), dim=-1)
        cost_l, w_l = composite_costs(spec, x_pos_vel, x_interp)
        grad = 0
        for cost, w in zip(cost_l, w_l):
            g = torch.autograd.grad([cost.sum()], [x_pos], retain_graph=True)[0]
            if spec.clip_grad:
                g = clip_grad_by_norm(g, spec.max_grad_norm)
            g[..., 0, :] = 0.0
            g[..., -1, :] = 0.0
            grad = grad + w * g
    return -1.0 * grad


def guide_manager_pos_grad(spec: GuideSpec, x_pos_normalized, velocity):
    """reference guides.py:60-118 (GuideManagerTrajectories.forward, use_velocity_from_finite_difference=False): the state
    is [unnormalised positions | the manager's velocity trajectory]; per cost the position and the velocity gradient are
    clipped separately, end rows zeroed, weighted; returns (-grad_pos, velocity - grad_velocity)."""
    q = spec.robot.q_dim
    x_pos = x_pos_normalized.clone()
    vel = velocity.clone()
    with torch.enable_grad():
        x_pos.requires_grad_(True)
        vel.requires_grad_(True)
        x_pos = limits_unnormalize(x_pos, spec.mins[:q].to(x_pos.dtype), spec.maxs[:q].to(x_pos.dtype))
        x_interp = interpolate_points(x_pos, spec.n_interp) if spec.interpolate else x_pos
        x_pos_vel = torch.cat((x_pos, vel), dim=-1)
        cost_l, w_l = composite_costs(spec, x_pos_vel, x_interp)
        grad, grad_velocity = 0, 0
        for cost, w in zip(cost_l, w_l):
            g, gv = torch.autograd.grad([cost.sum()], [x_pos, vel], retain_graph=True, allow_unused=True)
            gv = torch.zeros_like(vel) if gv is None else gv
            if spec.clip_grad:
                g = clip_grad_by_norm(g, spec.max_grad_norm)
                gv = clip_grad_by_norm(gv, spec.max_grad_norm)
            g[..., 0, :] = 0.0
            g[..., -1, :] = 0.0
            gv[..., 0, :] = 0.0
            gv[..., -1, :] = 0.0
            grad = grad + w * g
            grad_velocity = grad_velocity + w * gv
    return -1.0 * grad, (velocity - grad_velocity).detach()


# ------------------------------------------------------------------------------------------------
# Sampler (reference diffusion_model_base.py:143-182, sample_functions.py:5-83)
# ------------------------------------------------------------------------------------------------
def apply_hard_conditioning(x, conditions):
    for t, val in conditions.items():
        x[:, t, :] = val.clone()
    return x


def _extract(a, t, ndim):
    return a.gather(-1, t).reshape(t.shape[0], *((1,) * (ndim - 1)))


class OracleDiffusion:
    """GaussianDiffusionModel restated for sampling (predict_epsilon / clip_denoised as configured)."""

    def __init__(self, unet_sd, n_diffusion_steps=25, variance_schedule="exponential", predict_epsilon=True,
                 clip_denoised=True, dtype=torch.float32):
        self.dtype = dtype
        self.sd = {k: torch.as_tensor(v).to(dtype) for k, v in unet_sd.items()}
        self.n_diffusion_steps = n_diffusion_steps
        self.predict_epsilon = predict_epsilon
        self.clip_denoised = clip_denoised
        self.buf = {k: v.to(dtype) for k, v in make_schedule(n_diffusion_steps, variance_schedule).items()}
        self.state_dim = self.sd["final_conv.1.weight"].shape[0]

    def model(self, x, t):
        return unet_forward(self.sd, x, t)

    def p_mean_variance(self, x, t):
        b = self.buf
        noise = self.model(x, t)
        if self.predict_epsilon:
            x_recon = (_extract(b["sqrt_recip_alphas_cumprod"], t, x.ndim) * x
                       - _extract(b["sqrt_recipm1_alphas_cumprod"], t, x.ndim) * noise)
        else:
            x_recon = noise
        if self.clip_denoised:
            x_recon = x_recon.clamp(-1.0, 1.0)
        mean = (_extract(b["posterior_mean_coef1"], t, x.ndim) * x_recon
                + _extract(b["posterior_mean_coef2"], t, x.ndim) * x)
        return mean, _extract(b["posterior_variance"], t, x.ndim), _extract(b["posterior_log_variance_clipped"], t, x.ndim)

    def guide_gradient_steps(self, x, hard_conds, guide, n_guide_steps=1, scale_grad_by_std=False, model_var=None):
        for _ in range(n_guide_steps):
            g = guide(x)
            if scale_grad_by_std:
                g = model_var * g
            x = x + g
            x = apply_hard_conditioning(x, hard_conds)
        return x

    def ddpm_step(self, x, hard_conds, t, noise, guide=None, n_guide_steps=1, scale_grad_by_std=False,
                  t_start_guide=float("inf"), noise_std=1.0, return_mean=False):
        """One `ddpm_sample_fn` call with the step noise injected (`noise` replaces randn_like)."""
        t_single = int(t[0])
        if t_single < 0:
            t = torch.zeros_like(t)
        mean, _, _ = self.p_mean_variance(x, t)
        logvar = _extract(self.buf["posterior_log_variance_clipped"], t, x.ndim)
        std, var = torch.exp(0.5 * logvar), torch.exp(logvar)
        x = mean
        if guide is not None and t_single < t_start_guide:
            x = self.guide_gradient_steps(x, hard_conds, guide, n_guide_steps, scale_grad_by_std, var)
        noise = noise.clone()
        noise[t == 0] = 0
        out = x + std * noise * noise_std
        return (out, mean) if return_mean else out

    def p_sample_loop(self, shape, hard_conds, noise=None, generator=None, return_chain=False,
                      n_diffusion_steps_without_noise=0, guide=None, n_guide_steps=1, scale_grad_by_std=False,
                      t_start_guide=float("inf"), noise_std_fn=None):
        """`noise`: [n_steps+1, B, H, D] injected (row 0 = initial x); else drawn in reference order
        from the global torch generator (randn(shape), then randn_like per step)."""
        steps = list(reversed(range(-n_diffusion_steps_without_noise, self.n_diffusion_steps)))
        draw = (lambda k: noise[k].to(self.dtype).clone()) if noise is not None else \
               (lambda k: torch.randn(shape, generator=generator).to(self.dtype))
        x = draw(0)
        x = apply_hard_conditioning(x, hard_conds)
        chain = [x] if return_chain else None
        for k, i in enumerate(steps):
            t = torch.full((shape[0],), i, dtype=torch.long)
            ns = 1.0 if noise_std_fn is None else float(noise_std_fn(t[0]))
            x = self.ddpm_step(x, hard_conds, t, draw(k + 1), guide, n_guide_steps, scale_grad_by_std, t_start_guide, ns)
            x = apply_hard_conditioning(x, hard_conds)
            if return_chain:
                chain.append(x)
        if return_chain:
            return x, torch.stack(chain, dim=1)
        return x


    def ddim_sample(self, shape, hard_conds, noise=None, generator=None, return_chain=False, guide=None,
                    t_start_guide=float("inf"), n_guide_steps=1, **sample_kwargs):
        """`ddim_sample` (diffusion_model_base.py:184-259): T // 5 steps on a linspace grid of timesteps, eta = 0 (so sigma
        = 0 and the per-step draw is consumed but multiplied away), no clamp of x_start, guide steps before the noise.
        As in the reference, `n_guide_steps` is a named argument that is NOT forwarded: `guide_gradient_steps` receives
        only `**sample_kwargs` (:240-245) and so runs its default single step unless the caller passes it there.
        `noise`: [T//5 + 1, B, H, D] injected (row 0 = initial x); else drawn in reference order."""
        b = self.buf
        total, sampling, eta = self.n_diffusion_steps, self.n_diffusion_steps // 5, 0.0
        times = torch.linspace(0, total - 1, steps=sampling + 1)
        times = torch.cat((torch.tensor([-1.0]), times))
        times = list(reversed(times.int().tolist()))
        draw = (lambda k: noise[k].to(self.dtype).clone()) if noise is not None else \
               (lambda k: torch.randn(shape, generator=generator).to(self.dtype))
        x = apply_hard_conditioning(draw(0), hard_conds)
        chain = [x] if return_chain else None
        for k, (time, time_next) in enumerate(zip(times[:-1], times[1:])):
            t = torch.full((shape[0],), time, dtype=torch.long)
            model_out = self.model(x, t)
            # predict_start_from_noise (:121-132) / predict_noise_from_start: with epsilon prediction the model output IS the noise
            if self.predict_epsilon:
                x_start = (_extract(b["sqrt_recip_alphas_cumprod"], t, x.ndim) * x
                           - _extract(b["sqrt_recipm1_alphas_cumprod"], t, x.ndim) * model_out)
                pred_noise = model_out
            else:
                x_start = model_out
                pred_noise = ((_extract(b["sqrt_recip_alphas_cumprod"], t, x.ndim) * x - model_out)
                              / _extract(b["sqrt_recipm1_alphas_cumprod"], t, x.ndim))
            if time_next < 0:
                x = apply_hard_conditioning(x_start, hard_conds)
                if return_chain:
                    chain.append(x)
                break
            t_next = torch.full((shape[0],), time_next, dtype=torch.long)
            alpha = _extract(b["alphas_cumprod"], t, x.ndim)
            alpha_next = _extract(b["alphas_cumprod"], t_next, x.ndim)
            sigma = eta * ((1 - alpha / alpha_next) * (1 - alpha_next) / (1 - alpha)).sqrt()
            c = (1 - alpha_next - sigma ** 2).sqrt()
            x = x_start * alpha_next.sqrt() + c * pred_noise
            if guide is not None and time_next < t_start_guide:
                x = self.guide_gradient_steps(x, hard_conds, guide, **sample_kwargs)
            x = x + sigma * draw(k + 1)
            x = apply_hard_conditioning(x, hard_conds)
            if return_chain:
                chain.append(x)
        if return_chain:
            return x, torch.stack(chain, dim=1)
        return x


# ------------------------------------------------------------------------------------------------
# Convenience: build the oracle-side guide from a synthetic.ProblemSpec
# ------------------------------------------------------------------------------------------------
def hard_conditions(problem, normalize=True, dtype=torch.float32):
    """reference trajectories.py:214-237: {0: [q_start, 0], H-1: [q_goal, 0]} (normalised)."""
    s = torch.cat([torch.as_tensor(problem.start), torch.zeros(problem.robot.q_dim)]).to(dtype)
    g = torch.cat([torch.as_tensor(problem.goal), torch.zeros(problem.robot.q_dim)]).to(dtype)
    if normalize:
        mins, maxs = torch.as_tensor(problem.mins).to(dtype), torch.as_tensor(problem.maxs).to(dtype)
        s, g = limits_normalize(s, mins, maxs), limits_normalize(g, mins, maxs)
    return {0: s, problem.n_support_points - 1: g}


def build_grid_fields(problem, texels_list=None):
    env = problem.env
    fields = []
    sets = [(env.spheres, env.boxes)]
    if np.asarray(env.extra_spheres).size or np.asarray(env.extra_boxes).size:
        sets.append((env.extra_spheres, env.extra_boxes))
    for k, (sp, bx) in enumerate(sets):
        if texels_list is not None:
            fields.append(GridSDF(env.limits, env.cell, texels_list[k], env.grid_shape))
        else:
            fields.append(GridSDF.build(env.limits, env.cell, env.grid_shape, sp, bx))
    return fields


def default_self_pairs(robot, min_frame_gap=4):
    """every pair of collision spheres whose frames are at least four links apart (nearer frames of the Panda sit at fixed or nearly fixed distances: their hinge would be a constant)"""
    fr = [int(f) for f in robot.sphere_frame]
    n = len(fr)
    return [(a, b) for a in range(n) for b in range(a + 1, n) if abs(fr[a] - fr[b]) >= min_frame_gap]


def make_guide_spec(problem, weight_collision, weight_smoothness, texels_list=None, n_interp=128, self_collision=True,
                    **kw) -> GuideSpec:
    if self_collision and problem.robot.kind == "panda" and "self_pairs" not in kw:
        kw["self_pairs"] = default_self_pairs(problem.robot)
    return GuideSpec(robot=problem.robot, mins=torch.as_tensor(problem.mins), maxs=torch.as_tensor(problem.maxs),
                     grid_fields=build_grid_fields(problem, texels_list), border_limits=problem.env.limits,
                     cutoff_margin=problem.cutoff_margin, dt=problem.dt, weight_collision=weight_collision,
                     weight_smoothness=weight_smoothness, n_interp=n_interp, **kw)


# ------------------------------------------------------------------------------------------------
# Post-sampling evaluation (reference inference.py:288-326; torch_robotics sources absent -> unpinned)
# ------------------------------------------------------------------------------------------------
def eval_trajectories(spec: GuideSpec, x_unnormalized, margin=0.0):
    """Per trajectory: #interpolated waypoints with sdf - radius < margin in any field, smoothness, path length,
    minimum clearance."""
    q = spec.robot.q_dim
    xi = interpolate_points(x_unnormalized, spec.n_interp)
    cen = sphere_centers(spec.robot, xi[..., :q])                      # [B, NI, S, ws]
    r = torch.as_tensor(spec.robot.sphere_radius, dtype=x_unnormalized.dtype)
    clear = []
    for g in spec.grid_fields:
        clear.append(g(cen) - r)
    if spec.border_limits is not None:
        clear.append(border_sdf(cen, spec.border_limits) - r)
    clear = torch.stack(clear, 0)                                      # [F, B, NI, S]
    bad = (clear < margin).any(0).any(-1)                              # [B, NI]
    pos, vel = x_unnormalized[..., :q], x_unnormalized[..., q:2 * q]
    smooth = torch.linalg.norm(torch.diff(vel, dim=-2), dim=-1).sum(-1)
    length = torch.linalg.norm(torch.diff(pos, dim=-2), dim=-1).sum(-1)
    return {"n_waypoints_in_collision": bad.sum(-1).float(), "smoothness": smooth, "path_length": length,
            "min_clearance": clear.amin(dim=(0, 2, 3))}
