"""GuideManagerTrajectoriesWithVelocity (reference `mpd/models/diffusion_models/guides.py:149-236`, the manager
inference.py builds) and the position-only GuideManagerTrajectories (:15-146).

Same constructor arguments (including the **kwargs that swallow inference.py's misspelt
`num_interpolated_points`) and call convention `guide(x_normalized [B,H,D]) -> grad [B,H,D]`.
The body — unnormalise, interpolate, composite cost, one backward per cost, clip-by-norm (+1e-6),
endpoint zeroing, weighting, negation — is one CUDA kernel with a hand-derived adjoint (csrc/guide.cu).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .costs import (CostCollision, CostComposite, CostGPTrajectory, GridSDFField, SelfCollisionField,
                    WorkspaceBoundaryField)


def build_guide_config(robot, mins, maxs, collision_costs, gp, clip_grad, max_grad_norm, n_interp, vel_from_fd=False):
    """mpdb_guide_config from a robot description, normaliser limits, [(field, weight, cutoff margin, sigma_coll)] and
    (dt, sigma_gp, weight) | None. Returns (cfg, tensors kept alive because the config references them by raw pointer)."""
    cfg = _lib.GuideConfig()
    cfg.robot_kind = 1 if robot.kind == "panda" else 0
    cfg.q_dim, cfg.ws_dim, cfg.n_spheres = robot.q_dim, robot.ws_dim, robot.n_spheres
    for i in range(robot.n_spheres):
        cfg.sphere_frame[i] = int(robot.sphere_frame[i])
        for k in range(3):
            cfg.sphere_offset[i][k] = float(robot.sphere_offset[i][k])
        cfg.sphere_radius[i] = float(robot.sphere_radius[i])
    if len(mins) == robot.q_dim:  # position-only trajectories (GuideManagerTrajectories): the velocity half is not normalised
        mins = list(mins) + [0.0] * robot.q_dim
        maxs = list(maxs) + [1.0] * robot.q_dim
    if len(mins) != 2 * robot.q_dim:
        raise RuntimeError("normaliser limits must cover [q, qdot] (or [q] for position-only trajectories)")
    for i in range(len(mins)):
        cfg.mins[i], cfg.maxs[i] = float(mins[i]), float(maxs[i])
    keep = []
    n_grid = 0
    cfg.has_border = cfg.has_self = 0
    order = []  # position of every cost in the kernel's order (grids, border, self): the reference sums them in list order
    for f, w, margin, sigma in collision_costs:
        if isinstance(f, GridSDFField):
            if n_grid >= _lib.MAX_GRID_FIELDS:
                raise RuntimeError(f"at most {_lib.MAX_GRID_FIELDS} grid-backed collision fields")
            if f.dim != robot.ws_dim:
                raise RuntimeError("grid field and robot disagree on the workspace dimension")
            for k in range(f.dim):
                cfg.grid_shape[n_grid][k] = f.shape[k]
                cfg.grid_lo[n_grid][k] = float(f.limits[0][k])
            cfg.grid_cell[n_grid] = f.cell
            cfg.grid_texels[n_grid] = f.texels.data_ptr()
            cfg.weight_grid[n_grid], cfg.margin_grid[n_grid], cfg.sigma_grid[n_grid] = float(w), float(margin), float(sigma)
            keep.append(f.texels)
            order.append(("grid", n_grid))
            n_grid += 1
        elif isinstance(f, WorkspaceBoundaryField):
            if cfg.has_border:
                raise NotImplementedError("one workspace-boundary field")
            cfg.has_border = 1
            for k in range(f.limits.shape[1]):
                cfg.border_lo[k], cfg.border_hi[k] = float(f.limits[0][k]), float(f.limits[1][k])
            cfg.weight_border, cfg.margin_border, cfg.sigma_border = float(w), float(margin), float(sigma)
            order.append(("border", 0))
        elif isinstance(f, SelfCollisionField):
            if cfg.has_self:
                raise NotImplementedError("one self-collision field")
            if robot.kind != "panda":
                continue  # a point mass has nothing to collide with
            cfg.has_self = 1
            for a, m in enumerate(f.pair_masks(robot.n_spheres)):
                cfg.self_pairs[a] = m
            cfg.weight_self, cfg.margin_self, cfg.sigma_self = float(w), float(margin), float(sigma)
            order.append(("self", 0))
        else:
            raise NotImplementedError(f"unsupported field {type(f).__name__}")
    # the kernel adds the weighted cost gradients as grids, border, self; fp32 addition is not associative, so a composite
    # listing them in another order is refused rather than silently summed differently from the reference's list order
    rank = {"grid": 0, "border": 1, "self": 2}
    if [rank[k] for k, _ in order] != sorted(rank[k] for k, _ in order):
        raise NotImplementedError("collision costs must be listed as: grid-backed fields, workspace boundary, self-collision")
    cfg.n_grid_fields = n_grid
    if gp is not None:
        cfg.use_gp = 1
        cfg.dt, cfg.sigma_gp, cfg.weight_gp = float(gp[0]), float(gp[1]), float(gp[2])
    else:
        cfg.use_gp = 0
        cfg.dt, cfg.sigma_gp = float(getattr(robot, "dt", 0.0) or 1.0), 1.0
    cfg.clip_grad = int(bool(clip_grad))
    cfg.max_grad_norm = float(max_grad_norm)
    cfg.n_interp = int(n_interp)
    cfg.vel_from_fd = int(bool(vel_from_fd))
    return cfg, keep


def _dataset_limits(dataset):
    """mins/maxs of the trajectory normaliser, as the reference reaches them (trajectories.py:196-197)."""
    nz = dataset.normalizer
    if hasattr(nz, "normalizers"):
        nz = nz.normalizers[dataset.field_key_traj]
    return nz.mins.detach().float().cpu().numpy(), nz.maxs.detach().float().cpu().numpy()


def const_vel_trajectory(start_state_pos, goal_state_pos, dt, num_steps, q_dim, set_initial_final_vel_to_zero=False,
                         device=None):
    """Restatement of `MultiMPPrior.const_vel_trajectory` (mp_baselines@8a50c3c, source absent; call site guides.py:46-53):
    num_steps + 1 states on the straight line start -> goal with the constant velocity (goal - start) / (num_steps * dt),
    optionally zero at both ends. Returns [num_steps + 1, 2 * q_dim]."""
    start = torch.as_tensor(start_state_pos, dtype=torch.float32, device=device)[:q_dim]
    goal = torch.as_tensor(goal_state_pos, dtype=torch.float32, device=device)[:q_dim]
    lam = torch.arange(num_steps + 1, dtype=torch.float32, device=device)[:, None] / float(num_steps)
    pos = start[None] * (1.0 - lam) + goal[None] * lam
    vel = ((goal - start) / (num_steps * dt))[None].repeat(num_steps + 1, 1)
    if set_initial_final_vel_to_zero:
        vel[0] = 0.0
        vel[-1] = 0.0
    return torch.cat((pos, vel), dim=-1)


class GuideManagerTrajectories(nn.Module):
    """Position-only guide manager with a persistent velocity trajectory — mirror of reference guides.py:15-146
    (scripts/inference/inference.py itself builds GuideManagerTrajectoriesWithVelocity, :229).

    `guide(x_pos_normalized [B,H,q]) -> grad [B,H,q]`; each call also moves `self.velocity` against the costs' velocity
    gradients (guides.py:110-112). One CUDA kernel (csrc/guide.cu, position-only mode): unnormalise the positions,
    interpolate, composite cost on [positions | velocity], per cost clip the position and the velocity gradient separately
    (norm + 1e-6), zero the end rows, weight, negate. With `use_velocity_from_finite_difference=True` (guides.py:77-79) the
    velocity half of the state is `robot.get_velocity(x_pos)` — torch_robotics' source is absent; restated as the central
    difference (p[h+1] - p[h-1]) / (2 dt) with zero end rows (oracle switch FD_CENTRAL, parity unpinned) — the costs reach the
    positions through it as well, a single position gradient is clipped per cost and the velocity trajectory is left alone."""
    _mpdb_fusable = False  # stateful (velocity): the reverse loop calls it step by step, as the reference does

    def __init__(self, dataset, cost, clip_grad=False, clip_grad_rule='norm', max_grad_norm=1., max_grad_value=0.1,
                 interpolate_trajectories_for_collision=False, num_interpolated_points_for_collision=128,
                 use_velocity_from_finite_difference=False, start_state_pos=None, goal_state_pos=None, num_steps=100,
                 robot=None, n_samples=1, tensor_args=None, **kwargs):
        super().__init__()
        if not isinstance(cost, CostComposite):
            raise NotImplementedError("cost must be a mpd_public_b200.CostComposite")
        if clip_grad and clip_grad_rule != 'norm':
            raise NotImplementedError("only clip_grad_rule='norm' is implemented (guides.py:127)")
        if start_state_pos is None or goal_state_pos is None:
            raise RuntimeError("start_state_pos and goal_state_pos are required (guides.py:46-53)")
        self.cost = cost
        self.dataset = dataset
        self.interpolate_trajectories_for_collision = interpolate_trajectories_for_collision
        self.num_interpolated_points_for_collision = num_interpolated_points_for_collision
        self.clip_grad = clip_grad
        self.clip_grad_rule = clip_grad_rule
        self.max_grad_norm = max_grad_norm
        self.max_grad_value = max_grad_value
        self.use_velocity_from_finite_difference = use_velocity_from_finite_difference
        self.robot = robot if robot is not None else cost.robot
        self.start_state_pos = start_state_pos
        self.goal_state_pos = goal_state_pos
        device = (tensor_args or {}).get("device", "cuda")
        dt = float(getattr(self.robot, "dt", 1.0))
        traj = const_vel_trajectory(start_state_pos, goal_state_pos, dt, num_steps, self.robot.q_dim,
                                    set_initial_final_vel_to_zero=True, device=device)
        vel = traj[:, self.robot.q_dim:]                                   # robot.get_velocity
        self.velocity = vel[None].repeat(n_samples, 1, 1).contiguous()      # 'H D -> B H D'
        self._handles = {}

    _config = None  # set below: shared with GuideManagerTrajectoriesWithVelocity
    _handle = None

    def __del__(self):
        try:
            for h in self._handles.values():
                _lib.lib().mpdb_guide_destroy(h[0])
            self._handles = {}
        except Exception:
            pass

    def forward(self, x_pos_normalized):
        _lib.require_cuda(x_pos_normalized, "x_pos_normalized")
        x = x_pos_normalized.detach().to(torch.float32).contiguous()
        B, H, q = x.shape
        if q != self.robot.q_dim:
            raise RuntimeError(f"expected position-only trajectories [B, H, {self.robot.q_dim}], got {tuple(x.shape)}")
        grad = torch.empty_like(x)
        if self.use_velocity_from_finite_difference:
            _lib.check(_lib.lib().mpdb_guide_grad_pos(self._handle(x.device, H), _lib.fptr(x), None, _lib.fptr(grad), B, H,
                                                      _lib.stream_ptr(x.device)))
            return grad
        if tuple(self.velocity.shape) != (B, H, q) or self.velocity.device != x.device:
            raise RuntimeError(f"velocity trajectory is {tuple(self.velocity.shape)} on {self.velocity.device}; the guide was "
                               f"built for n_samples={self.velocity.shape[0]}, num_steps={self.velocity.shape[1] - 1}")
        _lib.check(_lib.lib().mpdb_guide_grad_pos(self._handle(x.device, H), _lib.fptr(x), _lib.fptr(self.velocity),
                                                  _lib.fptr(grad), B, H, _lib.stream_ptr(x.device)))
        return grad


class GuideManagerTrajectoriesWithVelocity(nn.Module):
    _mpdb_fusable = True

    def __init__(self, dataset, cost, clip_grad=False, clip_grad_rule='norm', max_grad_norm=1., max_grad_value=0.1,
                 interpolate_trajectories_for_collision=False, num_interpolated_points_for_collision=128,
                 start_state_pos=None, goal_state_pos=None, num_steps=100, robot=None, n_samples=1, tensor_args=None,
                 **kwargs):
        super().__init__()
        if not isinstance(cost, CostComposite):
            raise NotImplementedError("cost must be a mpd_public_b200.CostComposite")
        if clip_grad and clip_grad_rule != 'norm':
            raise NotImplementedError("only clip_grad_rule='norm' is on the inference path (guides.py:151)")
        self.cost = cost
        self.dataset = dataset
        self.interpolate_trajectories_for_collision = interpolate_trajectories_for_collision
        self.num_interpolated_points_for_collision = num_interpolated_points_for_collision
        self.clip_grad = clip_grad
        self.clip_grad_rule = clip_grad_rule
        self.max_grad_norm = max_grad_norm
        self.max_grad_value = max_grad_value
        self._handles = {}

    # ---- C-ABI handle ----
    def _signature(self):
        """Everything the device-side configuration is built from, cheap to compare: the reference re-reads these attributes
        on every call, so a handle built from older values must not be reused (weights, clip settings, limits, ...)."""
        mins_t = self.dataset.normalizer
        if hasattr(mins_t, "normalizers"):
            mins_t = mins_t.normalizers[self.dataset.field_key_traj]
        costs = tuple((id(c), getattr(c, "sigma_coll", None), getattr(c, "cutoff_margin", None), getattr(c, "dt", None),
                       getattr(c, "sigma_gp", None)) for c in self.cost.cost_l)
        return (costs, tuple(float(w) for w in self.cost.weights_cost_l), bool(self.clip_grad), float(self.max_grad_norm),
                bool(self.interpolate_trajectories_for_collision), int(self.num_interpolated_points_for_collision),
                mins_t.mins.data_ptr(), mins_t.mins._version, mins_t.maxs.data_ptr(), mins_t.maxs._version,
                bool(getattr(self, "use_velocity_from_finite_difference", False)), float(getattr(self.cost.robot, "dt", 0.0) or 0.0))

    def _config(self, H):
        mins, maxs = _dataset_limits(self.dataset)
        coll, gp = [], None
        for c, w in zip(self.cost.cost_l, self.cost.weights_cost_l):
            if isinstance(c, CostCollision):
                coll.append((c.field, float(w), c.cutoff_margin, c.sigma_coll))
            elif isinstance(c, CostGPTrajectory):
                gp = (c.dt, c.sigma_gp, float(w))
        n_interp = int(self.num_interpolated_points_for_collision) if self.interpolate_trajectories_for_collision else int(H)
        return build_guide_config(self.cost.robot, mins, maxs, coll, gp, self.clip_grad, self.max_grad_norm, n_interp,
                                  vel_from_fd=getattr(self, "use_velocity_from_finite_difference", False))

    def _handle(self, device, H):
        device = torch.device(device)
        idx = device.index if device.index is not None else torch.cuda.current_device()
        key = (idx, int(H))
        sig = self._signature()
        h = self._handles.get(key)
        if h is not None and h[3] != sig:  # an attribute changed since the handle was built: rebuild (captured graphs are
            _lib.lib().mpdb_guide_destroy(h[0])  # keyed on the handle value + the config hash, so they are re-captured too)
            h = None
        if h is None:
            cfg, keep = self._config(H)
            for t in keep:
                if t.device.index != idx:
                    raise RuntimeError("collision grids live on a different device than the trajectories")
            handle = C.c_void_p()
            _lib.check(_lib.lib().mpdb_guide_create(C.byref(cfg), idx, C.byref(handle)))
            h = (handle, keep, cfg, sig)
            self._handles[key] = h
        return h[0]

    def batch_dependent_clamps(self, reset=False):
        """(trajectory, evaluation) pairs, over all handles of this guide, whose normaliser clamp was decided by other
        trajectories of the batch (normalization.py:160-162 looks at the whole tensor). 0 = results so far do not depend
        on the batch composition, i.e. any sharding of the batch reproduces them bit for bit."""
        return sum(int(_lib.lib().mpdb_guide_batch_dependent_clamps(h[0], int(bool(reset)))) for h in self._handles.values())

    # ---- parity instrumentation: the discrete decisions (texel index, hinge / wall / pair activity) of every evaluation ----
    def record_decisions(self, device, H, B, capacity_evals):
        """Starts recording into a fresh int32 buffer [capacity_evals, B, n_costs, n_interp, n_spheres]; returns it. Loops run
        without CUDA graphs while recording. `stop_recording` ends it."""
        handle = self._handle(device, H)
        cfg = self._handles[(torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device(), int(H))][2]
        n_costs = int(_lib.lib().mpdb_guide_num_collision_costs(handle))
        buf = torch.zeros((int(capacity_evals), int(B), n_costs, int(cfg.n_interp), int(cfg.n_spheres)), device=device, dtype=torch.int32)
        _lib.check(_lib.lib().mpdb_guide_record_decisions(handle, C.c_void_p(buf.data_ptr()), int(capacity_evals), int(B)))
        self._recording = (handle, buf)
        return buf

    def decisions_recorded(self):
        return int(_lib.lib().mpdb_guide_decisions_recorded(self._recording[0])) if getattr(self, "_recording", None) else 0

    def stop_recording(self):
        rec = getattr(self, "_recording", None)
        if rec is not None:
            _lib.check(_lib.lib().mpdb_guide_record_decisions(rec[0], None, 0, 0))
            self._recording = None

    def __del__(self):
        try:
            for h in self._handles.values():
                _lib.lib().mpdb_guide_destroy(h[0])
            self._handles = {}
        except Exception:
            pass

    # ---- reference interface ----
    def forward(self, x_normalized):
        _lib.require_cuda(x_normalized, "x_normalized")
        x = x_normalized.detach().to(torch.float32).contiguous()
        B, H, D = x.shape
        grad = torch.empty_like(x)
        _lib.check(_lib.lib().mpdb_guide_grad(self._handle(x.device, H), _lib.fptr(x), _lib.fptr(grad), B, H,
                                              _lib.stream_ptr(x.device)))
        return grad

    def guide_steps(self, x, hard_conds, n_guide_steps, model_var=None, return_chain=False):
        """n x {x <- x + guide(x) [* model_var]; hard conditioning} as n kernel launches; returns a new tensor — with
        `return_chain` also every iterate [n, B, H, D] (the diffusion_prior_then_guide post-loop, inference.py:263-282)."""
        _lib.require_cuda(x, "x")
        x = x.detach().to(torch.float32).clone(memory_format=torch.contiguous_format)
        B, H, D = x.shape
        chain = torch.empty((int(n_guide_steps), B, H, D), device=x.device, dtype=torch.float32) if return_chain else None
        rows = list(hard_conds.keys())
        hc_rows = (C.c_int32 * max(len(rows), 1))(*[int(r) % H for r in rows])
        hc = torch.stack([hard_conds[r].to(device=x.device, dtype=torch.float32).expand(B, D) for r in rows]).contiguous() \
            if rows else None
        mv = None
        if model_var is not None:
            mv = model_var.to(device=x.device, dtype=torch.float32).reshape(-1).expand(B).contiguous()
        _lib.check(_lib.lib().mpdb_guide_steps_chain(
            self._handle(x.device, H), _lib.fptr(x), int(n_guide_steps), _lib.fptr(mv) if mv is not None else None,
            len(rows), hc_rows, _lib.fptr(hc) if hc is not None else None,
            _lib.fptr(chain) if chain is not None and chain.numel() else None, B * H * D, B, H, _lib.stream_ptr(x.device)))
        return (x, chain) if return_chain else x

    def clip_gradient(self, grad):
        if self.clip_grad:
            if self.clip_grad_rule == 'norm':
                return self.clip_grad_by_norm(grad)
            raise NotImplementedError
        return grad

    def clip_grad_by_norm(self, grad):
        """reference :224-230 (torch ops; the fused kernel applies the same rule per cost)."""
        if self.clip_grad:
            grad_norm = torch.linalg.norm(grad + 1e-6, dim=-1, keepdims=True)
            scale_ratio = torch.clip(grad_norm, 0., self.max_grad_norm) / grad_norm
            grad = scale_ratio * grad
        return grad


GuideManagerTrajectories._signature = GuideManagerTrajectoriesWithVelocity._signature
GuideManagerTrajectories._config = GuideManagerTrajectoriesWithVelocity._config
GuideManagerTrajectories._handle = GuideManagerTrajectoriesWithVelocity._handle
