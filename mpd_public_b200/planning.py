"""Synthetic stand-ins for the objects inference.py pulls from `torch_robotics` and the dataset
(`trajectories.py:21-237`): robot, environment/task with its collision fields, and a dataset that
exposes the attributes the hot path touches (normaliser limits, hard conditions).
Real dataset / checkpoint ingestion is SURVEY §8(f).3 ("next").
"""
from __future__ import annotations

import numpy as np
import torch

from . import synthetic as S
from .costs import GridSDFField, SelfCollisionField, WorkspaceBoundaryField
from .normalization import DatasetNormalizer, LimitsNormalizer


class Robot:
    """What CostCollision/CostGPTrajectory and get_hard_conditions need (torch_robotics robots, absent)."""

    def __init__(self, spec: S.RobotSpec, cutoff_margin=0.05):
        self.spec = spec
        self.kind, self.q_dim, self.ws_dim = spec.kind, spec.q_dim, spec.ws_dim
        self.sphere_frame, self.sphere_offset, self.sphere_radius = spec.sphere_frame, spec.sphere_offset, spec.sphere_radius
        self.n_spheres = spec.n_spheres
        self.cutoff_margin = cutoff_margin
        self.dt = spec.dt

    def get_position(self, x):
        return x[..., :self.q_dim]

    def get_velocity(self, x):
        return x[..., self.q_dim:2 * self.q_dim]


class PlanningTask:
    def __init__(self, env: S.EnvSpec, robot: Robot, device, obstacle_cutoff_margin=0.05, use_extra_objects=True,
                 use_self_collision=True, self_collision_margin=0.05):
        self.env, self.robot, self.device = env, robot, torch.device(device)
        robot.cutoff_margin = obstacle_cutoff_margin
        self.use_extra_objects = use_extra_objects
        self.use_self_collision = use_self_collision
        self.self_collision_margin = self_collision_margin
        self._fields = None

    def get_collision_fields(self):
        """[objects grid, (extra objects grid), workspace boundaries, (robot self-collision: articulated robots only)] —
        inference.py:193."""
        if self._fields is None:
            e = self.env
            f = [GridSDFField.from_primitives(e.limits, e.cell, e.grid_shape, e.spheres, e.boxes, self.device)]
            if self.use_extra_objects and (np.asarray(e.extra_spheres).size or np.asarray(e.extra_boxes).size):
                f.append(GridSDFField.from_primitives(e.limits, e.cell, e.grid_shape, e.extra_spheres, e.extra_boxes,
                                                      self.device))
            f.append(WorkspaceBoundaryField(e.limits))
            if self.use_self_collision and self.robot.kind == "panda":
                f.append(SelfCollisionField(self.robot, cutoff_margin=self.self_collision_margin))
            self._fields = f
        return self._fields

    def random_coll_free_q(self, n_samples=1, max_tries=10000, clearance=0.08, seed=None):
        """torch_robotics PlanningTask.random_coll_free_q (call site inference.py:161): uniformly sampled configurations whose
        collision spheres keep `clearance` from every object and stay inside the workspace. Host-side rejection sampling on the
        analytic primitives (not on the timed path). Returns [n_samples, q_dim] on the task's device."""
        rng = np.random.default_rng(seed if seed is not None else int(torch.randint(0, 2 ** 31 - 1, (1,))))
        spec, e = self.robot.spec, self.env
        sph = np.concatenate([np.asarray(e.spheres).reshape(-1, e.dim + 1), np.asarray(e.extra_spheres).reshape(-1, e.dim + 1)])
        box = np.concatenate([np.asarray(e.boxes).reshape(-1, 2 * e.dim), np.asarray(e.extra_boxes).reshape(-1, 2 * e.dim)])
        out = []
        for _ in range(max_tries):
            q = rng.uniform(spec.q_min, spec.q_max)
            c = S.robot_sphere_centers_numpy(spec, q)
            d = S.sdf_analytic_numpy(c, sph, box) - spec.sphere_radius
            if d.min() > clearance and np.all((c > e.limits[0] + 0.05) & (c < e.limits[1] - 0.05)):
                out.append(q)
                if len(out) == n_samples:
                    return torch.as_tensor(np.stack(out), dtype=torch.float32, device=self.device)
        raise ValueError("No collision free configuration was found")

    def get_collision_fields_extra_objects(self):
        return [f for f in self.get_collision_fields()[1:] if isinstance(f, GridSDFField)]

    # ---- post-sampling evaluation (reference inference.py:288-326; torch_robotics PlanningTask, sources absent) ----
    def evaluate_trajectories(self, trajs, margin=0.0, n_interp=128):
        """One fused kernel over UNNORMALISED trajectories [B, H, D] -> dict of per-trajectory tensors:
        n_waypoints_in_collision, smoothness, path_length, min_clearance, in_collision (bool)."""
        import ctypes as C
        from . import _lib
        from .guides import build_guide_config
        _lib.require_cuda(trajs, "trajs")
        x = trajs.detach().to(torch.float32).contiguous()
        B, H, D = x.shape
        idx = x.device.index if x.device.index is not None else torch.cuda.current_device()
        key = (idx, int(n_interp))
        cache = self.__dict__.setdefault("_eval_handles", {})
        if key not in cache:
            zeros = np.zeros(2 * self.robot.q_dim, dtype=np.float32)
            fields = [f for f in self.get_collision_fields() if not isinstance(f, SelfCollisionField)]
            cfg, keep = build_guide_config(self.robot, zeros, zeros + 1, [(f, 1.0, 0.0, 1.0) for f in fields], None, False, 1.0,
                                           n_interp)
            handle = C.c_void_p()
            _lib.check(_lib.lib().mpdb_guide_create(C.byref(cfg), idx, C.byref(handle)))
            cache[key] = (handle, keep)
        stats = torch.empty((B, 4), device=x.device, dtype=torch.float32)
        _lib.check(_lib.lib().mpdb_eval_trajectories(cache[key][0], _lib.fptr(x), _lib.fptr(stats), float(margin), B, H,
                                                     _lib.stream_ptr(x.device)))
        return {"n_waypoints_in_collision": stats[:, 0], "smoothness": stats[:, 1], "path_length": stats[:, 2],
                "min_clearance": stats[:, 3], "in_collision": stats[:, 0] > 0, "n_interp": n_interp}

    def get_trajs_collision_and_free(self, trajs, return_indices=False, **kw):
        """(trajs_coll, [idxs_coll], trajs_free, [idxs_free], None) — None in place of an empty set, as upstream."""
        ev = self.evaluate_trajectories(trajs, **kw)
        coll = ev["in_collision"]
        idx_coll, idx_free = torch.nonzero(coll).flatten(), torch.nonzero(~coll).flatten()
        tc = trajs[idx_coll] if idx_coll.numel() else None
        tf = trajs[idx_free] if idx_free.numel() else None
        if return_indices:
            return tc, idx_coll, tf, idx_free, None
        return tc, tf, None

    def compute_fraction_free_trajs(self, trajs, **kw):
        return float((~self.evaluate_trajectories(trajs, **kw)["in_collision"]).float().mean())

    def compute_collision_intensity_trajs(self, trajs, **kw):
        ev = self.evaluate_trajectories(trajs, **kw)
        return float(ev["n_waypoints_in_collision"].sum() / (trajs.shape[0] * ev["n_interp"]))

    def compute_success_free_trajs(self, trajs, **kw):
        return int(bool((~self.evaluate_trajectories(trajs, **kw)["in_collision"]).any()))

    def best_free_trajectory(self, trajs, **kw):
        """argmin of path length + smoothness over the collision-free plans (inference.py:316-320); (index, cost) or None."""
        ev = self.evaluate_trajectories(trajs, **kw)
        cost = ev["path_length"] + ev["smoothness"]
        cost = torch.where(ev["in_collision"], torch.full_like(cost, float("inf")), cost)
        i = int(torch.argmin(cost))
        return None if bool(ev["in_collision"][i]) else (i, float(cost[i]))


def compute_smoothness(trajs, robot):
    """sum_h |v_{h+1} - v_h| per trajectory (torch_robotics metric, call site inference.py:312)."""
    return torch.linalg.norm(torch.diff(robot.get_velocity(trajs), dim=-2), dim=-1).sum(-1)


def compute_path_length(trajs, robot):
    """sum_h |p_{h+1} - p_h| per trajectory (call site inference.py:315)."""
    return torch.linalg.norm(torch.diff(robot.get_position(trajs), dim=-2), dim=-1).sum(-1)


def compute_variance_waypoints(trajs, robot):
    """variance of the waypoint positions across the trajectory set, summed over dimensions, averaged over the horizon
    (call site inference.py:325)."""
    return float(torch.var(robot.get_position(trajs), dim=0).sum(-1).mean())


class TrajectoryDataset:
    """Synthetic dataset: normaliser limits = joint / velocity limits."""

    field_key_traj = 'traj'

    def __init__(self, problem: S.ProblemSpec, device, include_velocity=True, use_extra_objects=True,
                 obstacle_cutoff_margin=0.05, use_self_collision=True, **kwargs):
        self.problem = problem
        self.device = torch.device(device)
        self.include_velocity = include_velocity
        self.n_support_points = problem.n_support_points
        self.robot = Robot(problem.robot, obstacle_cutoff_margin)
        self.env = problem.env
        self.task = PlanningTask(problem.env, self.robot, device, obstacle_cutoff_margin, use_extra_objects,
                                 use_self_collision=use_self_collision)
        self.state_dim = problem.robot.state_dim if include_velocity else problem.robot.q_dim
        self.threshold_start_goal_pos = 1.83 if problem.robot.kind == "panda" else 1.0
        # position-only datasets (include_velocity=False, trajectories.py:60-70) normalise the q positions only
        lim = torch.stack([torch.as_tensor(problem.mins), torch.as_tensor(problem.maxs)])[:, :self.state_dim].to(self.device)
        self.normalizer = DatasetNormalizer({self.field_key_traj: lim}, LimitsNormalizer)

    def unnormalize_trajectories(self, x):
        return self.normalizer.unnormalize(x, self.field_key_traj)

    def normalize_trajectories(self, x):
        return self.normalizer.normalize(x, self.field_key_traj)

    def get_hard_conditions(self, traj, horizon=None, normalize=False):
        """reference trajectories.py:214-237"""
        # start and goal go through the same elementwise operations as in the reference, as one [2, D] batch (half the
        # launches of treating them one after the other; the values are identical)
        ends = traj if traj.shape[0] == 2 else torch.stack((traj[0], traj[-1]))
        pos = self.robot.get_position(ends)
        if normalize and self.include_velocity:
            # cat(pos, zeros) normalised in one launch (a call used to be seven small torch launches on the critical path of
            # run_inference: 0.1 ms of idle GPU per end-to-end call)
            state = self.normalizer.normalize(pos, key=self.field_key_traj, pad_to=2 * pos.shape[-1])
        else:
            state = torch.cat((pos, torch.zeros_like(pos)), dim=-1) if self.include_velocity else pos
            if normalize:
                state = self.normalizer.normalize(state, key=self.field_key_traj)
        if horizon is None:
            horizon = self.n_support_points
        return {0: state[0], horizon - 1: state[1]}
