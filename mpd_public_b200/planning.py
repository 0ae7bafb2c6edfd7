"""Synthetic stand-ins for the objects inference.py pulls from `torch_robotics` and the dataset
(`trajectories.py:21-237`): robot, environment/task with its collision fields, and a dataset that
exposes the attributes the hot path touches (normaliser limits, hard conditions).
Real dataset / checkpoint ingestion is SURVEY §8(f).3 ("next").
"""
from __future__ import annotations

import numpy as np
import torch

from . import synthetic as S
from .costs import GridSDFField, WorkspaceBoundaryField
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
    def __init__(self, env: S.EnvSpec, robot: Robot, device, obstacle_cutoff_margin=0.05, use_extra_objects=True):
        self.env, self.robot, self.device = env, robot, torch.device(device)
        robot.cutoff_margin = obstacle_cutoff_margin
        self.use_extra_objects = use_extra_objects
        self._fields = None

    def get_collision_fields(self):
        """[objects grid, (extra objects grid), workspace boundaries] — inference.py:193 (Appendix C.4)."""
        if self._fields is None:
            e = self.env
            f = [GridSDFField.from_primitives(e.limits, e.cell, e.grid_shape, e.spheres, e.boxes, self.device)]
            if self.use_extra_objects and (np.asarray(e.extra_spheres).size or np.asarray(e.extra_boxes).size):
                f.append(GridSDFField.from_primitives(e.limits, e.cell, e.grid_shape, e.extra_spheres, e.extra_boxes,
                                                      self.device))
            f.append(WorkspaceBoundaryField(e.limits))
            self._fields = f
        return self._fields

    def get_collision_fields_extra_objects(self):
        return self.get_collision_fields()[1:-1]


class TrajectoryDataset:
    """Synthetic dataset: normaliser limits = joint / velocity limits."""

    field_key_traj = 'traj'

    def __init__(self, problem: S.ProblemSpec, device, include_velocity=True, use_extra_objects=True,
                 obstacle_cutoff_margin=0.05, **kwargs):
        self.problem = problem
        self.device = torch.device(device)
        self.include_velocity = include_velocity
        self.n_support_points = problem.n_support_points
        self.robot = Robot(problem.robot, obstacle_cutoff_margin)
        self.env = problem.env
        self.task = PlanningTask(problem.env, self.robot, device, obstacle_cutoff_margin, use_extra_objects)
        self.state_dim = problem.robot.state_dim
        self.threshold_start_goal_pos = 1.83 if problem.robot.kind == "panda" else 1.0
        lim = torch.stack([torch.as_tensor(problem.mins), torch.as_tensor(problem.maxs)]).to(self.device)
        self.normalizer = DatasetNormalizer({self.field_key_traj: lim}, LimitsNormalizer)

    def unnormalize_trajectories(self, x):
        return self.normalizer.unnormalize(x, self.field_key_traj)

    def normalize_trajectories(self, x):
        return self.normalizer.normalize(x, self.field_key_traj)

    def get_hard_conditions(self, traj, horizon=None, normalize=False):
        """reference trajectories.py:214-237"""
        start_state_pos = self.robot.get_position(traj[0])
        goal_state_pos = self.robot.get_position(traj[-1])
        if self.include_velocity:
            start_state = torch.cat((start_state_pos, torch.zeros_like(start_state_pos)), dim=-1)
            goal_state = torch.cat((goal_state_pos, torch.zeros_like(goal_state_pos)), dim=-1)
        else:
            start_state, goal_state = start_state_pos, goal_state_pos
        if normalize:
            start_state = self.normalizer.normalize(start_state, key=self.field_key_traj)
            goal_state = self.normalizer.normalize(goal_state, key=self.field_key_traj)
        if horizon is None:
            horizon = self.n_support_points
        return {0: start_state, horizon - 1: goal_state}
