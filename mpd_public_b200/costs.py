"""Planning-cost descriptors: CostCollision, CostGPTrajectory, CostComposite and the distance fields.

The reference builds these from `mp_baselines` / `torch_robotics` (inference.py:14,195-225), whose
sources are absent; constructor signatures follow the call sites, arithmetic follows
the frozen spec of SURVEY Appendix C. These objects hold *descriptions* (fields, weights, sigmas); the
arithmetic — FK, SDF lookup, hinge, adjoint, GP stencil — runs in csrc/guide.cu.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


class GridSDFField:
    """Voxel grid of {sdf, grad} texels over the environment limits (Appendix C.5), resident in HBM."""

    def __init__(self, limits, cell, texels, shape, cutoff_margin=None):
        self.limits = np.asarray(limits, dtype=np.float64)
        self.cutoff_margin = cutoff_margin  # None: the robot's / task's obstacle_cutoff_margin (inference.py:110)
        self.cell = float(cell)
        self.shape = tuple(int(s) for s in shape)
        self.dim = len(self.shape)
        self.texels = texels  # CUDA fp32 [prod(shape), 1 + dim]
        _lib.require_cuda(texels, "texels")

    @classmethod
    def from_primitives(cls, limits, cell, shape, spheres, boxes, device, cutoff_margin=None):
        """Samples analytic spheres/boxes on the grid with the CUDA builder (mpdb_sdf_grid_build)."""
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("mpd_public_b200 has no CPU path: GridSDFField must be built on a CUDA device")
        dim = len(shape)
        n = int(np.prod(shape))
        tex = torch.empty((n, 1 + dim), device=device, dtype=torch.float32)
        sp = np.ascontiguousarray(np.asarray(spheres, dtype=np.float32).reshape(-1, dim + 1))
        bx = np.ascontiguousarray(np.asarray(boxes, dtype=np.float32).reshape(-1, 2 * dim))
        shp = (C.c_int32 * 3)(*list(shape) + [1] * (3 - dim))
        lo = (C.c_float * 3)(*[float(v) for v in np.asarray(limits)[0]] + [0.0] * (3 - dim))
        with torch.cuda.device(device):
            _lib.check(_lib.lib().mpdb_sdf_grid_build(
                dim, shp, lo, float(cell), sp.ctypes.data_as(C.POINTER(C.c_float)), sp.shape[0],
                bx.ctypes.data_as(C.POINTER(C.c_float)), bx.shape[0], _lib.fptr(tex), _lib.stream_ptr(device)))
        return cls(limits, cell, tex, shape, cutoff_margin)


class WorkspaceBoundaryField:
    """Analytic box: distance to the nearest workspace wall, positive inside (Appendix C.4)."""

    def __init__(self, limits, cutoff_margin=None):
        self.limits = np.asarray(limits, dtype=np.float64)
        self.cutoff_margin = cutoff_margin


class SelfCollisionField:
    """Robot self-collision field (Appendix C.4; torch_robotics source absent -> restated, PARITY UNPINNED): the cost of an
    interpolated row is sum over the listed sphere pairs (a, b) of relu(margin - (|c_a - c_b| - r_a - r_b)).

    `pairs`: iterable of (a, b) sphere indices; default = every pair of spheres whose frames are at least four links
    apart (nearer frames of the Panda sit at fixed or nearly fixed distances: their hinge would be a constant)."""

    def __init__(self, robot, pairs=None, cutoff_margin=0.05, min_frame_gap=4):
        n = int(robot.n_spheres)
        if pairs is None:
            fr = [int(f) for f in robot.sphere_frame]
            pairs = [(a, b) for a in range(n) for b in range(a + 1, n) if abs(fr[a] - fr[b]) >= min_frame_gap]
        self.pairs = sorted({(min(int(a), int(b)), max(int(a), int(b))) for a, b in pairs})
        for a, b in self.pairs:
            if not (0 <= a < b < n):
                raise ValueError(f"bad sphere pair {(a, b)} for a robot with {n} spheres")
        self.cutoff_margin = float(cutoff_margin)

    def pair_masks(self, n_spheres):
        m = [0] * n_spheres
        for a, b in self.pairs:
            m[a] |= 1 << b
            m[b] |= 1 << a
        return m


class CostCollision:
    """CostCollision(robot, n_support_points, field=..., sigma_coll=1.0, tensor_args=...) — inference.py:197-202"""

    def __init__(self, robot, n_support_points, field=None, sigma_coll=1.0, tensor_args=None, **kwargs):
        if field is None:
            raise ValueError("CostCollision needs a field")
        if not float(sigma_coll) > 0.0:
            raise ValueError("sigma_coll must be positive")
        self.robot, self.n_support_points, self.field, self.sigma_coll = robot, n_support_points, field, float(sigma_coll)

    @property
    def cutoff_margin(self):
        """the field's own margin, else the robot's / task's obstacle_cutoff_margin (inference.py:110), else 0.05"""
        m = getattr(self.field, "cutoff_margin", None)
        if m is None:
            m = getattr(self.robot, "cutoff_margin", None)
        return 0.05 if m is None else float(m)


class CostGPTrajectory:
    """CostGPTrajectory(robot, n_support_points, dt, sigma_gp=1.0, tensor_args=...) — inference.py:208-211"""

    def __init__(self, robot, n_support_points, dt, sigma_gp=1.0, tensor_args=None, **kwargs):
        self.robot, self.n_support_points, self.dt, self.sigma_gp = robot, n_support_points, float(dt), float(sigma_gp)


class CostComposite:
    """CostComposite(robot, n_support_points, cost_list, weights_cost_l=..., tensor_args=...) — inference.py:221-225"""

    def __init__(self, robot, n_support_points, cost_list, weights_cost_l=None, tensor_args=None, **kwargs):
        self.robot, self.n_support_points = robot, n_support_points
        self.cost_l = list(cost_list)
        self.weights_cost_l = list(weights_cost_l) if weights_cost_l is not None else [1.0] * len(self.cost_l)
        if len(self.weights_cost_l) != len(self.cost_l):
            raise ValueError("weights_cost_l must have one weight per cost")
        n_gp = sum(isinstance(c, CostGPTrajectory) for c in self.cost_l)
        if n_gp > 1:
            raise NotImplementedError("at most one CostGPTrajectory per composite")
        for c in self.cost_l:
            if not isinstance(c, (CostCollision, CostGPTrajectory)):
                raise NotImplementedError(f"unsupported cost {type(c).__name__}")
