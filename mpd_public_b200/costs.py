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

    def __init__(self, limits, cell, texels, shape):
        self.limits = np.asarray(limits, dtype=np.float64)
        self.cell = float(cell)
        self.shape = tuple(int(s) for s in shape)
        self.dim = len(self.shape)
        self.texels = texels  # CUDA fp32 [prod(shape), 1 + dim]
        _lib.require_cuda(texels, "texels")

    @classmethod
    def from_primitives(cls, limits, cell, shape, spheres, boxes, device):
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
        return cls(limits, cell, tex, shape)


class WorkspaceBoundaryField:
    """Analytic box: distance to the nearest workspace wall, positive inside (Appendix C.4)."""

    def __init__(self, limits):
        self.limits = np.asarray(limits, dtype=np.float64)


class CostCollision:
    """CostCollision(robot, n_support_points, field=..., sigma_coll=1.0, tensor_args=...) — inference.py:197-202"""

    def __init__(self, robot, n_support_points, field=None, sigma_coll=1.0, tensor_args=None, **kwargs):
        if field is None:
            raise ValueError("CostCollision needs a field")
        if float(sigma_coll) != 1.0:
            raise NotImplementedError("only sigma_coll=1.0 is on the inference path (inference.py:200)")
        self.robot, self.n_support_points, self.field, self.sigma_coll = robot, n_support_points, field, sigma_coll


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
