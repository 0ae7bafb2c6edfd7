"""Seeded synthetic problems for the guided-sampling hot path.

There is no network, so neither trained checkpoints nor the trajectory datasets of the
reference (README.md:68-73) exist here. Everything the path needs as *input* is generated
deterministically from numpy's PCG64 (same numpy on every box of this image):

* UNet weights in the reference's state-dict layout (SURVEY.md Appendix A;
  reference `mpd/models/diffusion_models/temporal_unet.py:22-116`,
  `mpd/models/layers/layers.py:229-355`);
* environments (analytic spheres / boxes) standing in for torch_robotics'
  EnvSimple2D / EnvDense2D / EnvNarrowPassageDense2D / EnvSpheres3D (sources absent);
* robot descriptions (point mass; Panda kinematic chain);
* normaliser limits, start/goal pairs.

This module is host-side *problem generation*; it is not on the timed path.
"""
from __future__ import annotations

import math
import zlib
from collections import OrderedDict
from dataclasses import dataclass, field

import numpy as np

UNET_DIM_MULTS = {0: (1, 2, 4), 1: (1, 2, 4, 8)}  # reference temporal_unet.py:14-17

TIME_EMB_DIM = 32  # reference temporal_unet.py:28 (time_emb_dim) and TimeEncoder(32, ...) at :66


# --------------------------------------------------------------------------------------
# UNet state-dict layout
# --------------------------------------------------------------------------------------
def unet_param_shapes(state_dim: int, unet_input_dim: int = 32, dim_mults=(1, 2, 4, 8)) -> "OrderedDict[str, tuple]":
    """Names and shapes of `TemporalUnet(...).state_dict()` for conditioning_type=None.

    Mirrors module registration order of reference temporal_unet.py:60-116.
    """
    dims = [state_dim] + [unet_input_dim * m for m in dim_mults]
    in_out = list(zip(dims[:-1], dims[1:]))
    shapes: "OrderedDict[str, tuple]" = OrderedDict()

    def conv_block(prefix, cin, cout, k=5):
        shapes[f"{prefix}.block.0.weight"] = (cout, cin, k)
        shapes[f"{prefix}.block.0.bias"] = (cout,)
        shapes[f"{prefix}.block.2.weight"] = (cout,)
        shapes[f"{prefix}.block.2.bias"] = (cout,)

    def rtb(prefix, cin, cout):
        conv_block(f"{prefix}.blocks.0", cin, cout)
        conv_block(f"{prefix}.blocks.1", cout, cout)
        shapes[f"{prefix}.cond_mlp.1.weight"] = (cout, TIME_EMB_DIM)
        shapes[f"{prefix}.cond_mlp.1.bias"] = (cout,)
        if cin != cout:
            shapes[f"{prefix}.residual_conv.weight"] = (cout, cin, 1)
            shapes[f"{prefix}.residual_conv.bias"] = (cout,)

    shapes["time_mlp.encoder.1.weight"] = (128, 32)
    shapes["time_mlp.encoder.1.bias"] = (128,)
    shapes["time_mlp.encoder.3.weight"] = (TIME_EMB_DIM, 128)
    shapes["time_mlp.encoder.3.bias"] = (TIME_EMB_DIM,)

    n = len(in_out)
    for i, (cin, cout) in enumerate(in_out):
        rtb(f"downs.{i}.0", cin, cout)
        rtb(f"downs.{i}.1", cout, cout)
        if i < n - 1:
            shapes[f"downs.{i}.4.conv.weight"] = (cout, cout, 3)
            shapes[f"downs.{i}.4.conv.bias"] = (cout,)
    mid = dims[-1]
    rtb("mid_block1", mid, mid)
    rtb("mid_block2", mid, mid)
    for i, (cin, cout) in enumerate(reversed(in_out[1:])):
        rtb(f"ups.{i}.0", cout * 2, cin)
        rtb(f"ups.{i}.1", cin, cin)
        shapes[f"ups.{i}.4.conv.weight"] = (cin, cin, 4)  # ConvTranspose1d: [Cin, Cout, k]
        shapes[f"ups.{i}.4.conv.bias"] = (cin,)
    conv_block("final_conv.0", unet_input_dim, unet_input_dim)
    shapes["final_conv.1.weight"] = (state_dim, unet_input_dim, 1)
    shapes["final_conv.1.bias"] = (state_dim,)
    return shapes


def _rng_for(seed: int, name: str) -> np.random.Generator:
    return np.random.default_rng([seed, zlib.crc32(name.encode())])


def make_unet_state_dict(seed: int, state_dim: int, unet_input_dim: int = 32, dim_mults=(1, 2, 4, 8)):
    """Seeded fp32 numpy weights, PyTorch-default-like scale (uniform ±1/sqrt(fan_in)).

    GroupNorm affine parameters are perturbed away from (1, 0) so the affine path is exercised.
    """
    out = OrderedDict()
    for name, shape in unet_param_shapes(state_dim, unet_input_dim, dim_mults).items():
        rng = _rng_for(seed, name)
        if ".block.2." in name:  # GroupNorm gamma / beta
            if name.endswith("weight"):
                a = 1.0 + 0.2 * rng.uniform(-1, 1, shape)
            else:
                a = 0.2 * rng.uniform(-1, 1, shape)
        else:
            if name.endswith("weight"):
                if "ups." in name and ".4.conv." in name:
                    fan_in = shape[1] * shape[2]  # torch's ConvTranspose fan_in convention
                else:
                    fan_in = int(np.prod(shape[1:]))
            else:
                # bias: use the fan_in of the matching weight (approx: recompute from sibling)
                fan_in = None
            if fan_in is None:
                wname = name[: -len("bias")] + "weight"
                wshape = unet_param_shapes(state_dim, unet_input_dim, dim_mults)[wname]
                fan_in = int(np.prod(wshape[1:]))
            bound = 1.0 / math.sqrt(fan_in)
            a = rng.uniform(-bound, bound, shape)
        out[name] = np.ascontiguousarray(a, dtype=np.float32)
    return out


# --------------------------------------------------------------------------------------
# Robots
# --------------------------------------------------------------------------------------
# Panda kinematic chain: public Franka URDF values (not from the reference tree).
PANDA_JOINT_XYZ = np.array([
    [0.0, 0.0, 0.333],
    [0.0, 0.0, 0.0],
    [0.0, -0.316, 0.0],
    [0.0825, 0.0, 0.0],
    [-0.0825, 0.384, 0.0],
    [0.0, 0.0, 0.0],
    [0.088, 0.0, 0.0],
], dtype=np.float64)
PANDA_JOINT_ROLL = np.array([0.0, -math.pi / 2, math.pi / 2, math.pi / 2, -math.pi / 2, math.pi / 2, math.pi / 2])
PANDA_FLANGE_XYZ = np.array([0.0, 0.0, 0.107])
PANDA_Q_MIN = np.array([-2.8973, -1.7628, -2.8973, -3.0718, -2.8973, -0.0175, -2.8973])
PANDA_Q_MAX = np.array([2.8973, 1.7628, 2.8973, -0.0698, 2.8973, 3.7525, 2.8973])


@dataclass
class RobotSpec:
    """What the collision cost needs to know about a robot (torch_robotics robots, absent).

    `kind` = 'pointmass' (FK = identity on q) or 'panda' (7-DoF chain).
    Collision spheres: sphere i rides on frame `sphere_frame[i]` (1..8 for panda: link 1..7, flange;
    ignored for pointmass) at local offset `sphere_offset[i]` with radius `sphere_radius[i]`.
    """
    kind: str
    q_dim: int
    ws_dim: int
    q_min: np.ndarray
    q_max: np.ndarray
    v_max: float
    sphere_frame: np.ndarray
    sphere_offset: np.ndarray
    sphere_radius: np.ndarray
    dt: float = 0.0

    @property
    def state_dim(self):
        return 2 * self.q_dim

    @property
    def n_spheres(self):
        return int(self.sphere_radius.shape[0])


def robot_pointmass(ws_dim: int = 2, radius: float = 0.01, limit: float = 1.0, v_max: float = 2.0) -> RobotSpec:
    return RobotSpec(
        kind="pointmass", q_dim=ws_dim, ws_dim=ws_dim,
        q_min=-limit * np.ones(ws_dim), q_max=limit * np.ones(ws_dim), v_max=v_max,
        sphere_frame=np.zeros(1, dtype=np.int32), sphere_offset=np.zeros((1, 3)),
        sphere_radius=np.array([radius]))


def robot_panda(v_max: float = 2.5) -> RobotSpec:
    # switch E1: 8 spheres on link-1..7 origins + flange origin, radii 0.1 (hand 0.05)
    return RobotSpec(
        kind="panda", q_dim=7, ws_dim=3, q_min=PANDA_Q_MIN.copy(), q_max=PANDA_Q_MAX.copy(), v_max=v_max,
        sphere_frame=np.arange(1, 9, dtype=np.int32), sphere_offset=np.zeros((8, 3)),
        sphere_radius=np.array([0.1] * 7 + [0.05]))


def panda_fk_numpy(q: np.ndarray):
    """Frame origins and rotations of link 1..7 and flange for q[..., 7] (float64 numpy).

    Returns (origins[..., 8, 3], rotations[..., 8, 3, 3]). T_i = T_{i-1} * Trans(xyz_i) * Rx(roll_i) * Rz(q_i).
    """
    q = np.asarray(q, dtype=np.float64)
    lead = q.shape[:-1]
    R = np.broadcast_to(np.eye(3), lead + (3, 3)).copy()
    o = np.zeros(lead + (3,))
    origins, rots = [], []
    for i in range(7):
        o = o + np.einsum("...ij,j->...i", R, PANDA_JOINT_XYZ[i])
        cr, sr = math.cos(PANDA_JOINT_ROLL[i]), math.sin(PANDA_JOINT_ROLL[i])
        Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
        c, s = np.cos(q[..., i]), np.sin(q[..., i])
        Rz = np.zeros(lead + (3, 3))
        Rz[..., 0, 0] = c; Rz[..., 0, 1] = -s; Rz[..., 1, 0] = s; Rz[..., 1, 1] = c; Rz[..., 2, 2] = 1
        R = R @ Rx @ Rz
        origins.append(o); rots.append(R)
    o = o + np.einsum("...ij,j->...i", R, PANDA_FLANGE_XYZ)
    origins.append(o); rots.append(R)
    return np.stack(origins, axis=-2), np.stack(rots, axis=-3)


def robot_sphere_centers_numpy(robot: RobotSpec, q: np.ndarray) -> np.ndarray:
    """World positions of the collision spheres, [..., n_spheres, ws_dim] (host-side setup helper)."""
    q = np.asarray(q, dtype=np.float64)
    if robot.kind == "pointmass":
        return q[..., None, :]
    o, R = panda_fk_numpy(q)
    f = robot.sphere_frame - 1
    return o[..., f, :] + np.einsum("...sij,sj->...si", R[..., f, :, :], robot.sphere_offset)


# --------------------------------------------------------------------------------------
# Environments
# --------------------------------------------------------------------------------------
@dataclass
class EnvSpec:
    """Analytic obstacle set + the voxel grid it is sampled on."""
    name: str
    dim: int
    limits: np.ndarray            # [2, dim] workspace box (lo, hi)
    spheres: np.ndarray           # [ns, dim+1]  centre, radius
    boxes: np.ndarray             # [nb, 2*dim]  centre, half-size
    cell: float
    extra_spheres: np.ndarray = field(default_factory=lambda: np.zeros((0, 0)))  # "extra objects" field
    extra_boxes: np.ndarray = field(default_factory=lambda: np.zeros((0, 0)))

    @property
    def grid_shape(self):
        n = np.round((self.limits[1] - self.limits[0]) / self.cell).astype(np.int64) + 1
        return tuple(int(v) for v in n)


def _rand_spheres(rng, n, dim, rlo, rhi, lim, keepout=None):
    out = []
    while len(out) < n:
        c = rng.uniform(-lim, lim, dim)
        r = rng.uniform(rlo, rhi)
        if keepout is not None and keepout(c, r):
            continue
        out.append(np.concatenate([c, [r]]))
    return np.array(out).reshape(n, dim + 1)


def _rand_boxes(rng, n, dim, hlo, hhi, lim):
    c = rng.uniform(-lim, lim, (n, dim))
    h = rng.uniform(hlo, hhi, (n, dim))
    return np.concatenate([c, h], axis=1).reshape(n, 2 * dim)


def make_env(name: str, seed: int = 1, cell: float | None = None) -> EnvSpec:
    """Synthetic stand-ins for the reference's four environments."""
    rng = np.random.default_rng([seed, zlib.crc32(name.encode())])
    lim2 = np.array([[-1.0, -1.0], [1.0, 1.0]])
    lim3 = np.array([[-1.0, -1.0, -1.0], [1.0, 1.0, 1.0]])
    z2 = np.zeros((0, 3)); zb2 = np.zeros((0, 4)); z3 = np.zeros((0, 4)); zb3 = np.zeros((0, 6))
    if name == "EnvSimple2D":
        return EnvSpec(name, 2, lim2, _rand_spheres(rng, 4, 2, 0.1, 0.25, 0.8), _rand_boxes(rng, 3, 2, 0.08, 0.2, 0.8),
                       cell or 0.005, _rand_spheres(rng, 1, 2, 0.08, 0.12, 0.6), zb2)
    if name == "EnvDense2D":
        return EnvSpec(name, 2, lim2, _rand_spheres(rng, 15, 2, 0.05, 0.13, 0.9), _rand_boxes(rng, 10, 2, 0.04, 0.1, 0.9),
                       cell or 0.005, z2, zb2)
    if name == "EnvNarrowPassageDense2D":
        sph = _rand_spheres(rng, 12, 2, 0.05, 0.12, 0.9, keepout=lambda c, r: abs(c[0]) < 0.2 + r)
        # a wall at x in [-0.06, 0.06] with one gap of half-height 0.07 around y = 0.25
        boxes = np.array([[0.0, 0.25 + 0.07 + 0.45, 0.06, 0.45], [0.0, 0.25 - 0.07 - 0.70, 0.06, 0.70]])
        boxes = np.concatenate([boxes, _rand_boxes(rng, 6, 2, 0.04, 0.09, 0.9)], axis=0)
        boxes = boxes[[i for i in range(len(boxes)) if i < 2 or abs(boxes[i, 0]) > 0.3]]
        return EnvSpec(name, 2, lim2, sph, boxes, cell or 0.005, z2, zb2)
    if name == "EnvSpheres3D":
        # keep obstacles away from the robot base column so that collision-free configurations exist
        sph = _rand_spheres(rng, 15, 3, 0.1, 0.2, 0.85,
                            keepout=lambda c, r: math.hypot(c[0], c[1]) < 0.25 + r and c[2] < 0.5)
        return EnvSpec(name, 3, lim3, sph, zb3, cell or 0.01, z3, zb3)
    raise KeyError(name)


def sdf_analytic_numpy(p: np.ndarray, spheres: np.ndarray, boxes: np.ndarray) -> np.ndarray:
    """min over primitives of the signed distance at p[..., dim] (float64; host-side setup helper)."""
    p = np.asarray(p, dtype=np.float64)
    dim = p.shape[-1]
    d = np.full(p.shape[:-1], np.inf)
    for s in np.asarray(spheres).reshape(-1, dim + 1):
        d = np.minimum(d, np.linalg.norm(p - s[:dim], axis=-1) - s[dim])
    for b in np.asarray(boxes).reshape(-1, 2 * dim):
        qv = np.abs(p - b[:dim]) - b[dim:]
        d = np.minimum(d, np.linalg.norm(np.maximum(qv, 0.0), axis=-1) + np.minimum(qv.max(axis=-1), 0.0))
    return d


# --------------------------------------------------------------------------------------
# Problems = env + robot + limits + start/goal
# --------------------------------------------------------------------------------------
@dataclass
class ProblemSpec:
    env: EnvSpec
    robot: RobotSpec
    n_support_points: int
    mins: np.ndarray              # [D] normaliser limits
    maxs: np.ndarray
    start: np.ndarray             # [q_dim]
    goal: np.ndarray
    cutoff_margin: float = 0.05   # reference inference.py:110 (obstacle_cutoff_margin)
    trajectory_duration: float = 5.0  # reference inference.py:61

    @property
    def dt(self):
        return self.trajectory_duration / self.n_support_points  # reference inference.py:120


def make_problem(env_name: str, robot_name: str, n_support_points: int = 64, env_seed: int = 1,
                 task_seed: int = 2, cell: float | None = None) -> ProblemSpec:
    env = make_env(env_name, env_seed, cell)
    robot = robot_panda() if robot_name == "RobotPanda" else robot_pointmass(env.dim)
    mins = np.concatenate([robot.q_min, -robot.v_max * np.ones(robot.q_dim)]).astype(np.float32)
    maxs = np.concatenate([robot.q_max, robot.v_max * np.ones(robot.q_dim)]).astype(np.float32)
    rng = np.random.default_rng([task_seed, zlib.crc32((env_name + robot_name).encode())])
    # threshold_start_goal_pos: reference launch_generate_trajectories.py:13-16
    thresh = 1.83 if robot.kind == "panda" else 1.0
    sph_all = np.concatenate([env.spheres.reshape(-1, env.dim + 1), env.extra_spheres.reshape(-1, env.dim + 1)])
    box_all = np.concatenate([env.boxes.reshape(-1, 2 * env.dim), env.extra_boxes.reshape(-1, 2 * env.dim)])

    def free(q):
        c = robot_sphere_centers_numpy(robot, q)
        d = sdf_analytic_numpy(c, sph_all, box_all) - robot.sphere_radius
        inside = np.all((c > env.limits[0] + 0.05) & (c < env.limits[1] - 0.05))
        return bool(d.min() > 0.08) and bool(inside)

    start = goal = None
    for _ in range(10000):
        a = rng.uniform(robot.q_min, robot.q_max)
        b = rng.uniform(robot.q_min, robot.q_max)
        if free(a) and free(b) and np.linalg.norm(a - b) > thresh:
            start, goal = a, b
            break
    if start is None:
        raise ValueError("No collision free configuration was found")  # reference inference.py:168-171
    prob = ProblemSpec(env, robot, n_support_points, mins, maxs, start.astype(np.float32), goal.astype(np.float32))
    robot.dt = prob.dt
    return prob


MODEL_IDS = {
    "EnvSimple2D-RobotPointMass": ("EnvSimple2D", "RobotPointMass"),
    "EnvDense2D-RobotPointMass": ("EnvDense2D", "RobotPointMass"),
    "EnvNarrowPassageDense2D-RobotPointMass": ("EnvNarrowPassageDense2D", "RobotPointMass"),
    "EnvSpheres3D-RobotPanda": ("EnvSpheres3D", "RobotPanda"),
}


def make_problem_by_id(model_id: str, n_support_points: int = 64, **kw) -> ProblemSpec:
    e, r = MODEL_IDS[model_id]
    return make_problem(e, r, n_support_points, **kw)
