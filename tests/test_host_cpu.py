"""CPU-only checks: the C-ABI library loads and exports what include/mpdb200.h declares, struct layouts
agree between C and ctypes, the host mirror keeps the reference's state-dict and schedule, the product
refuses to run without CUDA, and the sharding helpers work under gloo with world_size 2."""
import ctypes
import json
import os
import re
import subprocess
import sys
import tempfile

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

from tests.golden import cases as C  # noqa: E402


def test_library_exports_every_declared_symbol():
    from mpd_public_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "mpdb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(mpdb_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    lib = _lib.lib()
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.mpdb_version() >= 100  # host-only call, no GPU needed


def test_struct_layouts_match_c():
    from mpd_public_b200 import _lib
    src = r'''
#include <stdio.h>
#include <stddef.h>
#include "mpdb200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(mpdb_engine_config), sizeof(mpdb_guide_config), sizeof(mpdb_loop_params),
         offsetof(mpdb_guide_config, grid_texels), offsetof(mpdb_guide_config, n_interp),
         offsetof(mpdb_loop_params, noise_std), offsetof(mpdb_loop_params, hard_cond_vals),
         offsetof(mpdb_guide_config, self_pairs), offsetof(mpdb_guide_config, margin_grid), offsetof(mpdb_guide_config, vel_from_fd),
         offsetof(mpdb_loop_params, state_dim), sizeof(mpdb_ddim_params), offsetof(mpdb_ddim_params, hard_cond_vals));
  return 0; }
'''
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "t"), os.path.join(d, "t.c")], check=True)
        out = subprocess.run([os.path.join(d, "t")], check=True, capture_output=True, text=True).stdout.split()
    got = [int(v) for v in out]
    G, L, Dd = _lib.GuideConfig, _lib.LoopParams, _lib.DdimParams
    want = [ctypes.sizeof(_lib.EngineConfig), ctypes.sizeof(G), ctypes.sizeof(L), G.grid_texels.offset, G.n_interp.offset,
            L.noise_std.offset, L.hard_cond_vals.offset, G.self_pairs.offset, G.margin_grid.offset, G.vel_from_fd.offset,
            L.state_dim.offset, ctypes.sizeof(Dd), Dd.hard_cond_vals.offset]
    assert got == want


def test_state_dict_layout_matches_reference():
    import mpd_public_b200 as M
    keys = json.load(open(os.path.join(C.GOLDEN_DIR, "state_dict_keys.json")))
    for case, ref in keys.items():
        d, h, opt, seed = C.UNET_CASES[case]
        unet = M.TemporalUnet(n_support_points=h, state_dim=d, unet_input_dim=32, dim_mults=M.UNET_DIM_MULTS[opt])
        model = M.GaussianDiffusionModel(model=unet, n_diffusion_steps=C.T_DIFF, predict_epsilon=True)
        mine = {k: list(v.shape) for k, v in model.state_dict().items()}
        assert mine == ref  # same keys, same shapes as the reference's GaussianDiffusionModel ...
        assert list(mine.keys()) == list(ref.keys())  # ... in the same order (dict equality ignores it): what torch.save would hold
        model.load_state_dict({"model." + k: torch.as_tensor(v) for k, v in C.unet_weights(case).items()}, strict=False)


def test_schedule_buffers_bit_exact_vs_reference():
    import mpd_public_b200 as M
    g = C.load("schedule")
    unet = M.TemporalUnet(n_support_points=64, state_dim=4, dim_mults=(1, 2, 4))
    for sched, T in (("exponential", 25), ("cosine", 20)):
        m = M.GaussianDiffusionModel(model=unet, variance_schedule=sched, n_diffusion_steps=T, predict_epsilon=True)
        for k, v in m.state_dict().items():
            if not k.startswith("model."):
                assert np.array_equal(v.numpy(), g[f"{sched}.{k}"]), (sched, k)
    with pytest.raises(NotImplementedError):
        M.GaussianDiffusionModel(model=unet, variance_schedule="linear")


def test_normalizer_matches_reference_golden():
    import mpd_public_b200 as M
    g = C.load("normalizer")
    prob = C.guide_problem("panda3d")
    nz = M.LimitsNormalizer(torch.stack([torch.as_tensor(prob.mins), torch.as_tensor(prob.maxs)]))
    assert np.array_equal(nz.unnormalize(torch.as_tensor(C.guide_input("panda3d"))).numpy(), g["un_in"])
    assert np.array_equal(nz.unnormalize(torch.as_tensor(C.guide_input("panda3d", out_of_range=True))).numpy(), g["un_out"])


def test_panda_kinematic_constants_are_pinned_independently():
    """the oracle and the product each hold their own copy of the Panda chain constants (the oracle does not
    import the product's). Both are checked against each other AND against answers from outside this repository: the public
    Franka Panda geometry puts the flange of the zero configuration at (0.088, 0, 0.926) and the link origins at the sums
    below. A wrong joint origin on both sides would have to be the same wrong number twice and still hit these."""
    import math
    from mpd_public_b200 import synthetic as S
    from oracle import mpd_oracle as O
    assert np.allclose(np.asarray(O.PANDA_JOINT_XYZ), S.PANDA_JOINT_XYZ, atol=0) and np.allclose(O.PANDA_JOINT_ROLL, S.PANDA_JOINT_ROLL, atol=0)
    assert np.allclose(O.PANDA_FLANGE_XYZ, S.PANDA_FLANGE_XYZ, atol=0)
    import inspect
    src = inspect.getsource(O)
    assert "synthetic" not in src.split("def sphere_centers")[1].split("def sdf_and_grad_analytic")[0], "the oracle's FK must not import product constants"
    zero = np.array([[0, 0, 0.333], [0, 0, 0.333], [0, 0, 0.649], [0.0825, 0, 0.649], [0, 0, 1.033], [0, 0, 1.033],
                     [0.088, 0, 1.033], [0.088, 0, 0.926]])
    robot = S.robot_panda()
    q0 = torch.zeros(1, 7, dtype=torch.float64)
    assert np.abs(O.sphere_centers(robot, q0)[0].numpy() - zero).max() < 1e-12
    assert np.abs(S.robot_sphere_centers_numpy(robot, np.zeros(7)) - zero).max() < 1e-12
    # joint 1 turns everything about the world z axis; joint 4 at -pi/2 folds the forearm: link 5 origin known in closed form
    q = torch.zeros(1, 7, dtype=torch.float64)
    q[0, 0] = math.pi / 2
    rotz = np.array([[0.0, -1.0, 0.0], [1.0, 0.0, 0.0], [0.0, 0.0, 1.0]])
    assert np.abs(O.sphere_centers(robot, q)[0].numpy() - zero @ rotz.T).max() < 1e-12
    q = torch.zeros(1, 7, dtype=torch.float64)
    q[0, 3] = -math.pi / 2
    c = O.sphere_centers(robot, q)[0].numpy()
    # link 4 frame at q4 = -pi/2: the (-0.0825, 0.384) offset of link 5 (in link 4's x / y, i.e. world x / z at q = 0) turns by -90 deg about
    # link 4's axis (world -y at q = 0): world offset (0.384, 0, 0.0825)
    assert np.abs(c[4] - (zero[3] + np.array([0.384, 0.0, 0.0825]))).max() < 1e-12, c[4]
    # rigid link lengths for a random configuration
    q = torch.tensor([[0.3, -0.5, 0.2, -1.9, 0.4, 1.7, -0.6]], dtype=torch.float64)
    c = O.sphere_centers(robot, q)[0].numpy()
    assert abs(np.linalg.norm(c[2] - c[0]) - 0.316) < 1e-12 and abs(np.linalg.norm(c[4] - c[3]) - math.hypot(0.0825, 0.384)) < 1e-12
    assert abs(np.linalg.norm(c[7] - c[6]) - 0.107) < 1e-12 and abs(np.linalg.norm(c[6] - c[5]) - 0.088) < 1e-12
    assert np.abs(c - S.robot_sphere_centers_numpy(robot, q[0].numpy())).max() < 1e-12


def test_oracle_decision_audit_and_forced_decisions():
    """oracle/parity.py's machinery on the CPU: decisions encoded the way the CUDA guide records them (include/mpdb200.h),
    produced here from the oracle's own choices, must (a) decode, (b) reproduce the free-running gradient exactly when taken
    over, (c) pass the audit; a corrupted decision away from any boundary must fail it."""
    from oracle import mpd_oracle as O
    prob = C.guide_problem("panda3d")
    spec = O.make_guide_spec(prob, 1e-2, 1e-7)
    spec.self_margin = 0.35
    x = torch.as_tensor(C.guide_input("panda3d"))
    rep = []
    ref = O.guide_manager_grad(spec, x, report=rep)
    B, NI, S_ = x.shape[0], spec.n_interp, prob.robot.n_spheres
    dec = torch.zeros((B, 3, NI, S_), dtype=torch.int64)
    dec[:, 0] = (rep[0]["flat"] << 1) | (rep[0]["hinge"] > 0).long()
    walls = rep[1]["walls"]
    w = walls.argmin(-1)
    dec[:, 1] = ((w % 3) << 2) | ((w < 3).long() << 1) | (rep[1]["hinge"] > 0).long()
    act = rep[2]["hinge"] > 0                                   # [B, NI, n_pairs]
    for k, (a, b) in enumerate(spec.self_pairs):
        dec[:, 2, :, a] |= act[..., k].long() << b
        dec[:, 2, :, b] |= act[..., k].long() << a
    assert int(act.sum()) > 0 and int((rep[0]["hinge"] > 0).sum()) > 0
    dec = dec.to(torch.int32)
    forced = O.guide_manager_grad(spec, x, decisions=dec)
    assert torch.equal(forced, ref)
    audit = O.audit_decisions(spec, x, dec)
    assert audit["unexplained"] == 0 and audit["index_diff"] == 0 and audit["hinge_diff"] == 0 and audit["wall_diff"] == 0
    bad = dec.clone()
    far = (rep[0]["hinge"].abs() > 0.01).nonzero()[0]
    bad[far[0], 0, far[1], far[2]] ^= 1                         # flip a hinge far from its boundary
    assert O.audit_decisions(spec, x, bad)["unexplained"] >= 1
    moved = dec.clone()
    inner = (rep[0]["round_dist"] > 0.1).nonzero()[0]
    moved[inner[0], 0, inner[1], inner[2]] += 2                 # neighbouring texel although the point sits well inside its cell
    assert O.audit_decisions(spec, x, moved)["unexplained"] >= 1


def test_no_cpu_fallback():
    import mpd_public_b200 as M
    unet = M.TemporalUnet(n_support_points=64, state_dim=4, dim_mults=(1, 2, 4))
    model = M.GaussianDiffusionModel(model=unet, n_diffusion_steps=25, predict_epsilon=True)
    with pytest.raises(RuntimeError, match="no CPU path"):
        model.run_inference(None, {0: torch.zeros(4)}, n_samples=2, horizon=64)
    with pytest.raises(RuntimeError, match="CUDA"):
        unet(torch.zeros(2, 64, 4), torch.zeros(2, dtype=torch.long), None)
    with pytest.raises(RuntimeError, match="no CPU path"):
        M.GridSDFField.from_primitives(np.array([[-1, -1], [1, 1.0]]), 0.1, (21, 21), np.zeros((0, 3)), np.zeros((0, 4)), "cpu")
    # the product never imports the oracle
    for root, _, files in os.walk(os.path.join(ROOT, "mpd_public_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(root, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_hard_conditions_and_problem_generation():
    import mpd_public_b200 as M
    from oracle import mpd_oracle as O
    for mid in C.S.MODEL_IDS:
        prob = C.S.make_problem_by_id(mid, 64, cell=0.05)
        ds = M.TrajectoryDataset(prob, "cpu")
        hc = ds.get_hard_conditions(torch.vstack((torch.as_tensor(prob.start), torch.as_tensor(prob.goal))), normalize=True)
        ref = O.hard_conditions(prob)
        assert set(hc) == {0, 63}
        for k in hc:
            assert torch.allclose(hc[k], ref[k], atol=1e-6)
            assert hc[k].abs().max() <= 1.0


def test_shard_bounds_cover_batch():
    from mpd_public_b200.parallel import shard_bounds
    for n in (1, 7, 100, 4096):
        for w in (1, 2, 3, 8):
            b = [shard_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1


_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from mpd_public_b200.parallel import sample_sharded, shard_bounds
dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{sys.argv[2]}", rank=int(sys.argv[3]), world_size=2)
for n_total in (8, 7):
    g = torch.Generator().manual_seed(0)
    noise = torch.randn((3, n_total, 4, 2), generator=g)
    def sample_local(n, nz):
        assert nz.shape[1] == n
        return nz.sum(0) * 2.0   # stand-in for the per-shard sampler: any per-trajectory function
    out = sample_sharded(sample_local, n_total, noise=noise)
    assert torch.equal(out, noise.sum(0) * 2.0), "gathered shards must equal the single-process result"
dist.destroy_process_group()
print("ok")
'''


def test_sharded_sampling_gloo_world2():
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "w.py")
        open(path, "w").write(_WORKER)
        procs = [subprocess.Popen([sys.executable, path, ROOT, str(port), str(r)], stdout=subprocess.PIPE,
                                  stderr=subprocess.STDOUT, text=True) for r in range(2)]
        outs = [p.communicate(timeout=120)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0 and "ok" in o, o


def test_checkpoint_and_dataset_ingestion_roundtrip():
    """SURVEY §8f.3 on a synthetic directory with the reference's on-disk layout."""
    import yaml
    import mpd_public_b200 as M
    from mpd_public_b200 import ingest
    with tempfile.TemporaryDirectory() as d:
        os.makedirs(os.path.join(d, "model", "checkpoints"))
        os.makedirs(os.path.join(d, "data", "0"))
        args = dict(variance_schedule="exponential", n_diffusion_steps=25, predict_epsilon=True, unet_input_dim=32,
                    unet_dim_mults_option=0, diffusion_model_class="GaussianDiffusionModel", use_ema=True)
        yaml.dump(args, open(os.path.join(d, "model", "args.yaml"), "w"))
        unet = M.TemporalUnet(n_support_points=64, state_dim=4, dim_mults=M.UNET_DIM_MULTS[0])
        ref = M.GaussianDiffusionModel(model=unet, n_diffusion_steps=25, predict_epsilon=True)
        ref.load_state_dict({"model." + k: torch.as_tensor(v) for k, v in C.unet_weights("pm2d_opt0_h64").items()}, strict=False)
        torch.save(ref.state_dict(), os.path.join(d, "model", "checkpoints", "ema_model_current_state_dict.pth"))
        g = torch.Generator().manual_seed(0)
        trajs = torch.rand((50, 64, 4), generator=g) * 2 - 1
        torch.save(trajs, os.path.join(d, "data", "0", "trajs-free.pt"))

        model, a = ingest.load_diffusion_model(os.path.join(d, "model"), state_dim=4, n_support_points=64, device="cpu")
        assert a["n_diffusion_steps"] == 25 and not model.training
        for (k1, v1), (k2, v2) in zip(model.state_dict().items(), ref.state_dict().items()):
            assert k1 == k2 and torch.equal(v1, v2)
        nz, h, dim = ingest.load_trajectory_limits(os.path.join(d, "data"), q_dim=2)
        assert (h, dim) == (64, 4)
        lim = nz.normalizers["traj"]
        assert torch.equal(lim.mins, trajs.reshape(-1, 4).min(0).values) and torch.equal(lim.maxs, trajs.reshape(-1, 4).max(0).values)
        x = nz.normalize(trajs, "traj")
        assert float(x.min()) == -1.0 and float(x.max()) == 1.0


def test_const_vel_trajectory_matches_the_oracle_restatement():
    """Velocity trajectory the position-only guide manager starts from (guides.py:46-53)."""
    import torch
    from mpd_public_b200.guides import const_vel_trajectory
    from oracle import mpd_oracle as O
    start, goal = torch.tensor([0.1, -0.4, 0.3]), torch.tensor([0.9, 0.2, -0.5])
    for zero_ends in (False, True):
        a = const_vel_trajectory(start, goal, 0.08, 63, 3, set_initial_final_vel_to_zero=zero_ends, device="cpu")
        b = O.const_vel_trajectory(start, goal, 0.08, 63, 3, set_initial_final_vel_to_zero=zero_ends)
        assert a.shape == (64, 6) and torch.allclose(a, b, atol=1e-6)
        assert torch.equal(a[0, :3], start) and torch.allclose(a[-1, :3], goal)
        assert (a[0, 3:].abs().sum() == 0) == zero_ends
