"""ctypes binding of libmpdb200.so (the C ABI declared in include/mpdb200.h).

The product path has no CPU or PyTorch fallback: if the shared library is missing or a call fails,
a RuntimeError is raised (the reference signals errors with Python exceptions too,
diffusion_model_base.py:72,275).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmpdb200.so")

MAX_LEVELS, MAX_STATE_DIM, MAX_SPHERES, MAX_GRID_FIELDS, MAX_HARD_CONDS = 8, 32, 16, 4, 8
INT32_MAX = 2 ** 31 - 1


class EngineConfig(C.Structure):
    _fields_ = [
        ("state_dim", C.c_int32), ("horizon", C.c_int32), ("unet_input_dim", C.c_int32), ("n_levels", C.c_int32),
        ("dim_mults", C.c_int32 * MAX_LEVELS), ("n_diffusion_steps", C.c_int32), ("predict_epsilon", C.c_int32),
        ("clip_denoised", C.c_int32), ("max_batch", C.c_int32),
    ]


class GuideConfig(C.Structure):
    _fields_ = [
        ("robot_kind", C.c_int32), ("q_dim", C.c_int32), ("ws_dim", C.c_int32), ("n_spheres", C.c_int32),
        ("sphere_frame", C.c_int32 * MAX_SPHERES), ("sphere_offset", (C.c_float * 3) * MAX_SPHERES),
        ("sphere_radius", C.c_float * MAX_SPHERES), ("mins", C.c_float * MAX_STATE_DIM), ("maxs", C.c_float * MAX_STATE_DIM),
        ("n_grid_fields", C.c_int32), ("grid_texels", C.c_void_p * MAX_GRID_FIELDS),
        ("grid_shape", (C.c_int32 * 3) * MAX_GRID_FIELDS), ("grid_lo", (C.c_float * 3) * MAX_GRID_FIELDS),
        ("grid_cell", C.c_float * MAX_GRID_FIELDS), ("margin_grid", C.c_float * MAX_GRID_FIELDS),
        ("sigma_grid", C.c_float * MAX_GRID_FIELDS), ("weight_grid", C.c_float * MAX_GRID_FIELDS),
        ("has_border", C.c_int32), ("border_lo", C.c_float * 3), ("border_hi", C.c_float * 3),
        ("margin_border", C.c_float), ("sigma_border", C.c_float), ("weight_border", C.c_float),
        ("has_self", C.c_int32), ("self_pairs", C.c_uint32 * MAX_SPHERES),
        ("margin_self", C.c_float), ("sigma_self", C.c_float), ("weight_self", C.c_float),
        ("dt", C.c_float), ("sigma_gp", C.c_float), ("weight_gp", C.c_float),
        ("use_gp", C.c_int32), ("clip_grad", C.c_int32), ("max_grad_norm", C.c_float), ("n_interp", C.c_int32),
        ("vel_from_fd", C.c_int32),
    ]


class LoopParams(C.Structure):
    _fields_ = [
        ("n_steps_without_noise", C.c_int32), ("t_start_guide", C.c_int32), ("n_guide_steps", C.c_int32),
        ("scale_grad_by_std", C.c_int32), ("noise_std", C.POINTER(C.c_float)), ("n_hard_conds", C.c_int32),
        ("hard_cond_rows", C.c_int32 * MAX_HARD_CONDS), ("hard_cond_vals", C.c_void_p), ("use_cuda_graph", C.c_int32),
        ("horizon", C.c_int32), ("state_dim", C.c_int32),
    ]


class DdimParams(C.Structure):
    _fields_ = [
        ("n_steps", C.c_int32), ("times", C.POINTER(C.c_int32)), ("times_next", C.POINTER(C.c_int32)),
        ("sqrt_alpha_next", C.POINTER(C.c_float)), ("coef_noise", C.POINTER(C.c_float)), ("t_start_guide", C.c_int32),
        ("n_guide_steps", C.c_int32), ("n_hard_conds", C.c_int32), ("hard_cond_rows", C.c_int32 * MAX_HARD_CONDS),
        ("hard_cond_vals", C.c_void_p), ("horizon", C.c_int32), ("state_dim", C.c_int32),
    ]


# every symbol include/mpdb200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "mpdb_last_error": (C.c_char_p, []),
    "mpdb_version": (C.c_int, []),
    "mpdb_engine_create": (C.c_int, [C.POINTER(EngineConfig), C.c_int, C.POINTER(_P)]),
    "mpdb_engine_destroy": (None, [_P]),
    "mpdb_engine_set_param": (C.c_int, [_P, C.c_char_p, _P, C.c_int64, _P]),
    "mpdb_engine_set_schedule": (C.c_int, [_P] + [C.POINTER(C.c_float)] * 7),
    "mpdb_engine_set_option": (C.c_int, [_P, C.c_char_p, C.c_double]),
    "mpdb_engine_finalize": (C.c_int, [_P, _P]),
    "mpdb_unet_forward": (C.c_int, [_P, _P, _P, _P, C.c_int32, _P]),
    "mpdb_engine_generation": (C.c_int64, [_P]),
    "mpdb_unet_forward_uniform": (C.c_int, [_P, _P, C.c_int32, _P, C.c_int32, _P]),
    "mpdb_engine_mega_info": (C.c_int, [_P, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                        C.POINTER(C.c_int32), C.c_char_p, C.c_int]),
    "mpdb_engine_read_mega_timeline": (C.c_int, [_P, C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.c_int32]),
    "mpdb_profile_unet_body": (C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_float), C.POINTER(C.c_double),
                                         C.POINTER(C.c_int32), _P]),
    "mpdb_p_mean": (C.c_int, [_P, _P, _P, _P, C.c_int32, _P]),
    "mpdb_add_noise": (C.c_int, [_P, _P, _P, _P, C.c_float, C.c_int32, _P]),
    "mpdb_sample_loop": (C.c_int, [_P, _P, C.POINTER(LoopParams), _P, _P, _P, C.c_int64, C.c_int64, C.c_int32, _P]),
    "mpdb_ddim_loop": (C.c_int, [_P, _P, C.POINTER(DdimParams), _P, _P, _P, C.c_int64, C.c_int64, C.c_int32, _P]),
    "mpdb_guide_steps_chain": (C.c_int, [_P, _P, C.c_int32, _P, C.c_int32, C.POINTER(C.c_int32), _P, _P, C.c_int64, C.c_int32,
                                         C.c_int32, _P]),
    "mpdb_normal_fill": (C.c_int, [_P, C.c_int64, C.c_int32, _P, C.c_int, _P]),
    "mpdb_limits_normalize": (C.c_int, [_P, C.c_int64, C.c_int32, _P, _P, _P, C.c_int32, C.c_int, _P]),
    "mpdb_normal_offset_increment": (C.c_int64, [C.c_int64, C.c_int]),
    "mpdb_launch_count": (C.c_int64, []),
    "mpdb_engine_num_ops": (C.c_int, [_P]),
    "mpdb_profile_forward": (C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_float), C.POINTER(C.c_double),
                                       C.POINTER(C.c_int32), _P]),
    "mpdb_profile_guide": (C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_float), _P]),
    "mpdb_profile_guide_steps": (C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_float), _P]),
    "mpdb_engine_step_precision": (C.c_int, [_P, C.c_int32]),
    "mpdb_guide_max_coresident": (C.c_int, [_P, C.c_int32]),
    "mpdb_debug_tc_conv5": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P]),
    "mpdb_engine_read_timeline": (C.c_int, [_P, C.POINTER(C.c_int64), C.c_int32]),
    "mpdb_engine_num_buffers": (C.c_int, [_P]),
    "mpdb_engine_buffer_info": (C.c_int, [_P, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "mpdb_engine_read_buffer": (C.c_int, [_P, C.c_int, _P, C.c_int32, _P]),
    "mpdb_guide_create": (C.c_int, [C.POINTER(GuideConfig), C.c_int, C.POINTER(_P)]),
    "mpdb_guide_destroy": (None, [_P]),
    "mpdb_guide_grad": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, _P]),
    "mpdb_guide_grad_pos": (C.c_int, [_P, _P, _P, _P, C.c_int32, C.c_int32, _P]),
    "mpdb_guide_steps": (C.c_int, [_P, _P, C.c_int32, _P, C.c_int32, C.POINTER(C.c_int32), _P, C.c_int32, C.c_int32, _P]),
    "mpdb_guide_record_decisions": (C.c_int, [_P, _P, C.c_int64, C.c_int32]),
    "mpdb_guide_decisions_recorded": (C.c_int64, [_P]),
    "mpdb_guide_num_collision_costs": (C.c_int, [_P]),
    "mpdb_guide_batch_dependent_clamps": (C.c_int64, [_P, C.c_int32]),
    "mpdb_debug_fk": (C.c_int, [_P, _P, _P, C.c_int32, _P]),
    "mpdb_eval_trajectories": (C.c_int, [_P, _P, _P, C.c_float, C.c_int32, C.c_int32, _P]),
    "mpdb_sdf_grid_build": (C.c_int, [C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_float), C.c_float,
                                      C.POINTER(C.c_float), C.c_int32, C.POINTER(C.c_float), C.c_int32, _P, _P]),
}

_lib = None


def lib():
    """Loads libmpdb200.so once; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: the CUDA library is not built. Run `python -c 'import __graft_entry__ as g; "
            f"g.build()'` (or `make -C mpd_public_b200/csrc`). There is no CPU fallback.")
    l = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(l, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = l
    return l


def check(status: int):
    if status != 0:
        msg = lib().mpdb_last_error()
        raise RuntimeError(f"mpdb200: {msg.decode() if msg else 'unknown error'} (status {status})")


def launch_count() -> int:
    return int(lib().mpdb_launch_count())


def require_cuda(t, what="tensor"):
    import torch
    if not (torch.is_tensor(t) and t.is_cuda):
        raise RuntimeError(f"mpd_public_b200: {what} must live on a CUDA device — this framework has no CPU path "
                           f"(got {getattr(t, 'device', type(t))})")


def stream_ptr(device):
    import torch
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def fptr(t):
    """Raw device pointer of a contiguous fp32 CUDA tensor."""
    import torch
    assert t.dtype == torch.float32 and t.is_contiguous(), (t.dtype, t.is_contiguous())
    return C.c_void_p(t.data_ptr())
