"""GaussianDiffusionModel — host-side mirror of reference
`mpd/models/diffusion_models/diffusion_model_base.py:46-316` (sampling part).

Same constructor, buffers (names and values bit-identical: the same torch ops on the same inputs),
methods and return conventions, so `scripts/inference/inference.py:138-257` runs unchanged against it.
The arithmetic of every reverse step runs in libmpdb200 (hand-written sm_100a kernels); when
`sample_fn` is this package's `ddpm_sample_fn` and the guide is this package's guide manager (or None)
the whole loop is enqueued by ONE C-ABI call (`mpdb_sample_loop`) with no host synchronisation inside
(the reference has ~150 host syncs per loop) and can be replayed as a CUDA graph.
"""
from __future__ import annotations

import ctypes as C
from copy import copy

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .sample_functions import apply_hard_conditioning, ddpm_sample_fn, extract, guide_gradient_steps  # noqa: F401


# ------------------------------------------------------------------------------------------------
# schedules (reference helpers.py:26-46) — same torch/numpy ops => bit-identical buffers
# ------------------------------------------------------------------------------------------------
def exponential_beta_schedule(n_diffusion_steps, beta_start=1e-4, beta_end=1.0):
    x = torch.linspace(0, n_diffusion_steps, n_diffusion_steps)
    beta_start = torch.tensor(beta_start)
    beta_end = torch.tensor(beta_end)
    a = 1 / n_diffusion_steps * torch.log(beta_end / beta_start)
    return beta_start * torch.exp(a * x)


def cosine_beta_schedule(n_diffusion_steps, s=0.008, a_min=0, a_max=0.999, dtype=torch.float32):
    steps = n_diffusion_steps + 1
    x = np.linspace(0, steps, steps)
    alphas_cumprod = np.cos(((x / steps) + s) / (1 + s) * np.pi * 0.5) ** 2
    alphas_cumprod = alphas_cumprod / alphas_cumprod[0]
    betas = 1 - (alphas_cumprod[1:] / alphas_cumprod[:-1])
    return torch.tensor(np.clip(betas, a_min=a_min, a_max=a_max), dtype=dtype)


class _DeviceNormal:
    """The loop's draws — randn(shape) + one randn_like per step (diffusion_model_base.py:165, sample_functions.py:51) — in
    ONE kernel (csrc/rng.cu) that reproduces, bit for bit, what those `normal_()` calls would have drawn from torch's CUDA
    generator, and advances that generator by exactly what they would have consumed. (seed, offset) travel through a small
    device tensor so the launch can be replayed inside a CUDA graph."""

    def __init__(self, device):
        self.device = torch.device(device)
        self.index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.state = torch.zeros(2, dtype=torch.int64, device=self.device)
        self.lib = _lib.lib()

    @staticmethod
    def supported(device):
        idx = torch.device(device).index
        gen = torch.cuda.default_generators[idx if idx is not None else torch.cuda.current_device()]
        return hasattr(gen, "get_offset") and hasattr(gen, "set_offset")

    def advance(self, numel, n_draws):
        """Reserve n_draws normal_() calls of `numel` elements in torch's generator; uploads (seed, first offset)."""
        gen = torch.cuda.default_generators[self.index]
        inc = int(self.lib.mpdb_normal_offset_increment(int(numel), self.index))
        if inc <= 0:
            raise RuntimeError("mpdb_normal_offset_increment failed")
        seed, off = int(gen.initial_seed()), int(gen.get_offset())
        gen.set_offset(off + n_draws * inc)
        # the seed may use all 64 bits: reinterpret as int64 for the tensor
        self.state.copy_(torch.tensor([seed - (1 << 64) if seed >= (1 << 63) else seed, off], dtype=torch.int64))

    def fill(self, buf):
        """buf: contiguous fp32 [n_draws, ...]; enqueues the fill on the current stream (reads self.state when it runs)."""
        n_draws = buf.shape[0]
        numel = buf[0].numel()
        _lib.check(self.lib.mpdb_normal_fill(_lib.fptr(buf), int(numel), int(n_draws), C.c_void_p(self.state.data_ptr()),
                                             self.index, _lib.stream_ptr(self.device)))


def make_timesteps(batch_size, i, device):
    return torch.full((batch_size,), i, device=device, dtype=torch.long)


# ------------------------------------------------------------------------------------------------
# engine: one libmpdb200 handle per (TemporalUnet, device, schedule)
# ------------------------------------------------------------------------------------------------
class Engine:
    """Owns an `mpdb_engine*`: packed weights, time-conditioning tables, schedule tables, workspace."""

    def __init__(self, unet, device, n_steps, schedule, predict_epsilon, clip_denoised, horizon=None):
        self.lib = _lib.lib()
        self.unet = unet
        # the UNet is fully convolutional (temporal_unet.py:118-171): the reference accepts any horizon divisible by
        # 2^(levels-1); the device plan (buffers, tiles, cluster program) is per horizon, so each horizon gets its own engine
        self.horizon = int(horizon if horizon is not None else unet.n_support_points)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("mpd_public_b200 has no CPU path: move the model to a CUDA device")
        self.n_steps = int(n_steps)
        cfg = _lib.EngineConfig()
        cfg.state_dim = unet.state_dim
        cfg.horizon = self.horizon
        cfg.unet_input_dim = unet.unet_input_dim
        cfg.n_levels = len(unet.dim_mults)
        for i, m in enumerate(unet.dim_mults):
            cfg.dim_mults[i] = m
        cfg.n_diffusion_steps = self.n_steps
        cfg.predict_epsilon = int(bool(predict_epsilon))
        cfg.clip_denoised = int(bool(clip_denoised))
        cfg.max_batch = 1
        self.handle = C.c_void_p()
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.dev_index = dev_index
        _lib.check(self.lib.mpdb_engine_create(C.byref(cfg), dev_index, C.byref(self.handle)))
        self._param_sig = None
        self._tc_mode = None
        self._set_schedule(schedule)

    TC_MODES = {"off": 0, "auto": 1, "force": 2}

    def set_tensor_cores(self, mode):
        """'auto' (default): tcgen05 path (22-bit scaled-fp16 operand split, as accurate as the fp32 FMA path) for every
        convolution, in the fused loop and in the per-call entry points; 'off': exact fp32 FMA path everywhere; 'force':
        tensor cores even when a finite `tc_amp_limit` carve-out is configured."""
        if mode != self._tc_mode:
            _lib.check(self.lib.mpdb_engine_set_option(self.handle, b"tc_mode", float(self.TC_MODES[mode])))
            self._tc_mode = mode

    def __del__(self):
        try:
            if getattr(self, "handle", None) and self.handle.value:
                self.lib.mpdb_engine_destroy(self.handle)
                self.handle = C.c_void_p()
        except Exception:
            pass

    def _set_schedule(self, s):
        T = self.n_steps
        if s is None:  # stand-alone TemporalUnet: identity-ish tables, only unet_forward is meaningful
            z = torch.zeros(T)
            tabs = [torch.ones(T), z, z, z, z, torch.ones(T), torch.ones(T)]
        else:
            logvar = s["posterior_log_variance_clipped"]
            tabs = [s["sqrt_recip_alphas_cumprod"], s["sqrt_recipm1_alphas_cumprod"], s["posterior_mean_coef1"],
                    s["posterior_mean_coef2"], logvar, torch.exp(0.5 * logvar), torch.exp(logvar)]
        arrs = [np.ascontiguousarray(t.detach().float().cpu().numpy()) for t in tabs]
        for a in arrs:
            assert a.shape == (T,)
        self._sched_host = arrs
        ptrs = [a.ctypes.data_as(C.POINTER(C.c_float)) for a in arrs]
        _lib.check(self.lib.mpdb_engine_set_schedule(self.handle, *ptrs))

    def _signature(self):
        # (module, name) slots are cached: walking the module tree costs ~1 ms per call, which would sit in front of every
        # sampling call; reading the slots still sees replaced Parameter objects, .to() and in-place updates
        slots = self.__dict__.get("_param_slots")
        if slots is None:
            slots = [(m, n) for m in self.unet.modules() for n in m._parameters if m._parameters[n] is not None]
            self._param_slots = slots
        return tuple((m._parameters[n].data_ptr(), m._parameters[n]._version) for m, n in slots)

    def sync_params(self):
        """(Re)uploads the UNet parameters when they changed (load_state_dict, .to(), optimiser step)."""
        sig = self._signature()
        if sig == self._param_sig:
            return
        st = _lib.stream_ptr(self.device)
        for name, p in self.unet.named_parameters():
            t = p.detach()
            if t.device != self.device or t.dtype != torch.float32 or not t.is_contiguous():
                t = t.to(device=self.device, dtype=torch.float32).contiguous()
            _lib.check(self.lib.mpdb_engine_set_param(self.handle, name.encode(), _lib.fptr(t), t.numel(), st))
            del t
        _lib.check(self.lib.mpdb_engine_finalize(self.handle, st))
        self._param_sig = sig

    # ---- entry points ----
    def _prep(self, x, t=None):
        _lib.require_cuda(x, "x")
        if x.dim() != 3 or x.shape[1] != self.horizon or x.shape[2] != self.unet.state_dim:
            raise RuntimeError(f"expected x of shape [B, {self.horizon}, {self.unet.state_dim}], got {tuple(x.shape)}")
        x = x.detach().to(torch.float32).contiguous()
        if t is not None:
            t = t.to(device=x.device, dtype=torch.long).contiguous()
            if t.shape != (x.shape[0],):
                raise RuntimeError(f"expected t of shape [{x.shape[0]}], got {tuple(t.shape)}")
        return x, t

    def unet_forward(self, x, t, check_t=True):
        x, t = self._prep(x, t)
        if check_t and (int(t.max()) >= self.n_steps or int(t.min()) < 0):
            raise RuntimeError(f"time index out of range [0, {self.n_steps})")
        self.sync_params()
        out = torch.empty_like(x)
        _lib.check(self.lib.mpdb_unet_forward(self.handle, _lib.fptr(x), C.c_void_p(t.data_ptr()), _lib.fptr(out),
                                              x.shape[0], _lib.stream_ptr(x.device)))
        return out

    def unet_forward_uniform(self, x, t):
        """The forward as the fused loop runs it: one timestep for the whole batch (make_timesteps,
        diffusion_model_base.py:25-27), the loop's tensor-core policy, one cluster-kernel launch when supported."""
        if not x.is_cuda:
            raise RuntimeError("mpd_public_b200 runs on CUDA tensors only (no CPU fallback)")
        x, _ = self._prep(x)
        self.sync_params()
        out = torch.empty_like(x)
        _lib.check(self.lib.mpdb_unet_forward_uniform(self.handle, _lib.fptr(x), int(t), _lib.fptr(out), x.shape[0],
                                                      _lib.stream_ptr(x.device)))
        return out

    def mega_info(self, B):
        """(in_use, samples per cluster, layers, A-buffer bytes, shared-memory bytes, reason if not in use)"""
        g, n, a, sm = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        why = C.create_string_buffer(256)
        ok = self.lib.mpdb_engine_mega_info(self.handle, int(B), C.byref(g), C.byref(n), C.byref(a), C.byref(sm), why, 256)
        return bool(ok), g.value, n.value, a.value, sm.value, why.value.decode()

    def p_mean(self, x, t):
        x, t = self._prep(x, t)
        self.sync_params()
        out = torch.empty_like(x)
        _lib.check(self.lib.mpdb_p_mean(self.handle, _lib.fptr(x), C.c_void_p(t.data_ptr()), _lib.fptr(out),
                                        x.shape[0], _lib.stream_ptr(x.device)))
        return out

    def add_noise_(self, x, t, noise, noise_std):
        x2, t = self._prep(x, t)
        assert x2.data_ptr() == x.data_ptr(), "add_noise_ needs a contiguous fp32 tensor"
        noise = noise.to(torch.float32).contiguous()
        _lib.check(self.lib.mpdb_add_noise(self.handle, _lib.fptr(x), C.c_void_p(t.data_ptr()), _lib.fptr(noise),
                                           float(noise_std), x.shape[0], _lib.stream_ptr(x.device)))
        return x

    def sample_loop(self, noise, hard_conds, guide_handle, n_extra, t_start_guide, n_guide_steps, scale_grad_by_std,
                    noise_std_list, return_chain, use_cuda_graph):
        """noise: [n_iters+1, B, H, D]. Returns (x [B,H,D], chain [n_iters+1, B, H, D] or None)."""
        self.sync_params()
        n_iters = self.n_steps + n_extra
        S, B, H, D = noise.shape
        assert S == n_iters + 1
        if H != self.horizon or D != self.unet.state_dim:
            raise RuntimeError(f"this engine was built for trajectories [B, {self.horizon}, {self.unet.state_dim}], got "
                               f"[{B}, {H}, {D}]")
        p = _lib.LoopParams()
        p.horizon, p.state_dim = H, D
        p.n_steps_without_noise = n_extra
        p.t_start_guide = _lib.INT32_MAX if t_start_guide == float("inf") or t_start_guide >= _lib.INT32_MAX else int(
            np.ceil(t_start_guide))
        p.n_guide_steps = int(n_guide_steps)
        p.scale_grad_by_std = int(bool(scale_grad_by_std))
        ns = (C.c_float * n_iters)(*[float(v) for v in noise_std_list])
        p.noise_std = C.cast(ns, C.POINTER(C.c_float))
        rows = list(hard_conds.keys())  # dict order = overwrite order of apply_hard_conditioning
        if len(rows) > _lib.MAX_HARD_CONDS:
            raise RuntimeError(f"at most {_lib.MAX_HARD_CONDS} hard conditions are supported")
        p.n_hard_conds = len(rows)
        hc = None
        if rows:
            for k, r in enumerate(rows):
                p.hard_cond_rows[k] = int(r) % H
            hc = torch.stack([hard_conds[r].to(device=noise.device, dtype=torch.float32).expand(B, D) for r in rows]).contiguous()
            p.hard_cond_vals = hc.data_ptr()
        p.use_cuda_graph = int(bool(use_cuda_graph))
        x_out = torch.empty((B, H, D), device=noise.device, dtype=torch.float32)
        chain = torch.empty((S, B, H, D), device=noise.device, dtype=torch.float32) if return_chain else None
        _lib.check(self.lib.mpdb_sample_loop(
            self.handle, guide_handle, C.byref(p), _lib.fptr(noise), _lib.fptr(x_out),
            _lib.fptr(chain) if chain is not None else None, B * H * D, H * D, B, _lib.stream_ptr(noise.device)))
        # keep the host arrays referenced by the call alive until it returned (it is synchronous on the host)
        del ns
        return x_out, chain

    def ddim_loop(self, x_init, hard_conds, guide_handle, times, times_next, sqrt_alpha_next, coef_noise, t_start_guide,
                  n_guide_steps, return_chain):
        """x_init: [B, H, D] (randn, hard conditions not yet applied). Returns (x [B,H,D], chain [n+1, B, H, D] or None)."""
        self.sync_params()
        x_init, _ = self._prep(x_init)
        B, H, D = x_init.shape
        n = len(times)
        p = _lib.DdimParams()
        p.n_steps = n
        t_arr = (C.c_int32 * n)(*[int(v) for v in times])
        tn_arr = (C.c_int32 * n)(*[int(v) for v in times_next])
        sa_arr = (C.c_float * n)(*[float(v) for v in sqrt_alpha_next])
        cn_arr = (C.c_float * n)(*[float(v) for v in coef_noise])
        p.times, p.times_next = C.cast(t_arr, C.POINTER(C.c_int32)), C.cast(tn_arr, C.POINTER(C.c_int32))
        p.sqrt_alpha_next, p.coef_noise = C.cast(sa_arr, C.POINTER(C.c_float)), C.cast(cn_arr, C.POINTER(C.c_float))
        p.t_start_guide = _lib.INT32_MAX if t_start_guide == float("inf") or t_start_guide >= _lib.INT32_MAX else int(
            np.ceil(t_start_guide))
        p.n_guide_steps = int(n_guide_steps)
        rows = list(hard_conds.keys())
        if len(rows) > _lib.MAX_HARD_CONDS:
            raise RuntimeError(f"at most {_lib.MAX_HARD_CONDS} hard conditions are supported")
        p.n_hard_conds = len(rows)
        hc = None
        if rows:
            for k, r in enumerate(rows):
                p.hard_cond_rows[k] = int(r) % H
            hc = torch.stack([hard_conds[r].to(device=x_init.device, dtype=torch.float32).expand(B, D) for r in rows]).contiguous()
            p.hard_cond_vals = hc.data_ptr()
        p.horizon, p.state_dim = H, D
        x_out = torch.empty_like(x_init)
        chain = torch.empty((n + 1, B, H, D), device=x_init.device, dtype=torch.float32) if return_chain else None
        _lib.check(self.lib.mpdb_ddim_loop(
            self.handle, guide_handle, C.byref(p), _lib.fptr(x_init), _lib.fptr(x_out),
            _lib.fptr(chain) if chain is not None else None, B * H * D, H * D, B, _lib.stream_ptr(x_init.device)))
        del t_arr, tn_arr, sa_arr, cn_arr
        return x_out, chain

    def set_option(self, name, value):
        _lib.check(self.lib.mpdb_engine_set_option(self.handle, name.encode(), float(value)))

    def read_buffers(self, B):
        """name -> [B, C, L] activation of the last forward (parity debugging; needs option alias_buffers = 0
        set before that forward)."""
        out = {}
        n = self.lib.mpdb_engine_num_buffers(self.handle)
        for i in range(n):
            name = C.create_string_buffer(128)
            c, l = C.c_int32(), C.c_int32()
            _lib.check(self.lib.mpdb_engine_buffer_info(self.handle, i, name, 128, C.byref(c), C.byref(l)))
            t = torch.empty((B, c.value, l.value), device=self.device, dtype=torch.float32)
            _lib.check(self.lib.mpdb_engine_read_buffer(self.handle, i, _lib.fptr(t), B, _lib.stream_ptr(self.device)))
            out[name.value.decode()] = t
        return out


def _engine_for_unet(unet, device, n_steps=None, schedule=None, predict_epsilon=True, clip_denoised=True, horizon=None):
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("mpd_public_b200 has no CPU path: move the model and its inputs to a CUDA device")
    idx = device.index if device.index is not None else torch.cuda.current_device()
    horizon = int(horizon if horizon is not None else unet.n_support_points)
    cache = unet.__dict__.setdefault("_mpdb_engines", {})
    if n_steps is None:
        # stand-alone UNet call: reuse any engine of this device and horizon, else a table of 1000 time steps
        for (d, _t, _pe, _cd, h), eng in cache.items():
            if d == idx and h == horizon:
                return eng
        n_steps = unet.__dict__.get("_mpdb_default_steps", 1000)
    key = (idx, int(n_steps), bool(predict_epsilon), bool(clip_denoised), horizon)
    eng = cache.get(key)
    if eng is None:
        eng = Engine(unet, torch.device("cuda", idx), n_steps, schedule, predict_epsilon, clip_denoised, horizon)
        cache[key] = eng
    return eng


# ------------------------------------------------------------------------------------------------
class GaussianDiffusionModel(nn.Module):
    """reference diffusion_model_base.py:46 — same constructor signature; extra kwargs are ignored
    (inference.py:138-144 passes the UNet kwargs here as well)."""

    def __init__(self, model=None, variance_schedule='exponential', n_diffusion_steps=100, clip_denoised=True,
                 predict_epsilon=False, loss_type='l2', context_model=None, **kwargs):
        super().__init__()
        self.model = model
        self.context_model = context_model
        self.n_diffusion_steps = n_diffusion_steps
        self.state_dim = self.model.state_dim

        if variance_schedule == 'cosine':
            betas = cosine_beta_schedule(n_diffusion_steps, s=0.008, a_min=0, a_max=0.999)
        elif variance_schedule == 'exponential':
            betas = exponential_beta_schedule(n_diffusion_steps, beta_start=1e-4, beta_end=1.0)
        else:
            raise NotImplementedError

        alphas = 1. - betas
        alphas_cumprod = torch.cumprod(alphas, axis=0)
        alphas_cumprod_prev = torch.cat([torch.ones(1), alphas_cumprod[:-1]])
        self.clip_denoised = clip_denoised
        self.predict_epsilon = predict_epsilon

        self.register_buffer('betas', betas)
        self.register_buffer('alphas_cumprod', alphas_cumprod)
        self.register_buffer('alphas_cumprod_prev', alphas_cumprod_prev)
        self.register_buffer('sqrt_alphas_cumprod', torch.sqrt(alphas_cumprod))
        self.register_buffer('sqrt_one_minus_alphas_cumprod', torch.sqrt(1. - alphas_cumprod))
        self.register_buffer('log_one_minus_alphas_cumprod', torch.log(1. - alphas_cumprod))
        self.register_buffer('sqrt_recip_alphas_cumprod', torch.sqrt(1. / alphas_cumprod))
        self.register_buffer('sqrt_recipm1_alphas_cumprod', torch.sqrt(1. / alphas_cumprod - 1))
        posterior_variance = betas * (1. - alphas_cumprod_prev) / (1. - alphas_cumprod)
        self.register_buffer('posterior_variance', posterior_variance)
        self.register_buffer('posterior_log_variance_clipped', torch.log(torch.clamp(posterior_variance, min=1e-20)))
        self.register_buffer('posterior_mean_coef1', betas * np.sqrt(alphas_cumprod_prev) / (1. - alphas_cumprod))
        self.register_buffer('posterior_mean_coef2', (1. - alphas_cumprod_prev) * np.sqrt(alphas) / (1. - alphas_cumprod))
        self.use_cuda_graph = True
        self.tensor_cores = "auto"  # 'auto' | 'force' | 'off'  (see Engine.set_tensor_cores)
        if model is not None:
            model.__dict__["_mpdb_default_steps"] = n_diffusion_steps

    # ------------------------------------------ engine ------------------------------------------#
    def _schedule_sig(self):
        b = self.posterior_mean_coef1
        return (b.data_ptr(), b._version, self.sqrt_recip_alphas_cumprod._version)

    def _engine(self, horizon=None):
        device = self.betas.device
        if device.type != "cuda":
            raise RuntimeError("mpd_public_b200 has no CPU path: call .to('cuda') on the model first")
        sched = {k: getattr(self, k) for k in (
            'sqrt_recip_alphas_cumprod', 'sqrt_recipm1_alphas_cumprod', 'posterior_mean_coef1', 'posterior_mean_coef2',
            'posterior_log_variance_clipped')}
        eng = _engine_for_unet(self.model, device, self.n_diffusion_steps, sched, self.predict_epsilon, self.clip_denoised,
                               horizon)
        sig = self._schedule_sig()
        if eng.__dict__.get("_sched_sig") != sig:  # buffers reloaded by load_state_dict
            eng._set_schedule(sched)
            eng._sched_sig = sig
        eng.set_tensor_cores(self.tensor_cores)
        return eng

    def _device_rng(self, device):
        """csrc/rng.cu generator for `device` (None: `self.device_rng = False`, or a torch without generator offsets)."""
        if not self.__dict__.get("device_rng", True) or not _DeviceNormal.supported(device):
            return None
        cache = self.__dict__.setdefault("_device_rngs", {})
        key = str(torch.device(device))
        if key not in cache:
            cache[key] = _DeviceNormal(device)
        return cache[key]

    # ------------------------------------------ sampling ------------------------------------------#
    def predict_noise_from_start(self, x_t, t, x0):
        if self.predict_epsilon:
            return x0
        return (extract(self.sqrt_recip_alphas_cumprod, t, x_t.shape) * x_t - x0) / \
            extract(self.sqrt_recipm1_alphas_cumprod, t, x_t.shape)

    def predict_start_from_noise(self, x_t, t, noise):
        if self.predict_epsilon:
            return (extract(self.sqrt_recip_alphas_cumprod, t, x_t.shape) * x_t -
                    extract(self.sqrt_recipm1_alphas_cumprod, t, x_t.shape) * noise)
        return noise

    def q_posterior(self, x_start, x_t, t):
        posterior_mean = (extract(self.posterior_mean_coef1, t, x_t.shape) * x_start +
                          extract(self.posterior_mean_coef2, t, x_t.shape) * x_t)
        posterior_variance = extract(self.posterior_variance, t, x_t.shape)
        posterior_log_variance_clipped = extract(self.posterior_log_variance_clipped, t, x_t.shape)
        return posterior_mean, posterior_variance, posterior_log_variance_clipped

    @torch.no_grad()
    def p_mean_variance(self, x, hard_conds, context, t):
        """reference :143-155. UNet + x0 reconstruction + clamp + posterior mean run in one fused kernel chain."""
        if context is not None:
            raise NotImplementedError("context conditioning is not on the guided-sampling path")
        model_mean = self._engine(x.shape[1]).p_mean(x, t)
        posterior_variance = extract(self.posterior_variance, t, x.shape)
        posterior_log_variance = extract(self.posterior_log_variance_clipped, t, x.shape)
        return model_mean, posterior_variance, posterior_log_variance

    @torch.no_grad()
    def p_sample_loop(self, shape, hard_conds, context=None, return_chain=False, sample_fn=ddpm_sample_fn,
                      n_diffusion_steps_without_noise=0, **sample_kwargs):
        """reference :158-182."""
        device = self.betas.device
        if device.type != "cuda":
            raise RuntimeError("mpd_public_b200 has no CPU path: call .to('cuda') on the model first")
        batch_size = shape[0]
        guide = sample_kwargs.get("guide", None)
        fused = (sample_fn is ddpm_sample_fn and context is None
                 and (guide is None or getattr(guide, "_mpdb_fusable", False)))
        if fused:
            return self._p_sample_loop_fused(shape, hard_conds, return_chain, n_diffusion_steps_without_noise,
                                             **sample_kwargs)

        # generic path: any sample_fn / any guide callable, one step at a time (still CUDA kernels underneath)
        x = torch.randn(shape, device=device)
        x = apply_hard_conditioning(x, hard_conds)
        chain = [x] if return_chain else None
        for i in reversed(range(-n_diffusion_steps_without_noise, self.n_diffusion_steps)):
            t = make_timesteps(batch_size, i, device)
            x, values = sample_fn(self, x, hard_conds, context, t, **sample_kwargs)
            x = apply_hard_conditioning(x, hard_conds)
            if return_chain:
                chain.append(x)
        if return_chain:
            chain = torch.stack(chain, dim=1)
            return x, chain
        return x

    def _p_sample_loop_fused(self, shape, hard_conds, return_chain, n_extra, guide=None, n_guide_steps=1,
                             scale_grad_by_std=False, t_start_guide=torch.inf, noise_std_extra_schedule_fn=None,
                             debug=False, noise=None, **kwargs):
        device = self.betas.device
        if len(shape) != 3 or shape[2] != self.state_dim:
            raise RuntimeError(f"expected shape (batch, horizon, {self.state_dim}), got {tuple(shape)}")
        eng = self._engine(shape[1])
        steps = list(reversed(range(-n_extra, self.n_diffusion_steps)))
        if (noise is None and self.use_cuda_graph and self.__dict__.get("graph_rng", True)
                and not torch.cuda.is_current_stream_capturing()):
            out = self._graphed_loop(eng, shape, hard_conds, return_chain, n_extra, steps, guide, n_guide_steps,
                                     scale_grad_by_std, t_start_guide, noise_std_extra_schedule_fn)
            if out is not None:
                return out
        if noise is None:
            # identical generator consumption to the reference: randn(shape), then one randn_like per step (randn is
            # empty + normal_). The buffer and its per-step views are kept between calls: allocating and slicing them
            # is most of the host time of a call, and the loop copies the noise into its own staging anyway.
            key = (len(steps) + 1, tuple(shape), str(device))
            cache = self.__dict__.get("_noise_cache")
            if cache is None or cache[0] != key:
                buf = torch.empty((len(steps) + 1, *shape), device=device, dtype=torch.float32)
                cache = (key, buf, list(buf.unbind(0)))
                self.__dict__["_noise_cache"] = cache
            noise = cache[1]
            rng = self._device_rng(device)
            if rng is not None:
                rng.advance(noise[0].numel(), noise.shape[0])
                rng.fill(noise)
            else:
                for view in cache[2]:
                    view.normal_()
        else:
            noise = noise.to(device=device, dtype=torch.float32).contiguous()
            if tuple(noise.shape) != (len(steps) + 1, *shape):
                raise RuntimeError(f"injected noise must have shape {(len(steps) + 1, *shape)}")
        if noise_std_extra_schedule_fn is None:
            ns = [1.0] * len(steps)
        else:
            # the reference calls the schedule with the step index as a tensor; the index tensors are cached
            tcache = self.__dict__.setdefault("_step_tensors", {})
            ns = []
            for i in steps:
                ti = tcache.get(i)
                if ti is None:
                    ti = tcache[i] = torch.tensor(i, dtype=torch.long)
                ns.append(float(noise_std_extra_schedule_fn(ti)))
        tsg = float(t_start_guide)
        handle = guide._handle(device, shape[1]) if guide is not None else None
        x, chain = eng.sample_loop(noise, hard_conds, handle, n_extra, tsg, n_guide_steps if guide is not None else 0,
                                   scale_grad_by_std, ns, return_chain, self.use_cuda_graph)
        if return_chain:
            return x, chain.transpose(0, 1)  # [B, steps+1, H, D] like torch.stack(chain, dim=1)
        return x

    def _graphed_loop(self, eng, shape, hard_conds, return_chain, n_extra, steps, guide, n_guide_steps, scale_grad_by_std,
                      t_start_guide, noise_std_extra_schedule_fn):
        """One CUDA-graph launch per call for the API path that draws its own noise: the per-step `normal_()` draws (the
        reference's generator consumption, graph-safe Philox offsets) AND the whole reverse loop are captured together, so
        a call costs one launch on the host instead of 31 RNG launches + the loop. Start/goal are copied into static
        buffers before the replay; the result is cloned out of the graph's memory. Returns None if capture is impossible."""
        device = self.betas.device
        if noise_std_extra_schedule_fn is None:
            ns = (1.0,) * len(steps)
        else:
            tcache = self.__dict__.setdefault("_step_tensors", {})
            ns = []
            for i in steps:
                ti = tcache.get(i)
                if ti is None:
                    ti = tcache[i] = torch.tensor(i, dtype=torch.long)
                ns.append(float(noise_std_extra_schedule_fn(ti)))
            ns = tuple(ns)
        handle = guide._handle(device, shape[1]) if guide is not None else None
        rows = tuple(hard_conds.keys())
        eng.sync_params()
        gen = int(eng.lib.mpdb_engine_generation(eng.handle))
        key = (tuple(shape), n_extra, handle.value if handle is not None else None, id(guide), int(n_guide_steps),
               bool(scale_grad_by_std), float(t_start_guide), ns, bool(return_chain), rows, str(device), gen,
               torch.cuda.current_device())
        cache = self.__dict__.setdefault("_loop_graphs", {})
        entry = cache.get(key)
        if entry is None:
            if len(cache) > 32:
                cache.clear()
            try:
                noise = torch.empty((len(steps) + 1, *shape), device=device, dtype=torch.float32)
                views = list(noise.unbind(0))
                static_hc = {r: torch.empty((shape[0], shape[2]), device=device, dtype=torch.float32) for r in rows}

                rng = self._device_rng(device)

                def run_once():
                    if rng is not None:
                        rng.fill(noise)  # one launch; (seed, offset) are read from rng.state at replay time
                    else:
                        for v in views:
                            v.normal_()
                    return eng.sample_loop(noise, static_hc, handle, n_extra, float(t_start_guide),
                                           n_guide_steps if guide is not None else 0, scale_grad_by_std, list(ns),
                                           return_chain, False)

                for r in rows:
                    static_hc[r].copy_(hard_conds[r].to(device=device, dtype=torch.float32).expand_as(static_hc[r]))
                rng_state = torch.cuda.get_rng_state(device)
                side = torch.cuda.Stream(device=device)
                side.wait_stream(torch.cuda.current_stream(device))
                with torch.cuda.stream(side):
                    run_once()  # warm-up: engine workspace, lazily configured kernels
                torch.cuda.current_stream(device).wait_stream(side)
                torch.cuda.synchronize(device)
                torch.cuda.set_rng_state(rng_state, device)  # the warm-up must not consume the caller's generator
                if int(eng.lib.mpdb_engine_generation(eng.handle)) != gen:  # the warm-up (re)allocated: re-key
                    gen = int(eng.lib.mpdb_engine_generation(eng.handle))
                    key = key[:11] + (gen,) + key[12:]
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    x_out, chain = run_once()
                entry = (graph, static_hc, x_out, chain, rng, noise)
                cache[key] = entry
            except Exception as exc:  # capture not possible in this context: the plain path still works
                self.__dict__["graph_rng"] = False
                self.__dict__["_graph_rng_error"] = repr(exc)
                return None
        graph, static_hc, x_out, chain, rng, noise_buf = entry
        if rng is not None:  # what the reference's eager draws would consume; torch's own graph-safe offsets otherwise
            rng.advance(noise_buf[0].numel(), noise_buf.shape[0])
        for r in rows:
            v = hard_conds[r]
            if v.device != device or v.dtype != torch.float32:
                v = v.to(device=device, dtype=torch.float32)
            static_hc[r].copy_(v)  # broadcasts [D] / [1, D] / [B, D]
        graph.replay()
        if return_chain:
            return x_out.clone(), chain.transpose(0, 1).clone()
        return x_out.clone()

    def ddim_schedule(self):
        """The (time, time_next) pairs of `ddim_sample` and the two coefficients of each update, computed with the torch
        expressions the reference evaluates per step (diffusion_model_base.py:196-209, :232-236; eta = 0) so that the fp32
        values the device multiplies by are bit-identical: (times, times_next, sqrt(alpha_next), sqrt(1 - alpha_next - sigma^2))."""
        total_timesteps, sampling_timesteps, eta = self.n_diffusion_steps, self.n_diffusion_steps // 5, 0.
        times = torch.linspace(0, total_timesteps - 1, steps=sampling_timesteps + 1)
        times = torch.cat((torch.tensor([-1]), times))
        times = list(reversed(times.int().tolist()))
        ac = self.alphas_cumprod.detach().cpu()
        t_l, tn_l, sa_l, cn_l = [], [], [], []
        for time, time_next in zip(times[:-1], times[1:]):
            t_l.append(time)
            tn_l.append(time_next)
            if time_next < 0:
                sa_l.append(1.0)
                cn_l.append(0.0)
                break
            alpha, alpha_next = ac[time], ac[time_next]
            sigma = eta * ((1 - alpha / alpha_next) * (1 - alpha_next) / (1 - alpha)).sqrt()
            sa_l.append(float(alpha_next.sqrt()))
            cn_l.append(float((1 - alpha_next - sigma ** 2).sqrt()))
        return t_l, tn_l, sa_l, cn_l

    @torch.no_grad()
    def ddim_sample(self, shape, hard_conds, context=None, return_chain=False, t_start_guide=torch.inf, guide=None,
                    n_guide_steps=1, noise=None, **sample_kwargs):
        """reference :184-259 (T // 5 steps, eta = 0) as ONE C-ABI call (`mpdb_ddim_loop`): per time pair a UNet forward whose
        last epilogue forms x_start / pred_noise and the DDIM update, then the guide evaluations when time_next <
        t_start_guide, then the hard conditions; no host synchronisation inside.

        As in the reference, the named argument `n_guide_steps` is NOT what the guide receives: `guide_gradient_steps` is
        called with `**sample_kwargs` only (:240-245), i.e. one step unless the caller passes n_guide_steps there — which
        Python routes to the named argument, so the reference always runs a single guide step per DDIM step; so does this.
        The generator is consumed like the reference's: randn(shape), then one randn_like per step (multiplied by sigma = 0).
        `noise` (optional, [B, H, D]) replaces the initial draw (parity tests)."""
        device = self.betas.device
        if device.type != "cuda":
            raise RuntimeError("mpd_public_b200 has no CPU path: call .to('cuda') on the model first")
        if context is not None:
            raise NotImplementedError("context conditioning is not on the guided-sampling path")
        if len(shape) != 3 or shape[2] != self.state_dim:
            raise RuntimeError(f"expected shape (batch, horizon, {self.state_dim}), got {tuple(shape)}")
        if guide is not None and not getattr(guide, "_mpdb_fusable", False):
            raise NotImplementedError("ddim_sample runs with this package's GuideManagerTrajectoriesWithVelocity (or guide=None)")
        if sample_kwargs.get("scale_grad_by_std", False):
            raise NotImplementedError("ddim_sample passes no model_var to guide_gradient_steps (the reference would fail too)")
        eng = self._engine(shape[1])
        t_l, tn_l, sa_l, cn_l = self.ddim_schedule()
        if noise is None:
            x = torch.randn(shape, device=device)
            scratch = torch.empty_like(x)
            for tn in tn_l:
                if tn >= 0:
                    scratch.normal_()  # the reference's per-step randn_like (times sigma = 0)
        else:
            x = noise.to(device=device, dtype=torch.float32)
            if tuple(x.shape) != tuple(shape):
                raise RuntimeError(f"injected noise must have shape {tuple(shape)}")
        handle = guide._handle(device, shape[1]) if guide is not None else None
        x, chain = eng.ddim_loop(x, hard_conds, handle, t_l, tn_l, sa_l, cn_l, float(t_start_guide),
                                 1 if guide is not None else 0, return_chain)
        if return_chain:
            return x, chain.transpose(0, 1)  # [B, steps + 1, H, D] like torch.stack(chain, dim=1)
        return x

    @torch.no_grad()
    def conditional_sample(self, hard_conds, horizon=None, batch_size=1, ddim=False, **sample_kwargs):
        """reference :262-272"""
        horizon = horizon or self.model.n_support_points
        shape = (batch_size, horizon, self.state_dim)
        if ddim:
            return self.ddim_sample(shape, hard_conds, **sample_kwargs)
        return self.p_sample_loop(shape, hard_conds, **sample_kwargs)

    def forward(self, cond, *args, **kwargs):
        raise NotImplementedError  # as in the reference (:274-276)

    @torch.no_grad()
    def warmup(self, horizon=64, device='cuda'):
        """reference :279-283"""
        shape = (2, horizon, self.state_dim)
        x = torch.randn(shape, device=device)
        t = make_timesteps(2, 1, device)
        self.model(x, t, context=None)

    @torch.no_grad()
    def run_inference(self, context=None, hard_conds=None, n_samples=1, return_chain=False, **diffusion_kwargs):
        """reference :286-316 — returns [steps+1, n_samples, H, D] if return_chain else [n_samples, H, D]."""
        hard_conds = copy(hard_conds)
        context = copy(context)
        for k, v in hard_conds.items():
            # einops.repeat(v, 'd -> b d', b=n_samples): the loop only reads the conditions (and copies them into its own
            # staging), so the broadcast view stands in for the materialised copy (one launch + one allocation less per row)
            hard_conds[k] = v.unsqueeze(0).expand(n_samples, -1)
        if context is not None:
            for k, v in context.items():
                context[k] = v.unsqueeze(0).repeat(n_samples, 1)
        if not return_chain:
            # the reference keeps the whole chain and returns its last entry; the last entry is the sample itself
            return self.conditional_sample(hard_conds, context=context, batch_size=n_samples, return_chain=False,
                                           **diffusion_kwargs)
        samples, chain = self.conditional_sample(hard_conds, context=context, batch_size=n_samples, return_chain=True,
                                                 **diffusion_kwargs)
        trajs_chain_normalized = chain.permute(1, 0, 2, 3)  # 'b diffsteps h d -> diffsteps b h d'
        if return_chain:
            return trajs_chain_normalized
        return trajs_chain_normalized[-1]

    @torch.no_grad()
    def sample(self, hard_conds, n_samples, horizon=None, noise=None, **diffusion_kwargs):
        """Throughput-oriented entry (not in the reference): final plans only, no chain kept."""
        hard_conds = {k: v.unsqueeze(0).repeat(n_samples, 1) for k, v in hard_conds.items()}
        horizon = horizon or self.model.n_support_points
        return self.p_sample_loop((n_samples, horizon, self.state_dim), hard_conds, return_chain=False, noise=noise,
                                  **diffusion_kwargs)
