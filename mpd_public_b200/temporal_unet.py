"""TemporalUnet — host-side mirror of reference `mpd/models/diffusion_models/temporal_unet.py:20-171`.

Same constructor arguments, same `state_dict()` keys and shapes, so reference
checkpoints load with `load_state_dict`. The module holds parameters only; the arithmetic runs in
libmpdb200's CUDA kernels (csrc/unet.cu): Conv1d+GroupNorm+Mish(+time cond)(+residual) fused per
launch, channel-major activations with zero halos, time-conditioning hoisted into [T, C] tables.
Only the configuration the inference path uses is supported: conditioning_type=None,
self_attention=False.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib
from .synthetic import UNET_DIM_MULTS, unet_param_shapes  # noqa: F401  (re-exported)


class _Node(nn.Module):
    """Anonymous container so that parameter paths match the reference's module tree."""


def _register(root: nn.Module, dotted: str, tensor: torch.Tensor):
    parts = dotted.split(".")
    m = root
    for p in parts[:-1]:
        if p not in m._modules:
            m.add_module(p, _Node())
        m = m._modules[p]
    m.register_parameter(parts[-1], nn.Parameter(tensor))


class TemporalUnet(nn.Module):

    def __init__(self, n_support_points=None, state_dim=None, unet_input_dim=32, dim_mults=(1, 2, 4, 8),
                 time_emb_dim=32, self_attention=False, conditioning_embed_dim=4, conditioning_type=None,
                 attention_num_heads=2, attention_dim_head=32, **kwargs):
        super().__init__()
        if conditioning_type not in (None, "None"):
            raise NotImplementedError("only conditioning_type=None is on the guided-sampling path (SURVEY §2 row 4)")
        if self_attention:
            raise NotImplementedError("self_attention=True is not used by the shipped configurations")
        if time_emb_dim != 32:
            raise NotImplementedError("time_emb_dim is fixed to 32 (reference temporal_unet.py:66)")
        if n_support_points is None or state_dim is None:
            raise ValueError("n_support_points and state_dim are required")
        self.state_dim = int(state_dim)
        self.n_support_points = int(n_support_points)
        self.unet_input_dim = int(unet_input_dim)
        self.dim_mults = tuple(int(m) for m in dim_mults)
        self.conditioning_type = None
        dims = [self.state_dim] + [self.unet_input_dim * m for m in self.dim_mults]
        in_out = list(zip(dims[:-1], dims[1:]))
        print(f'[ models/temporal ] Channel dimensions: {in_out}')
        # default-initialised like nn.Conv1d/nn.Linear/nn.GroupNorm would be (uniform +-1/sqrt(fan_in); GN = 1, 0)
        # registration order = the reference's state_dict() order: its __init__ creates the (still empty) `downs` and `ups`
        # ModuleLists before the mid blocks (temporal_unet.py:60-116), so `ups.*` precedes `mid_block*` in a saved checkpoint
        shapes_all = unet_param_shapes(self.state_dim, self.unet_input_dim, self.dim_mults)
        rank = {"time_mlp": 0, "downs": 1, "ups": 2, "mid_block1": 3, "mid_block2": 4, "final_conv": 5}
        for name, shape in sorted(shapes_all.items(), key=lambda kv: rank[kv[0].split(".")[0]]):
            if ".block.2." in name:
                t = torch.ones(shape) if name.endswith("weight") else torch.zeros(shape)
            else:
                wshape = shape if name.endswith("weight") else \
                    unet_param_shapes(self.state_dim, self.unet_input_dim, self.dim_mults)[name[:-4] + "weight"]
                fan_in = 1
                for s in wshape[1:]:
                    fan_in *= s
                bound = 1.0 / fan_in ** 0.5
                t = torch.empty(shape).uniform_(-bound, bound)
            _register(self, name, t)

    def forward(self, x, time, context=None):
        """x: [batch, horizon, state_dim] (CUDA fp32), time: [batch] integer -> eps [batch, horizon, state_dim]."""
        from .diffusion_model import _engine_for_unet
        _lib.require_cuda(x, "x")
        if context is not None:
            raise NotImplementedError("context conditioning is not on the guided-sampling path")
        if x.dim() != 3:
            raise RuntimeError(f"expected x of shape [batch, horizon, {self.state_dim}], got {tuple(x.shape)}")
        eng = _engine_for_unet(self, x.device, horizon=x.shape[1])  # fully convolutional: one device plan per horizon
        return eng.unet_forward(x, time)
