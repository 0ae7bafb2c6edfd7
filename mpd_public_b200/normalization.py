"""LimitsNormalizer / DatasetNormalizer — mirror of reference `mpd/datasets/normalization.py:12-167`.

Used outside the timed loop (building hard conditions, unnormalising the final plans). Inside the
loop the unnormalisation — including the batch-global clip branch of `unnormalize` (:160-162) — is
fused into the guide kernel (csrc/guide.cu).
"""
from __future__ import annotations

import torch


class Normalizer:
    def __init__(self, X):
        self.X = X
        self.mins = X.min(dim=0).values
        self.maxs = X.max(dim=0).values

    def __call__(self, x):
        return self.normalize(x)

    def normalize(self, *args, **kwargs):
        raise NotImplementedError()

    def unnormalize(self, *args, **kwargs):
        raise NotImplementedError()


class LimitsNormalizer(Normalizer):
    """maps [xmin, xmax] to [-1, 1] (reference :144-167)"""

    def _range_tensor(self):
        rng = self.__dict__.get("_range")  # maxs - mins, the same fp32 subtraction as the reference's, done once
        if rng is None or rng.device != self.mins.device:
            rng = self.__dict__["_range"] = self.maxs - self.mins
        return rng

    def normalize(self, x, pad_to=None):
        """`pad_to`: normalise `cat(x, zeros)` with `pad_to` columns in all (what `get_hard_conditions` needs) without
        materialising the concatenation. fp32 CUDA tensors take ONE launch of `mpdb_limits_normalize` (same operations, same
        roundings as the four eager kernels below: torch.equal, tests/test_gpu_parity.py); anything else runs the torch ops."""
        rng = self._range_tensor()
        d_out = x.shape[-1] if pad_to is None else int(pad_to)
        if (x.is_cuda and x.dtype == torch.float32 and self.mins.is_cuda and self.mins.device == x.device
                and self.mins.dtype == torch.float32 and self.mins.numel() == d_out and x.numel() > 0):
            from . import _lib
            xc = x.contiguous()
            out = torch.empty((*x.shape[:-1], d_out), dtype=torch.float32, device=x.device)
            _lib.check(_lib.lib().mpdb_limits_normalize(_lib.fptr(xc), xc.numel() // x.shape[-1], x.shape[-1],
                                                        _lib.fptr(self.mins.contiguous()), _lib.fptr(rng.contiguous()),
                                                        _lib.fptr(out), d_out, x.device.index, _lib.stream_ptr(x.device)))
            return out
        if pad_to is not None and d_out > x.shape[-1]:
            x = torch.cat((x, torch.zeros((*x.shape[:-1], d_out - x.shape[-1]), dtype=x.dtype, device=x.device)), dim=-1)
        x = (x - self.mins) / rng
        x = 2 * x - 1
        return x

    def unnormalize(self, x, eps=1e-4):
        if x.max() > 1 + eps or x.min() < -1 - eps:
            x = torch.clip(x, -1, 1)
        x = (x + 1) / 2.
        return x * (self.maxs - self.mins) + self.mins


class DatasetNormalizer:
    """reference :12-45, built directly from per-field limits instead of a loaded dataset."""

    def __init__(self, dataset, normalizer=LimitsNormalizer):
        if isinstance(normalizer, str):
            normalizer = {"LimitsNormalizer": LimitsNormalizer}[normalizer]
        self.normalizers = {key: normalizer(val.reshape(-1, val.shape[-1])) for key, val in dataset.items()}

    def __call__(self, *args, **kwargs):
        return self.normalize(*args, **kwargs)

    def normalize(self, x, key, **kwargs):
        return self.normalizers[key].normalize(x, **kwargs)

    def unnormalize(self, x, key):
        return self.normalizers[key].unnormalize(x)

    def get_field_normalizers(self):
        return self.normalizers
