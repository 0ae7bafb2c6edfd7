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

    def normalize(self, x):
        rng = self.__dict__.get("_range")  # maxs - mins, the same fp32 subtraction as the reference's, done once
        if rng is None or rng.device != self.mins.device:
            rng = self.__dict__["_range"] = self.maxs - self.mins
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

    def normalize(self, x, key):
        return self.normalizers[key].normalize(x)

    def unnormalize(self, x, key):
        return self.normalizers[key].unnormalize(x)

    def get_field_normalizers(self):
        return self.normalizers
