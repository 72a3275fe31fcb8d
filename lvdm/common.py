"""Helpers with the reference's semantics (lvdm/common.py:25-44)."""
from inspect import isfunction

import torch


def exists(v):
    return v is not None


def default(v, d):
    if v is not None:
        return v
    return d() if isfunction(d) else d


def extract_into_tensor(a, t, x_shape):
    """a[t] broadcast to x_shape's rank: [b] -> [b,1,1,...] (common.py:25-28)."""
    return a.gather(-1, t).reshape(t.shape[0], *((1,) * (len(x_shape) - 1)))


def noise_like(shape, device, repeat=False):
    """Fresh Gaussian noise on `device` from the global generator (common.py:31-34)."""
    if repeat:
        return torch.randn((1, *shape[1:]), device=device).repeat(shape[0], *((1,) * (len(shape) - 1)))
    return torch.randn(shape, device=device)


def autocast(f, enabled=True):
    """Decorator with the semantics of the reference one (lvdm/common.py:16-22): run `f` under CUDA autocast with the ambient settings."""
    def do_autocast(*args, **kwargs):
        with torch.cuda.amp.autocast(enabled=enabled, dtype=torch.get_autocast_gpu_dtype(),
                                     cache_enabled=torch.is_autocast_cache_enabled()):
            return f(*args, **kwargs)
    return do_autocast
