"""Deterministic, platform-independent fill for parameters and synthetic inputs.

The golden generator (reference classes, build container) and the parity tests (our classes, any box)
both fill a module's state-dict with ``fill_state_dict``: every entry is a pure function of its KEY NAME
and SHAPE (splitmix64 counter stream -> uniform floats), so no weights have to be stored in the fixture,
and a key that exists on one side only — or with another shape — fails the test by construction.
"""
import zlib

import numpy as np
import torch

_MASK = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x):
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _MASK
        z = x
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _MASK
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _MASK
        return z ^ (z >> np.uint64(31))


def uniform(name, shape, lo=-1.0, hi=1.0):
    """float32 array of ``shape`` in [lo, hi): element i = f(crc32(name), i)."""
    n = int(np.prod(shape)) if len(shape) else 1
    seed = np.uint64(zlib.crc32(name.encode()) * 0x100000001B3 % (1 << 64))
    with np.errstate(over="ignore"):
        ctr = (np.arange(n, dtype=np.uint64) * np.uint64(0xD1342543DE82EF95) + seed) & _MASK
    bits = _splitmix64(ctr) >> np.uint64(40)                       # 24 random bits
    u = bits.astype(np.float64) / float(1 << 24)
    return (lo + (hi - lo) * u).astype(np.float32).reshape(shape)


def fill_state_dict(module, scale=1.0):
    """In-place deterministic values for every parameter and buffer of ``module``; returns the sorted keys.

    weights (>= 2-D): U(-a, a), a = scale * sqrt(3 / fan_in)   (variance-preserving)
    norm weights (1-D '...weight'): U(0.5, 1.5); running_var: U(0.5, 1.5); running_mean / biases / other 1-D:
    U(-0.3, 0.3); integer buffers (num_batches_tracked) untouched.
    """
    sd = module.state_dict()
    with torch.no_grad():
        for key in sorted(sd):
            t = sd[key]
            if not t.is_floating_point():
                continue
            shape = tuple(t.shape)
            if t.dim() >= 2:
                fan_in = int(np.prod(shape[1:]))
                a = scale * float(np.sqrt(3.0 / fan_in))
                v = uniform(key, shape, -a, a)
            elif key.endswith("running_var") or (key.endswith("weight") and t.dim() == 1):
                v = uniform(key, shape, 0.5, 1.5)
            else:
                v = uniform(key, shape, -0.3, 0.3)
            t.copy_(torch.from_numpy(v).to(t.device))
    return sorted(sd)


def state_dict_signature(module):
    return sorted((k, tuple(v.shape)) for k, v in module.state_dict().items())
