"""Deterministic parameter values by NAME, so that fixtures need not store multi-megabyte weights.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  ``fill_by_name(module, seed)`` overwrites every
parameter ``p`` named ``k`` in ``module.named_parameters()`` with uniform(-a, a) values drawn from a CPU
``torch.Generator`` seeded with ``seed + crc32(k)``; ``a = scale / sqrt(fan_in)`` (``fan_in`` = last dim, or
the numel of a vector).  The reference module (in oracle/make_golden.py) and the drop-in module (in the
tests) have the same parameter names, hence receive the same values without exchanging them.  The CPU
generator stream is a function of the torch build only; the GPU box runs the same image.

``subsample`` / ``digest`` reduce a large gradient to a strided sample (<= ``n`` elements) + its max-abs, which
is what the MMGCN / DAG-ERC fixtures store for their 64 x [400,200] and 4 x 1.2 M-float weight gradients.
"""
import math
import zlib

import numpy as np
import torch


def fill_by_name(module, seed, scale=1.0, only=None):
    with torch.no_grad():
        for k, p in module.named_parameters():
            if only is not None and not only(k):
                continue
            g = torch.Generator().manual_seed((int(seed) + zlib.crc32(k.encode())) % (2 ** 31))
            fan_in = p.shape[-1] if p.dim() > 1 else max(p.numel(), 1)
            a = scale / math.sqrt(fan_in)
            p.copy_((torch.rand(p.shape, generator=g) * 2 - 1) * a)


def subsample(arr, n=2048):
    flat = np.asarray(arr).reshape(-1)
    step = max(1, flat.size // n)
    return flat[::step][:n].copy()


def digest(arr, n=2048):
    """(strided sample, max-abs, sum) of a tensor."""
    a = np.asarray(arr, dtype=np.float32)
    return subsample(a, n), np.float32(np.abs(a).max() if a.size else 0.0), np.float64(a.astype(np.float64).sum())
