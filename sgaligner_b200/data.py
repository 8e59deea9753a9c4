"""Host <-> device movement of the collated Scan3R batch dict (mirrors ``utils/torch_util.py:26-36``:
only torch tensors move, numpy arrays and ints stay on the host)."""
from __future__ import annotations

import numpy as np
import torch


def to_cuda(x, device=None, non_blocking: bool = True):
    if isinstance(x, list):
        return [to_cuda(i, device, non_blocking) for i in x]
    if isinstance(x, tuple):
        return tuple(to_cuda(i, device, non_blocking) for i in x)
    if isinstance(x, dict):
        return {k: to_cuda(v, device, non_blocking) for k, v in x.items()}
    if torch.is_tensor(x):
        return x.cuda(device, non_blocking=non_blocking) if device is None or isinstance(device, int) else x.to(device, non_blocking=non_blocking)
    return x


def pin(x):
    """Pinned-host copy of every tensor in a batch dict (so that to_cuda is a true async H2D copy)."""
    if isinstance(x, dict):
        return {k: pin(v) for k, v in x.items()}
    if torch.is_tensor(x) and not x.is_cuda:
        return x.pin_memory()
    return x


def h2d_bytes(data: dict) -> int:
    n = 0
    for v in data.values():
        if torch.is_tensor(v):
            n += v.numel() * v.element_size()
    for k in ('e1i', 'e2i', 'e1j', 'e2j'):
        if k in data:
            n += np.asarray(data[k]).size * 4
    return int(n)
