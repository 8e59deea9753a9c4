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


# which collated tensors each modality of MultiModalEncoder.forward reads (sg_aligner.py:72-122)
_MODULE_INPUTS = {
    'point': ('tot_obj_pts',),
    'pct': ('tot_obj_pts',),
    'gat': ('tot_rel_pose', 'edges'),
    'rel': ('tot_bow_vec_object_edge_feats',),
    'attr': ('tot_bow_vec_object_attr_feats',),
}


def needed_keys(modules) -> set:
    """Tensor keys of the batch dict the given modalities consume.  The reference moves the whole dict to the
    device (``torch_util.to_cuda``); a serving loop only has to move what its encoder reads -- at C2
    (point + gat) the unused bag-of-words tensors are 20 % of the bytes."""
    keys = set()
    for m in modules:
        keys.update(_MODULE_INPUTS.get(m, ()))
    return keys


def h2d_bytes(data: dict, keys=None) -> int:
    n = 0
    for k, v in data.items():
        if torch.is_tensor(v) and (keys is None or k in keys):
            n += v.numel() * v.element_size()
    for k in ('e1i', 'e2i', 'e1j', 'e2j'):
        if k in data:
            n += np.asarray(data[k]).size * 4
    return int(n)


def to_cuda_streamed(data: dict, device, n_chunks: int = 4, copy_stream=None, keys=None) -> dict:
    """``to_cuda`` for serving: the copies run on a side stream so that they overlap the kernels of the
    same step.  The small tensors (edges, BoW, poses) go first, then ``tot_obj_pts`` in ``n_chunks``
    object ranges, each with its own event; ``MultiModalEncoder.forward`` waits for the small tensors,
    runs the graph branch, and launches the point encoder chunk by chunk as the copies land.
    Host tensors should be pinned (:func:`pin`), otherwise the copies are synchronous.
    ``keys``: only these tensors are copied (:func:`needed_keys`); the others stay on the host."""
    import torch
    dev = torch.device(device) if not isinstance(device, torch.device) else device
    cs = copy_stream if copy_stream is not None else _copy_stream(dev)
    cur = torch.cuda.current_stream(dev)
    cs.wait_stream(cur)          # do not overwrite buffers a previous step may still be reading
    out = {}
    with torch.cuda.stream(cs):
        for k, v in data.items():
            if torch.is_tensor(v) and k != 'tot_obj_pts':
                out[k] = v.to(dev, non_blocking=True) if (keys is None or k in keys) else v
            elif not torch.is_tensor(v):
                out[k] = v
        ev_small = torch.cuda.Event()
        ev_small.record(cs)
        pts = data['tot_obj_pts']
        N = pts.shape[0]
        dpts = torch.empty(pts.shape, dtype=pts.dtype, device=dev)
        chunks = []
        per = (N + n_chunks - 1) // n_chunks
        for s in range(0, N, per):
            e = min(N, s + per)
            dpts[s:e].copy_(pts[s:e], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(cs)
            chunks.append((s, e, ev))
    out['tot_obj_pts'] = dpts
    for v in out.values():       # allocated on the copy stream, consumed on the compute stream
        if torch.is_tensor(v) and v.is_cuda:
            v.record_stream(cur)
    out['_sga_ready'] = {'small': ev_small, 'pts': chunks, 'stream': cs}
    return out


_COPY_STREAMS = {}


def _copy_stream(dev):
    import torch
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    if key not in _COPY_STREAMS:
        _COPY_STREAMS[key] = torch.cuda.Stream(device=dev)
    return _COPY_STREAMS[key]
