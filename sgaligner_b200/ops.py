"""Thin torch-side launch wrappers over the C ABI (``include/sga_b200.h``).

PyTorch is used for device memory, streams and autograd bookkeeping only; every arithmetic step
of the hot path runs in ``libsga_b200.so``.  All wrappers launch on the caller's current CUDA
stream and never synchronise.
"""
from __future__ import annotations

import ctypes
import os
from typing import List, Optional, Sequence

import numpy as np
import torch

from ._lib import check, get_lib

POINTNET_SIMT = 0
POINTNET_TC = 1

# ---- instrumentation used by bench.py (never changes what is launched)
LAUNCHES = 0            # kernels of libsga_b200 launched so far (counted per C-ABI call)
KERNEL_EVENTS = None    # when a list: (name, start_event, end_event) of the dominant kernels


def _count(n: int):
    global LAUNCHES
    LAUNCHES += n


class _timed:
    """Brackets one C-ABI call with CUDA events on the launching stream when bench.py asks for it."""

    def __init__(self, name):
        self.name = name
        self.on = KERNEL_EVENTS is not None

    def __enter__(self):
        if self.on:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record(torch.cuda.current_stream())

    def __exit__(self, *a):
        if self.on:
            self.e1.record(torch.cuda.current_stream())
            KERNEL_EVENTS.append((self.name, self.e0, self.e1))


def _ptr(t: Optional[torch.Tensor]):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError('sgaligner_b200: the hot path runs on a CUDA device only (no CPU fallback); '
                               'got a CPU tensor')


def _f32c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


# --------------------------------------------------------------------------------- PointNet
def pointnet_set_max_ctas(n: int):
    """Cap on the SMs the persistent tensor-core PointNet forward occupies (0 = all)."""
    check(get_lib().sga_pointnet_set_max_ctas(int(n)), 'sga_pointnet_set_max_ctas')


_SIDE_STREAMS = []      # side streams that carry work of the current training step (see sg_aligner.MultiModalEncoder.forward)


def note_side_stream(stream):
    if stream is not None and all(stream is not s_ for s_ in _SIDE_STREAMS):
        _SIDE_STREAMS.append(stream)


def join_side_streams():
    """The current stream waits for every side stream noted since the last join.  Autograd orders its own tensors
    across streams, but a backward kernel that accumulates straight into ``p.grad`` (direct gradient accumulation)
    is invisible to it: whoever consumes those buffers (``FlatAdam.allreduce_grads`` / ``step``) joins first."""
    if _SIDE_STREAMS:
        cur = torch.cuda.current_stream()
        for s_ in _SIDE_STREAMS:
            cur.wait_stream(s_)
        del _SIDE_STREAMS[:]


def sm_count() -> int:
    sm, maj, mnr = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int(0)
    check(get_lib().sga_device_info(ctypes.byref(sm), ctypes.byref(maj), ctypes.byref(mnr)), 'sga_device_info')
    return int(sm.value)


def _pointnet_args(pts, W1, b1, W2, b2, W3, b3, mode):
    _need_cuda(pts, W1, W3)
    pts = _f32c(pts)
    C3 = W3.shape[0]
    W1c, W2c, W3c = _f32c(W1.reshape(64, 3)), _f32c(W2.reshape(128, 64)), _f32c(W3.reshape(C3, 128))
    if mode == POINTNET_TC and (W2c.data_ptr() % 16 or W3c.data_ptr() % 16):
        W2c, W3c = W2c.clone(), W3c.clone()
    return pts, (W1c, _f32c(b1), W2c, _f32c(b2), W3c, _f32c(b3))


def pointnet_forward(pts, W1, b1, W2, b2, W3, b3, want_argmax: bool, mode: int = POINTNET_TC, chunks=None, out=None):
    """chunks: optional [(obj_start, obj_end, cuda_event or None)] -- one launch per object range; a range with an
    event becomes valid when it fires (streamed H2D copy, ``data.to_cuda_streamed``), ranges without one only cut the
    encoder into shorter launches (finer-grained SM sharing between concurrent serving steps).
    ``out``: optional preallocated [N, C3] f32 result buffer (a static buffer a captured graph reads)."""
    pts, w = _pointnet_args(pts, W1, b1, W2, b2, W3, b3, mode)
    N, P, _ = pts.shape
    C3 = W3.shape[0]
    if out is None:
        out = torch.empty((N, C3), device=pts.device, dtype=torch.float32)
    else:
        assert out.shape == (N, C3) and out.dtype == torch.float32 and out.is_contiguous() and out.device == pts.device
    arg = torch.empty((N, C3), device=pts.device, dtype=torch.int32) if want_argmax else None
    lib = get_lib()
    if chunks:
        cur = torch.cuda.current_stream()
        for (s, e, ev) in chunks:
            if ev is not None:
                cur.wait_event(ev)
            check(lib.sga_pointnet_fwd(ctypes.c_void_p(pts.data_ptr() + s * P * 12), e - s, P, *[_ptr(t) for t in w], C3,
                                       ctypes.c_void_p(out.data_ptr() + s * C3 * 4),
                                       ctypes.c_void_p(0 if arg is None else arg.data_ptr() + s * C3 * 4), mode, _stream()),
                  'sga_pointnet_fwd')
            _count(2 if (want_argmax and mode == POINTNET_TC) else 1)
        return out, arg
    with _timed('pointnet_fwd'):
        check(lib.sga_pointnet_fwd(_ptr(pts), N, P, *[_ptr(t) for t in w], C3, _ptr(out), _ptr(arg), mode, _stream()),
              'sga_pointnet_fwd')
    _count(2 if (want_argmax and mode == POINTNET_TC) else 1)     # + the fp32 near-tie re-check of the max-pool
    return out, arg


def pointnet_tie_stats(reset: bool = False):
    """Diagnostics: (near-ties of the max-pool re-evaluated in fp32, of those reordered) by the tensor-core forward
    since the last reset.  Synchronises."""
    buf = (ctypes.c_ulonglong * 2)()
    check(get_lib().sga_debug_tie_stats(ctypes.cast(buf, ctypes.c_void_p), 1 if reset else 0), 'sga_debug_tie_stats')
    return int(buf[0]), int(buf[1])


def pointnet_forward_stats(pts, W1, b1, W2, b2, W3, b3, want_argmax: bool):
    """Train-mode forward on the tensor cores: pooled feature, argmax AND the f64 moments of
    ``pointnet_bn_moments`` out of one launch (+ a finalize launch).  Returns (out, argmax, moments)."""
    pts, w = _pointnet_args(pts, W1, b1, W2, b2, W3, b3, POINTNET_TC)
    N, P, _ = pts.shape
    C3 = W3.shape[0]
    out = torch.empty((N, C3), device=pts.device, dtype=torch.float32)
    arg = torch.empty((N, C3), device=pts.device, dtype=torch.int32) if want_argmax else None
    lib = get_lib()
    n_mom = 2 * (64 + 128 + C3)
    sbytes = int(lib.sga_pointnet_stats_scratch_bytes(C3))
    buf = torch.zeros(n_mom + sbytes // 8, device=pts.device, dtype=torch.float64)
    with _timed('pointnet_fwd_stats'):
        check(lib.sga_pointnet_fwd_stats(_ptr(pts), N, P, *[_ptr(t) for t in w], C3, _ptr(out), _ptr(arg), _ptr(buf),
                                         ctypes.c_void_p(buf.data_ptr() + n_mom * 8), sbytes, _stream()), 'sga_pointnet_fwd_stats')
    _count(3 if want_argmax else 2)
    return out, arg, buf[:n_mom]


_DIRECT_GRAD = {}      # id(parameter) -> weakref: parameters whose owner opted in to direct gradient accumulation


def enable_direct_grad(params, on: bool = True):
    """Opt the given parameters in to (or out of) DIRECT gradient accumulation: the backward kernels then add a
    parameter's gradient straight into its kept-allocated ``.grad`` tensor and hand ``None`` to autograd
    (``trainer.FlatAdam`` does this for its flat gradient buffer: no zero-fill and no ``grad += new`` kernel per
    parameter, ~40 launches per step).  It is strictly opt-in, because in that state ``torch.autograd.grad``,
    ``backward(inputs=...)``, gradient hooks and DDP-style reducers do not see these gradients; without it every
    backward returns ordinary gradient tensors to autograd."""
    import weakref
    for p in params:
        if on:
            _DIRECT_GRAD[id(p)] = weakref.ref(p)
        else:
            _DIRECT_GRAD.pop(id(p), None)


def grad_target(p: torch.Tensor, needed: bool = True):
    """The tensor a backward kernel may accumulate a PARAMETER gradient into directly, or None: ``p.grad`` of a
    parameter that was opted in with :func:`enable_direct_grad` (and needs a gradient at all).  All
    parameter-gradient kernels accumulate (atomics / beta = 1)."""
    if not needed:
        return None
    r = _DIRECT_GRAD.get(id(p))
    if r is None or r() is not p:
        return None
    g = getattr(p, 'grad', None)
    if (g is not None and p.is_leaf and g.dtype == torch.float32 and g.is_cuda and g.is_contiguous()
            and g.numel() == p.numel() and g.device == p.device):
        return g
    return None


def _grad_buf(shape, dev, into):
    """(buffer for the kernel, tensor to hand back to autograd or None when accumulated in place)"""
    if into is not None:
        return into, None
    t = torch.zeros(shape, device=dev, dtype=torch.float32)
    return t, t


def pointnet_backward(pts, W1, b1, W2, b2, W3, b3, out, arg, gout, mode: int = POINTNET_SIMT, into=None):
    """``into``: optional 6 tensors (or None entries) to accumulate gW1, gb1, gW2, gb2, gW3, gb3 into; the
    corresponding returned entries are None."""
    pts, w = _pointnet_args(pts, W1, b1, W2, b2, W3, b3, mode)
    N, P, _ = pts.shape
    C3 = W3.shape[0]
    dev = pts.device
    into = into or [None] * 6
    pairs = [_grad_buf(s, dev, t) for s, t in zip(((64, 3), (64,), (128, 64), (128,), (C3, 128), (C3,)), into)]
    g = [b for b, _ in pairs]
    with _timed('pointnet_bwd'):
        check(get_lib().sga_pointnet_bwd_mode(_ptr(pts), N, P, *[_ptr(t) for t in w], C3, _ptr(out), _ptr(arg), _ptr(_f32c(gout)),
                                              *[_ptr(t) for t in g], mode, _stream()), 'sga_pointnet_bwd')
    _count(1)
    return [r for _, r in pairs]


def pointnet_bn_moments(pts, W1, b1, W2, b2, W3, b3):
    pts = _f32c(pts)
    N, P, _ = pts.shape
    C3 = W3.shape[0]
    mom = torch.zeros(2 * (64 + 128 + C3), device=pts.device, dtype=torch.float64)
    check(get_lib().sga_pointnet_bn_moments(_ptr(pts), N, P, _ptr(_f32c(W1.reshape(64, 3))), _ptr(_f32c(b1)),
                                            _ptr(_f32c(W2.reshape(128, 64))), _ptr(_f32c(b2)),
                                            _ptr(_f32c(W3.reshape(C3, 128))), _ptr(_f32c(b3)), C3, _ptr(mom), _stream()),
          'sga_pointnet_bn_moments')
    _count(1)
    return mom


def pointnet_bn_moments_gram(pts, W1, b1, W2, b2, W3, b3):
    """BatchNorm batch statistics from tensor-core Gram matrices (csrc/pointnet_gram.cu); same layout as
    :func:`pointnet_bn_moments`."""
    pts, w = _pointnet_args(pts, W1, b1, W2, b2, W3, b3, POINTNET_TC)
    N, P, _ = pts.shape
    C3 = W3.shape[0]
    lib = get_lib()
    mom = torch.zeros(2 * (64 + 128 + C3), device=pts.device, dtype=torch.float64)
    nb = int(lib.sga_pointnet_gram_scratch_bytes())
    scratch = _workspace_named('gram', nb, pts.device)
    check(lib.sga_pointnet_bn_moments_gram(_ptr(pts), N, P, *[_ptr(t) for t in w], C3, _ptr(mom), _ptr(scratch), scratch.numel(),
                                           _stream()), 'sga_pointnet_bn_moments_gram')
    _count(3)
    return mom


_NAMED_WS = {}


def _workspace_named(name: str, nbytes: int, device) -> torch.Tensor:
    key = (name, device.index if device.index is not None else torch.cuda.current_device())
    ws = _NAMED_WS.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(int(nbytes) + 256, device=device, dtype=torch.uint8)
        _NAMED_WS[key] = ws
    return ws


def bn_running_update(mom: torch.Tensor, cnt: float, bns):
    """One launch for the train-mode running-statistics update of the three (output-discarded) BatchNorm layers."""
    b1, b2, b3 = bns
    assert b1.momentum == b2.momentum == b3.momentum and b1.momentum is not None
    check(get_lib().sga_bn_running_update(_ptr(mom), float(cnt), float(b1.momentum), int(b3.running_mean.numel()),
                                          _ptr(b1.running_mean), _ptr(b1.running_var), _ptr(b2.running_mean), _ptr(b2.running_var),
                                          _ptr(b3.running_mean), _ptr(b3.running_var), _ptr(b1.num_batches_tracked),
                                          _ptr(b2.num_batches_tracked), _ptr(b3.num_batches_tracked), _stream()),
          'sga_bn_running_update')
    _count(1)


# --------------------------------------------------------------------------------- NaivePCT (pct.py:275-317)
def _f64z(n, dev):
    return torch.zeros(n, device=dev, dtype=torch.float64)


def pct_point_moments(pts: torch.Tensor) -> torch.Tensor:
    pts = _f32c(pts)
    mom = _f64z(9, pts.device)
    check(get_lib().sga_pct_point_moments(_ptr(pts), pts.shape[0] * pts.shape[1], _ptr(mom), _stream()), 'sga_pct_point_moments')
    _count(1)
    return mom


def pct_affine_stats(mom: torch.Tensor, W: torch.Tensor) -> torch.Tensor:
    C = W.shape[0]
    stats = torch.empty(2 * C, device=mom.device, dtype=torch.float64)
    check(get_lib().sga_pct_affine_stats(_ptr(mom), _ptr(_f32c(W.reshape(C, 3))), C, _ptr(stats), _stream()), 'sga_pct_affine_stats')
    _count(1)
    return stats


def bn_fold(bn: torch.nn.BatchNorm1d, stats: Optional[torch.Tensor], cnt: float, training: bool, lin_bias=None):
    """(a, b) with BN(z + lin_bias) = a z + b for the stored tensor z; in training the module's running statistics get
    torch's momentum update as a side effect (one launch)."""
    C = bn.num_features
    dev = bn.weight.device
    a = torch.empty(C, device=dev, dtype=torch.float32)
    b = torch.empty(C, device=dev, dtype=torch.float32)
    mom = 0.1 if bn.momentum is None else float(bn.momentum)
    check(get_lib().sga_bn_fold(_ptr(stats), float(cnt), _ptr(lin_bias), _ptr(bn.weight), _ptr(bn.bias), _ptr(bn.running_mean),
                                _ptr(bn.running_var), _ptr(bn.num_batches_tracked), 1 if training else 0, mom, float(bn.eps), C,
                                _ptr(a), _ptr(b), _stream()), 'sga_bn_fold')
    _count(1)
    return a, b


def pct_embed(pts, W1, a1, b1, W2, want_stats: bool):
    pts = _f32c(pts)
    N, P, _ = pts.shape
    z2 = torch.empty((N, P, 128), device=pts.device, dtype=torch.float32)
    stats = _f64z(256, pts.device) if want_stats else None
    check(get_lib().sga_pct_embed(_ptr(pts), N, P, _ptr(_f32c(W1.reshape(128, 3))), _ptr(a1), _ptr(b1), _ptr(_f32c(W2.reshape(128, 128))),
                                  _ptr(z2), _ptr(stats), _stream()), 'sga_pct_embed')
    _count(1)
    return z2, stats


def pct_pointwise(src1, ab1, src2, ab2, W, bias, c0: int, want_x: bool, want_stats: bool):
    """Y = (g1(src1) + g2(src2)) W^T + bias; ab = (a, b) -> g(s) = relu(a s + b), None -> identity.  Returns
    (out0 [N,P,c0], out1 [N,P,Cout-c0] or None, X or None, stats or None)."""
    N, P, _ = src1.shape
    Cout = W.shape[0]
    dev = src1.device
    out0 = torch.empty((N, P, c0), device=dev, dtype=torch.float32)
    out1 = torch.empty((N, P, Cout - c0), device=dev, dtype=torch.float32) if c0 < Cout else None
    out_x = torch.empty((N, P, 128), device=dev, dtype=torch.float32) if want_x else None
    stats = _f64z(2 * Cout, dev) if want_stats else None
    a1, b1 = ab1 if ab1 is not None else (None, None)
    a2, b2 = ab2 if ab2 is not None else (None, None)
    with _timed('pct_pointwise'):
        check(get_lib().sga_pct_pointwise(_ptr(src1), _ptr(a1), _ptr(b1), _ptr(src2), _ptr(a2), _ptr(b2), N, P, _ptr(W), _ptr(bias),
                                          Cout, c0, _ptr(out_x), _ptr(out0), _ptr(out1), _ptr(stats), _stream()), 'sga_pct_pointwise')
    _count(1)
    return out0, out1, out_x, stats


def pct_pointwise_kv(src1, ab1, src2, ab2, W, bias, want_x: bool):
    """The fused k | v convolution of an SA layer (W [160,128]) that also records max |v| per object for the backward.
    -> (k [N,P,32], v [N,P,128], X or None, v_absmax [N])."""
    N, P, _ = src1.shape
    dev = src1.device
    k = torch.empty((N, P, 32), device=dev, dtype=torch.float32)
    v = torch.empty((N, P, 128), device=dev, dtype=torch.float32)
    out_x = torch.empty((N, P, 128), device=dev, dtype=torch.float32) if want_x else None
    v_absmax = torch.zeros(N, device=dev, dtype=torch.float32)
    a1, b1 = ab1 if ab1 is not None else (None, None)
    a2, b2 = ab2 if ab2 is not None else (None, None)
    with _timed('pct_pointwise'):
        check(get_lib().sga_pct_pointwise_kv(_ptr(src1), _ptr(a1), _ptr(b1), _ptr(src2), _ptr(a2), _ptr(b2), N, P, _ptr(W), _ptr(bias),
                                             _ptr(out_x), _ptr(k), _ptr(v), _ptr(v_absmax), _stream()), 'sga_pct_pointwise_kv')
    _count(1)
    return k, v, out_x, v_absmax


def pct_attention(k, v, want_c2: bool = False):
    """x_s = torch.bmm(x_v, softmax(x_k^T x_k / sqrt(32), -1)) for every object (pct.py:217-224), [N,P,128]."""
    N, P, _ = k.shape
    Ppad = (P + 127) // 128 * 128
    c2 = torch.empty((N, 2, Ppad), device=k.device, dtype=torch.float32)      # per row: max * log2(e)/sqrt(32), log2(sum of exponentials)
    xs = torch.empty((N, P, 128), device=k.device, dtype=torch.float32)
    lib = get_lib()
    with _timed('pct_attn_stats'):
        check(lib.sga_pct_attn_stats(_ptr(k), N, P, _ptr(c2), _stream()), 'sga_pct_attn_stats')
    with _timed('pct_attn'):
        check(lib.sga_pct_attn(_ptr(k), _ptr(v), _ptr(c2), N, P, _ptr(xs), _stream()), 'sga_pct_attn')
    _count(2)
    return (xs, c2) if want_c2 else xs


def pct_cat_linear(x1, x2, x3, t4, ab4, WL, track: bool = False):
    """-> (zmax, zmin, stats, imax, imin); the index tensors only with ``track`` (training: the backward needs the arg-max)."""
    N, P, _ = x1.shape
    dev = x1.device
    zmax = torch.empty((N, 2, 1024), device=dev, dtype=torch.float32)
    zmin = torch.empty((N, 2, 1024), device=dev, dtype=torch.float32)
    imax = torch.empty((N, 2, 1024), device=dev, dtype=torch.int32) if track else None
    imin = torch.empty((N, 2, 1024), device=dev, dtype=torch.int32) if track else None
    stats = _f64z(2048, dev)
    lib = get_lib()
    img = None
    if os.environ.get('SGA_PCT_CAT', '') != 'v1':
        # operands pre-split once into the image the MMA reads (the 8 channel-block CTAs then stream it with TMA copies)
        img = _workspace_named('pct_cat_img', int(lib.sga_pct_cat_image_bytes(N, P)), dev)
        with _timed('pct_cat_pack'):
            check(lib.sga_pct_cat_pack(_ptr(x1), _ptr(x2), _ptr(x3), _ptr(t4), _ptr(ab4[0]), _ptr(ab4[1]), N, P, _ptr(img), _stream()),
                  'sga_pct_cat_pack')
        _count(1)
    with _timed('pct_cat_linear'):
        check(lib.sga_pct_cat_linear(_ptr(x1), _ptr(x2), _ptr(x3), _ptr(t4), _ptr(ab4[0]), _ptr(ab4[1]), N, P, _ptr(WL), _ptr(zmax),
                                     _ptr(zmin), _ptr(stats), _ptr(imax), _ptr(imin), _ptr(img), _stream()), 'sga_pct_cat_linear')
    _count(1)
    return zmax, zmin, stats, imax, imin


def pct_pool_act(zmax, zmin, a, b, P: int, imax=None, imin=None):
    """pooled [N,1024]; with the tracked indices -> (pooled, pstar [N,1024] int32, zsel [N,1024])."""
    N = zmax.shape[0]
    dev = zmax.device
    out = torch.empty((N, 1024), device=dev, dtype=torch.float32)
    pstar = torch.empty((N, 1024), device=dev, dtype=torch.int32) if imax is not None else None
    zsel = torch.empty((N, 1024), device=dev, dtype=torch.float32) if imax is not None else None
    check(get_lib().sga_pct_pool_act(_ptr(zmax), _ptr(zmin), _ptr(imax), _ptr(imin), _ptr(a), _ptr(b), N, int(P), _ptr(out), _ptr(pstar),
                                     _ptr(zsel), _stream()), 'sga_pct_pool_act')
    _count(1)
    return out if imax is None else (out, pstar, zsel)


# ---- NaivePCT backward building blocks (csrc/pct_bwd.cu, pct_attn.cu, pct_cat.cu)
def bn_backward(g, y, ab, bn, stats, cnt: float, training: bool, mask=None, scale: float = 1.0, slope: float = 0.0, lin_bias=None,
                want_dy: bool = True, want_absmax: bool = False):
    """Backward of  act(BN(y + lin_bias)) (* mask * scale)  for the STORED y [rows, C]: (dy or None, dgamma, dbeta, (e, f, mean))
    with dy = a gy - e - f (y - mean).
    act = ReLU (slope 0) or LeakyReLU(slope); ``stats`` = the forward's batch statistics {sum y, sum y^2} (training).
    ``want_absmax`` (y [N, P, C]): the last element becomes (e, f, mean, max |dy| per object [N])."""
    C = y.shape[-1]
    rows = y.numel() // C
    dev = y.device
    lib = get_lib()
    sums = _f64z(2 * C, dev)
    check(lib.sga_bn_bwd_stats(_ptr(g), _ptr(y), _ptr(ab[0]), _ptr(ab[1]), _ptr(mask), float(scale), float(slope), rows, C, _ptr(sums),
                               _stream()), 'sga_bn_bwd_stats')
    e = torch.empty(C, device=dev, dtype=torch.float32)
    f = torch.empty(C, device=dev, dtype=torch.float32)
    mean = torch.empty(C, device=dev, dtype=torch.float32)
    dgamma = torch.empty(C, device=dev, dtype=torch.float32)
    dbeta = torch.empty(C, device=dev, dtype=torch.float32)
    check(lib.sga_bn_bwd_coef(_ptr(sums), _ptr(stats if training else None), float(cnt), _ptr(lin_bias), _ptr(bn.weight), _ptr(bn.running_mean),
                              _ptr(bn.running_var), 1 if training else 0, float(bn.eps), C, _ptr(e), _ptr(f), _ptr(mean), _ptr(dgamma),
                              _ptr(dbeta), _stream()), 'sga_bn_bwd_coef')
    dy = None
    if want_dy and want_absmax:
        dy = torch.empty_like(y)
        amax = torch.zeros(y.shape[0], device=dev, dtype=torch.float32)
        check(lib.sga_bn_bwd_apply_absmax(_ptr(g), _ptr(y), _ptr(ab[0]), _ptr(ab[1]), _ptr(mask), float(scale), float(slope), _ptr(e), _ptr(f),
                                          _ptr(mean), rows, C, _ptr(dy), rows // y.shape[0], _ptr(amax), _stream()), 'sga_bn_bwd_apply_absmax')
        _count(3)
        return dy, dgamma, dbeta, (e, f, mean, amax)
    if want_dy:
        dy = torch.empty_like(y)
        check(lib.sga_bn_bwd_apply(_ptr(g), _ptr(y), _ptr(ab[0]), _ptr(ab[1]), _ptr(mask), float(scale), float(slope), _ptr(e), _ptr(f),
                                   _ptr(mean), rows, C, _ptr(dy), _stream()), 'sga_bn_bwd_apply')
    _count(3 if want_dy else 2)
    return dy, dgamma, dbeta, (e, f, mean)


def bn_backward_apply(g, y, ab, slope: float = 0.0):
    """a * g * act'(a y + b)  (the BatchNorm scale and the activation mask only; no batch-statistics terms)."""
    C = y.shape[-1]
    out = torch.empty_like(y)
    check(get_lib().sga_bn_bwd_apply(_ptr(g), _ptr(y), _ptr(ab[0]), _ptr(ab[1]), None, 1.0, float(slope), None, None, None, y.numel() // C, C,
                                     _ptr(out), _stream()), 'sga_bn_bwd_apply')
    _count(1)
    return out


def pct_pow2_scale(x, y=None, target: float = 4096.0, per_object: bool = True):
    """[N,2] = {s, 1/s}: per-object (or, ``per_object=False``, whole-tensor) power-of-two scale of a gradient operand."""
    N = x.shape[0] if per_object else 1
    per = x.numel() // N
    scale = torch.empty((N, 2), device=x.device, dtype=torch.float32)
    check(get_lib().sga_pct_pow2_scale(_ptr(x), _ptr(y), N, per, float(target), _ptr(scale), _stream()), 'sga_pct_pow2_scale')
    _count(1)
    return scale


def pct_attention_backward(k, v, c2, dxs, dxs_absmax=None, v_absmax=None):
    """(dk_row, dk_col [N,P,32] -- their sum is d k --, dv [N,P,128]) of x_s = bmm(x_v, softmax(k^T k / sqrt(32))).
    ``dxs_absmax`` / ``v_absmax`` [N]: the per-object maxima if the producing kernels recorded them (saves a pass over both)."""
    N, P, _ = k.shape
    lib = get_lib()
    # |v_i . dxs_j| <= 128 max|v| max|dxs|: the scaled products (and T = E (dA - delta)) stay below 2^15 < 65504
    if dxs_absmax is not None and v_absmax is not None:
        scale = torch.empty((N, 2), device=k.device, dtype=torch.float32)
        check(lib.sga_pct_scale_from_absmax_pair(_ptr(dxs_absmax), _ptr(v_absmax), N, 16384.0 / 128.0, _ptr(scale), _stream()),
              'sga_pct_scale_from_absmax_pair')
        _count(1)
    else:
        scale = pct_pow2_scale(dxs, v, target=16384.0 / 128.0)
    dv = torch.empty_like(v)
    fused = os.environ.get('SGA_PCT_ATTN', '') != 'v1'      # the two-CTA kernel also records sum_rows dv and max |dv| per object
    dv_colsum = _f64z(128, k.device) if fused else None
    dv_absmax = torch.zeros(N, device=k.device, dtype=torch.float32) if fused else None
    with _timed('pct_attn_bwd_dv'):
        check(lib.sga_pct_attn_bwd_dv(_ptr(k), _ptr(dxs), _ptr(c2), _ptr(scale), N, P, _ptr(dv), _ptr(dv_colsum), _ptr(dv_absmax), _stream()),
              'sga_pct_attn_bwd_dv')
    # delta_i = sum_j A[i,j] dA[i,j] (in scaled units): summed by the row half from ITS OWN dA values in a first sweep, so
    # that the errors of dA and delta cancel in (dA - delta) where the softmax is peaked.  SGA_PCT_DELTA=dot takes the
    # mathematically equal v_i . dv_i from the FMA pipe instead (one S / dA sweep less, -9 ms at 4096 x 512): measured
    # 10x worse on d k_conv.weight at peaked attention (2e-3 .. 2e-2 vs fp64, tools/dbg_pct_attn_bwd.py) -- not the default.
    sweep = os.environ.get('SGA_PCT_DELTA', 'sweep') != 'dot'
    delta = torch.empty((N, P), device=k.device, dtype=torch.float32)
    if not sweep:
        check(lib.sga_pct_rowdot_scaled(_ptr(v), _ptr(dv), _ptr(scale), N, P, _ptr(delta), _stream()), 'sga_pct_rowdot_scaled')
    dk1 = torch.empty_like(k)
    dk2 = torch.empty_like(k)
    with _timed('pct_attn_bwd_dk'):
        check(lib.sga_pct_attn_bwd_dk(_ptr(k), _ptr(v), _ptr(dxs), _ptr(c2), _ptr(delta), _ptr(scale), N, P, 0, 1 if sweep else 0, _ptr(dk1),
                                      _stream()), 'sga_pct_attn_bwd_dk')
        check(lib.sga_pct_attn_bwd_dk(_ptr(k), _ptr(dxs), _ptr(v), _ptr(c2), _ptr(delta), _ptr(scale), N, P, 1, 0, _ptr(dk2), _stream()),
              'sga_pct_attn_bwd_dk')
    _count(3)
    return dk1, dk2, dv, dv_colsum, dv_absmax


def pw2_records_absmax() -> bool:
    """The per-object maxima come from the second-generation pointwise kernel (``SGA_PCT_PW=v1`` selects the first)."""
    return os.environ.get('SGA_PCT_PW', '') != 'v1'


def pct_pointwise_grad(src, Wt, absmax=None, want_absmax: bool = False):
    """src [N,P,128] @ Wt^T (Wt [128,128] contiguous): the input-gradient products; src is scaled per object into the fp16
    range on the way in, the result scaled back.  ``absmax`` [N]: max |src| per object if the producer recorded it (saves
    the reduction pass).  ``want_absmax``: -> (out, max |out| per object [N] or None)."""
    N, P, _ = src.shape
    out = torch.empty_like(src)
    if absmax is not None:
        scale = torch.empty((N, 2), device=src.device, dtype=torch.float32)
        check(get_lib().sga_pct_scale_from_absmax(_ptr(absmax), N, 4096.0, _ptr(scale), _stream()), 'sga_pct_scale_from_absmax')
        _count(1)
    else:
        scale = pct_pow2_scale(src)
    out_absmax = torch.zeros(N, device=src.device, dtype=torch.float32) if (want_absmax and pw2_records_absmax()) else None
    with _timed('pct_pointwise_bwd'):
        if out_absmax is not None:
            check(get_lib().sga_pct_pointwise_scaled_absmax(_ptr(src), _ptr(scale), N, P, _ptr(Wt), _ptr(out), _ptr(out_absmax), _stream()),
                  'sga_pct_pointwise_scaled_absmax')
        else:
            check(get_lib().sga_pct_pointwise_scaled(_ptr(src), _ptr(scale), N, P, _ptr(Wt), _ptr(out), _stream()), 'sga_pct_pointwise_scaled')
    _count(1)
    return (out, out_absmax) if want_absmax else out


def pct_sa_input_grad(gx, gcat, dxv, dk1, dk2, Wk):
    """gx (+ gcat) + dxv + (dk1 + dk2) Wk -> [N,P,128]; dk1 becomes dk1 + dk2 (in place)."""
    out = torch.empty_like(gx)
    check(get_lib().sga_pct_sa_input_grad(_ptr(gx), _ptr(gcat), _ptr(dxv), _ptr(dk1), _ptr(dk2), _ptr(Wk), gx.numel() // 128, _ptr(out),
                                          _stream()), 'sga_pct_sa_input_grad')
    _count(1)
    return out


def pct_residual(x, t, ab):
    out = torch.empty_like(x)
    check(get_lib().sga_pct_residual(_ptr(x), _ptr(t), _ptr(ab[0]), _ptr(ab[1]), x.numel() // 128, _ptr(out), _stream()), 'sga_pct_residual')
    _count(1)
    return out


def pct_embed_a1(pts, W1, ab1):
    N, P, _ = pts.shape
    out = torch.empty((N, P, 128), device=pts.device, dtype=torch.float32)
    check(get_lib().sga_pct_embed_a1(_ptr(pts), _ptr(W1), _ptr(ab1[0]), _ptr(ab1[1]), N * P, _ptr(out), _stream()), 'sga_pct_embed_a1')
    _count(1)
    return out


def pct_embed1_backward(g, pts, W1, ab1, bn, stats, mom, cnt: float, training: bool):
    """(dW1 [128,3], dgamma, dbeta) of relu(bn1(conv1(points))) given g = d/d(that activation) [N,P,128]."""
    dev = g.device
    lib = get_lib()
    sums5 = _f64z(5 * 128, dev)
    check(lib.sga_pct_embed1_bwd_stats(_ptr(g), _ptr(pts), _ptr(W1), _ptr(ab1[0]), _ptr(ab1[1]), g.numel() // 128, _ptr(sums5), _stream()),
          'sga_pct_embed1_bwd_stats')
    e = torch.empty(128, device=dev, dtype=torch.float32)
    f = torch.empty(128, device=dev, dtype=torch.float32)
    mean = torch.empty(128, device=dev, dtype=torch.float32)
    dgamma = torch.empty(128, device=dev, dtype=torch.float32)
    dbeta = torch.empty(128, device=dev, dtype=torch.float32)
    check(lib.sga_bn_bwd_coef(_ptr(sums5), _ptr(stats if training else None), float(cnt), None, _ptr(bn.weight), _ptr(bn.running_mean),
                              _ptr(bn.running_var), 1 if training else 0, float(bn.eps), 128, _ptr(e), _ptr(f), _ptr(mean), _ptr(dgamma),
                              _ptr(dbeta), _stream()), 'sga_bn_bwd_coef')
    dW1 = torch.empty((128, 3), device=dev, dtype=torch.float32)
    check(lib.sga_pct_embed1_wgrad(_ptr(sums5), _ptr(mom), float(cnt), _ptr(W1), _ptr(ab1[0]), _ptr(e), _ptr(f), _ptr(dW1), _stream()),
          'sga_pct_embed1_wgrad')
    _count(3)
    return dW1, dgamma, dbeta


def pct_cat_dense_backward(x1, x2, x3, x4, M, u, xbar):
    """g_x[a] = -u[a] - (M (xcat - xbar))[a]  for the four 128-channel blocks of the concatenation."""
    N, P, _ = x1.shape
    gs = [torch.empty_like(x1) for _ in range(4)]
    scale = pct_pow2_scale(M, per_object=False)
    with _timed('pct_cat_dense_bwd'):
        check(get_lib().sga_pct_cat_dense_bwd(_ptr(x1), _ptr(x2), _ptr(x3), _ptr(x4), N, P, _ptr(M), _ptr(scale), _ptr(u), _ptr(xbar),
                                              *[_ptr(g) for g in gs], _stream()), 'sga_pct_cat_dense_bwd')
    _count(1)
    return gs


def pct_cat_sparse_backward(coef, pstar, WL, xs4, gs, dWL):
    """gs[a][n, pstar, :] += coef WL[c, a-th block]  and  dWL[c, :] += coef xcat[n, pstar, :]."""
    N, P, _ = xs4[0].shape
    lib = get_lib()
    check(lib.sga_pct_cat_sparse_bwd_x(_ptr(coef), _ptr(pstar), _ptr(WL), N, P, *[_ptr(g) for g in gs], _stream()), 'sga_pct_cat_sparse_bwd_x')
    check(lib.sga_pct_cat_sparse_bwd_w(_ptr(coef), _ptr(pstar), *[_ptr(x) for x in xs4], N, P, _ptr(dWL), _stream()), 'sga_pct_cat_sparse_bwd_w')
    _count(2)


def axpby_rows(dst, alpha, src=None, beta=0.0, rowscale=None, gamma=0.0, rowscale2=None, colvec=None):
    R, C = dst.shape
    check(get_lib().sga_axpby_rows(_ptr(dst), float(alpha), _ptr(src), float(beta), _ptr(rowscale), float(gamma), _ptr(rowscale2), _ptr(colvec),
                                   R, C, _stream()), 'sga_axpby_rows')
    _count(1)
    return dst


def pct_wgrad(A, Bs, Cs, transpose: bool = False):
    """C_b[m, n] += sum_r A[r, m] B_b[r, n]  (transpose: C_b[n, m]) for up to four B_b [R, 128 | 32] sharing A [R, 128]; the
    contraction runs over all R rows on the tensor cores (csrc/pct_wgrad.cu)."""
    n = len(Bs)
    R = A.shape[0]
    assert A.shape[1] == 128 and A.is_contiguous() and 1 <= n <= 4
    for b, c in zip(Bs, Cs):
        assert b.shape[0] == R and b.is_contiguous() and b.shape[1] in (128, 32) and c.stride(1) == 1
        assert tuple(c.shape) == ((b.shape[1], 128) if transpose else (128, b.shape[1]))
    arr_p, arr_l, arr_i = ctypes.c_void_p * n, ctypes.c_int64 * n, ctypes.c_int * n
    with _timed('pct_wgrad'):
        check(get_lib().sga_pct_wgrad(_ptr(A), R, arr_p(*[b.data_ptr() for b in Bs]), arr_i(*[b.shape[1] for b in Bs]), n,
                                      arr_p(*[c.data_ptr() for c in Cs]), arr_l(*[c.stride(0) for c in Cs]), 1 if transpose else 0,
                                      _stream()), 'sga_pct_wgrad')
    _count(1)


def wgrad_group(problems):
    """``problems``: list of (A [K,M], B [K,N], C [M,N]); C += A^T B on the tensor cores, one grouped launch per 16."""
    n = len(problems)
    if n == 0:
        return
    K = problems[0][0].shape[0]
    arr_p, arr_l, arr_i = ctypes.c_void_p * n, ctypes.c_int64 * n, ctypes.c_int * n
    A = arr_p(*[a.data_ptr() for a, _, _ in problems])
    B = arr_p(*[b.data_ptr() for _, b, _ in problems])
    C = arr_p(*[c.data_ptr() for _, _, c in problems])
    lda = arr_l(*[a.stride(0) for a, _, _ in problems])
    ldb = arr_l(*[b.stride(0) for _, b, _ in problems])
    ldc = arr_l(*[c.stride(0) for _, _, c in problems])
    M = arr_i(*[a.shape[1] for a, _, _ in problems])
    Nn = arr_i(*[b.shape[1] for _, b, _ in problems])
    for a, b, c in problems:
        assert a.shape[0] == K and b.shape[0] == K and c.shape == (a.shape[1], b.shape[1])
    with _timed('pct_wgrad'):
        check(get_lib().sga_wgrad_group(A, lda, M, B, ldb, Nn, C, ldc, n, K, _stream()), 'sga_wgrad_group')
    _count((n + 15) // 16)


def col_stats(x):
    N, C = x.shape
    stats = _f64z(2 * C, x.device)
    check(get_lib().sga_col_stats(_ptr(x), N, C, _ptr(stats), _stream()), 'sga_col_stats')
    _count(1)
    return stats


def bn_act_rows(x, a, b, mask=None, scale: float = 1.0):
    N, C = x.shape
    out = torch.empty_like(x)
    check(get_lib().sga_bn_act_rows(_ptr(x), _ptr(a), _ptr(b), _ptr(mask), float(scale), N, C, _ptr(out), _stream()), 'sga_bn_act_rows')
    _count(1)
    return out


# --------------------------------------------------------------------------------- graphs
class GraphLayout:
    """Per-graph node / edge offsets of a collated batch (host prefix sums of ``graph_per_obj_count`` /
    ``graph_per_edge_count``, copied to the device once); depends on the counts only, not on the edges."""

    def __init__(self, obj_count: np.ndarray, edge_count: np.ndarray, device):
        oc = np.asarray(obj_count, dtype=np.int64).reshape(-1)
        ec = np.asarray(edge_count, dtype=np.int64).reshape(-1)
        self.G = int(oc.shape[0])
        self.N = int(oc.sum())
        self.E = int(ec.sum())
        self.max_nodes = int(oc.max()) if self.G else 0
        node_off = np.concatenate([[0], np.cumsum(oc)]).astype(np.int32)
        edge_off = np.concatenate([[0], np.cumsum(ec)]).astype(np.int64)
        self.node_off = torch.from_numpy(node_off).to(device, non_blocking=True)
        self.edge_off = torch.from_numpy(edge_off).to(device, non_blocking=True)


class BatchGraph:
    """Device-side block-diagonal CSR (by destination) of all 2B graphs of a collated batch."""

    def __init__(self, edges: torch.Tensor, obj_count: np.ndarray = None, edge_count: np.ndarray = None, layout: GraphLayout = None):
        _need_cuda(edges)
        dev = edges.device
        lay = layout if layout is not None else GraphLayout(obj_count, edge_count, dev)
        self.layout = lay
        self.G, self.N, self.E, self.max_nodes = lay.G, lay.N, lay.E, lay.max_nodes
        self.node_off, self.edge_off = lay.node_off, lay.edge_off
        edges = edges.to(torch.int64).contiguous()
        assert edges.shape[0] == self.E, 'edge tensor does not match graph_per_edge_count'
        self.row_beg = torch.empty(self.N, device=dev, dtype=torch.int32)
        self.row_cnt = torch.empty(self.N, device=dev, dtype=torch.int32)
        self.col = torch.empty(self.E + self.N, device=dev, dtype=torch.int32)
        check(get_lib().sga_csr_build(_ptr(edges), _ptr(self.node_off), _ptr(self.edge_off), self.G, self.max_nodes,
                                      _ptr(self.row_beg), _ptr(self.row_cnt), _ptr(self.col), _stream()), 'sga_csr_build')
        _count(1)


def gat_linear(x: torch.Tensor, W, att_src, att_dst, H: int, C: int):
    _need_cuda(x, W)
    is64 = 1 if x.dtype == torch.float64 else 0
    if not is64:
        x = _f32c(x)
    x = x.contiguous()
    N, in_dim = x.shape
    dev = x.device
    xs = torch.empty((H, N, C), device=dev, dtype=torch.float32)
    a_s = torch.empty((N, H), device=dev, dtype=torch.float32)
    a_d = torch.empty((N, H), device=dev, dtype=torch.float32)
    check(get_lib().sga_gat_linear(_ptr(x), is64, N, in_dim, _ptr(_f32c(W)), _ptr(_f32c(att_src).reshape(-1)),
                                   _ptr(_f32c(att_dst).reshape(-1)), H, C, _ptr(xs), _ptr(a_s), _ptr(a_d), _stream()),
          'sga_gat_linear')
    _count(1)
    return xs, a_s, a_d


def gat_aggregate(xs, a_s, a_d, graph: BatchGraph, bias, apply_elu: bool):
    H, N, C = xs.shape
    out = torch.empty((N, H * C), device=xs.device, dtype=torch.float32)
    check(get_lib().sga_gat_aggregate(_ptr(xs), _ptr(a_s), _ptr(a_d), _ptr(graph.row_beg), _ptr(graph.row_cnt),
                                      _ptr(graph.col), _ptr(graph.node_off), graph.G, graph.max_nodes, N, H, C,
                                      _ptr(_f32c(bias)), 1 if apply_elu else 0, _ptr(out), _stream()), 'sga_gat_aggregate')
    _count(1)
    return out


def as_f32(x: torch.Tensor) -> torch.Tensor:
    """f32 contiguous view/copy of a dataloader tensor; f64 inputs go through the library's cast
    kernel (the reference's `.float()`, sg_aligner.py:73-75)."""
    if x.dtype == torch.float64:
        x = x.contiguous()
        out = torch.empty(x.shape, device=x.device, dtype=torch.float32)
        check(get_lib().sga_cast_f64_f32(_ptr(x), _ptr(out), x.numel(), _stream()), 'sga_cast_f64_f32')
        _count(1)
        return out
    return _f32c(x)


def gat_aggregate_backward(xs, a_s, a_d, graph: BatchGraph, apply_elu: bool, out, gout, into_bias=None):
    H, N, C = xs.shape
    dev = xs.device
    g_xs = torch.zeros_like(xs)
    g_as = torch.zeros((N, H), device=dev, dtype=torch.float32)
    g_ad = torch.zeros((N, H), device=dev, dtype=torch.float32)
    g_bias, g_bias_ret = _grad_buf(H * C, dev, into_bias)
    check(get_lib().sga_gat_aggregate_bwd(_ptr(xs), _ptr(a_s), _ptr(a_d), _ptr(graph.row_beg), _ptr(graph.row_cnt),
                                          _ptr(graph.col), _ptr(graph.node_off), graph.G, graph.max_nodes, N, H, C,
                                          1 if apply_elu else 0, _ptr(out), _ptr(_f32c(gout)),
                                          _ptr(g_xs), _ptr(g_as), _ptr(g_ad), _ptr(g_bias), _stream()),
          'sga_gat_aggregate_bwd')
    _count(1)
    return g_xs, g_as, g_ad, g_bias_ret


def gat_linear_backward(x, W, att_src, att_dst, H, C, xs, g_xs, g_as, g_ad, need_gx: bool, into=None):
    """``into``: optional (gW, g_att_src, g_att_dst) accumulation targets (see :func:`grad_target`)."""
    x = as_f32(x)
    N, in_dim = x.shape
    dev = x.device
    into = into or (None, None, None)
    gW, gW_ret = _grad_buf((H * C, in_dim), dev, into[0])
    g_att_s, g_att_s_ret = _grad_buf(H * C, dev, into[1])
    g_att_d, g_att_d_ret = _grad_buf(H * C, dev, into[2])
    gx = torch.empty((N, in_dim), device=dev, dtype=torch.float32) if need_gx else None
    check(get_lib().sga_gat_linear_bwd(_ptr(x), N, in_dim, _ptr(_f32c(W)), _ptr(_f32c(att_src).reshape(-1)),
                                       _ptr(_f32c(att_dst).reshape(-1)), H, C, _ptr(xs), _ptr(g_xs), _ptr(g_as), _ptr(g_ad),
                                       _ptr(gW), _ptr(g_att_s), _ptr(g_att_d), _ptr(gx), _stream()), 'sga_gat_linear_bwd')
    _count(1 + H * (2 if need_gx else 1))
    return gW_ret, g_att_s_ret, g_att_d_ret, gx


# --------------------------------------------------------------------------------- projection + fusion
def project_fuse(x, W, b, joint: Optional[torch.Tensor], joint_col: int, fusion_w, M: int, m: int):
    _need_cuda(x, W)
    is64 = 1 if x.dtype == torch.float64 else 0
    if not is64:
        x = _f32c(x)
    x = x.contiguous()
    N, in_dim = x.shape
    out_dim = W.shape[0]
    emb = torch.empty((N, out_dim), device=x.device, dtype=torch.float32)
    fw = None if joint is None else _f32c(fusion_w).reshape(-1)
    check(get_lib().sga_project_fuse_fwd(_ptr(x), is64, N, in_dim, _ptr(_f32c(W)), _ptr(_f32c(b)), out_dim, _ptr(emb),
                                         _ptr(joint), 0 if joint is None else joint.shape[1], joint_col, _ptr(fw), M, m,
                                         _stream()), 'sga_project_fuse_fwd')
    _count(1)
    return emb


def project_fuse_multi(xs, Ws, bs, fusion_w, want_joint: bool):
    """Every modality's Linear + its slice of the fusion in ONE launch.  All modalities project to the same
    width (emb_dim).  Returns ([emb_m], joint or None)."""
    M = len(xs)
    _need_cuda(*xs, *Ws)
    N = xs[0].shape[0]
    out_dim = Ws[0].shape[0]
    dev = xs[0].device
    xs = [(x if x.dtype == torch.float64 else _f32c(x)).contiguous() for x in xs]
    Wc = [_f32c(W) for W in Ws]
    bc = [_f32c(b) for b in bs]
    embs = [torch.empty((N, out_dim), device=dev, dtype=torch.float32) for _ in range(M)]
    joint = torch.empty((N, M * out_dim), device=dev, dtype=torch.float32) if want_joint else None
    fw = _f32c(fusion_w).reshape(-1) if want_joint else None
    arr_p = ctypes.c_void_p * M
    arr_i = ctypes.c_int * M
    check(get_lib().sga_project_fuse_fwd_multi(arr_p(*[x.data_ptr() for x in xs]), arr_i(*[1 if x.dtype == torch.float64 else 0 for x in xs]),
                                               arr_i(*[int(x.shape[1]) for x in xs]), arr_p(*[w.data_ptr() for w in Wc]),
                                               arr_p(*[b.data_ptr() for b in bc]), arr_p(*[e.data_ptr() for e in embs]), M, N, out_dim,
                                               _ptr(joint), 0 if joint is None else joint.shape[1], _ptr(fw), _stream()),
          'sga_project_fuse_fwd_multi')
    _count(1)
    return embs, joint


def project_fuse_backward(x, W, emb, g_emb, g_joint, joint_col: int, fusion_w, M: int, m: int, need_gx: bool, into=None):
    """``into``: optional (gW, gb, g_fusion_w) accumulation targets (see :func:`grad_target`); the kernels add
    into them, so one ``g_fusion_w`` buffer can be shared by the M calls of a step."""
    x = as_f32(x)
    N, in_dim = x.shape
    out_dim = W.shape[0]
    dev = x.device
    into = into or (None, None, None)
    gW, gW_ret = _grad_buf((out_dim, in_dim), dev, into[0])
    gb, gb_ret = _grad_buf(out_dim, dev, into[1])
    gfw, gfw_ret = _grad_buf(M, dev, into[2])
    gx = torch.empty((N, in_dim), device=dev, dtype=torch.float32) if need_gx else None
    ws = torch.empty(N * out_dim + 64, device=dev, dtype=torch.float32)
    fw = None if g_joint is None else _f32c(fusion_w).reshape(-1)
    check(get_lib().sga_project_fuse_bwd(_ptr(x), N, in_dim, _ptr(_f32c(W)), out_dim, _ptr(emb),
                                         _ptr(None if g_emb is None else _f32c(g_emb)),
                                         _ptr(None if g_joint is None else _f32c(g_joint)),
                                         0 if g_joint is None else g_joint.shape[1], joint_col, _ptr(fw), M, m,
                                         _ptr(gW), _ptr(gb), _ptr(gfw), _ptr(gx), _ptr(ws), ws.numel() * 4, _stream()),
          'sga_project_fuse_bwd')
    _count(3 + (1 if g_joint is not None else 0) + (1 if need_gx else 0))
    return gW_ret, gb_ret, gfw_ret, gx


# --------------------------------------------------------------------------------- matching
class PairLayout:
    """Per-pair node offsets of a collated batch, resident on the device."""

    def __init__(self, obj_count: np.ndarray, device):
        n = np.asarray(obj_count, dtype=np.int64).reshape(-1, 2).sum(1)
        self.B = int(n.shape[0])
        self.n = n
        self.N = int(n.sum())
        self.max_n = int(n.max()) if self.B else 0
        self.pair_off_host = np.concatenate([[0], np.cumsum(n)]).astype(np.int32)
        self.sim_off_host = np.concatenate([[0], np.cumsum(n * n)]).astype(np.int64)
        node_pair = np.repeat(np.arange(self.B, dtype=np.int32), n)
        self.pair_off = torch.from_numpy(self.pair_off_host).to(device, non_blocking=True)
        self.sim_off = torch.from_numpy(self.sim_off_host).to(device, non_blocking=True)
        self.node_pair = torch.from_numpy(node_pair).to(device, non_blocking=True)


def match_sim(emb: torch.Tensor, lay: PairLayout):
    _need_cuda(emb)
    emb = _f32c(emb)
    N, D = emb.shape
    norms = torch.empty(N, device=emb.device, dtype=torch.float32)
    sim = torch.empty(int(lay.sim_off_host[-1]), device=emb.device, dtype=torch.float32)
    check(get_lib().sga_match_sim(_ptr(emb), N, D, _ptr(lay.pair_off), _ptr(lay.sim_off), lay.B, lay.max_n, _ptr(norms),
                                  _ptr(sim), _stream()), 'sga_match_sim')
    _count(2)
    return sim


def match_topk_tc(emb: torch.Tensor, lay: PairLayout, K: int, want_sim: bool):
    """Fused tcgen05 Gram + per-row top-K (K <= 8); returns (topk_idx, topk_dist, sim or None)."""
    _need_cuda(emb)
    emb = _f32c(emb)
    N, D = emb.shape
    dev = emb.device
    norms = torch.empty(N, device=dev, dtype=torch.float32)
    topk_idx = torch.empty((N, K), device=dev, dtype=torch.int32) if K > 0 else None
    topk_dist = torch.empty((N, K), device=dev, dtype=torch.float32) if K > 0 else None
    sim = torch.empty(int(lay.sim_off_host[-1]), device=dev, dtype=torch.float32) if want_sim else None
    with _timed('match_topk_tc'):
        check(get_lib().sga_match_topk_tc(_ptr(emb), N, D, _ptr(lay.pair_off), _ptr(lay.sim_off), lay.B, lay.max_n, K, _ptr(norms),
                                          _ptr(topk_idx), _ptr(topk_dist), _ptr(sim), _stream()), 'sga_match_topk_tc')
    _count(2)
    return topk_idx, topk_dist, sim


def match_rank(sim: torch.Tensor, lay: PairLayout, K: int, full: bool):
    dev = sim.device
    topk_idx = torch.empty((lay.N, K), device=dev, dtype=torch.int32) if K > 0 else None
    topk_dist = torch.empty((lay.N, K), device=dev, dtype=torch.float32) if K > 0 else None
    rank = torch.empty(sim.numel(), device=dev, dtype=torch.int32) if full else None
    check(get_lib().sga_match_rank(_ptr(sim), lay.N, _ptr(lay.pair_off), _ptr(lay.sim_off), _ptr(lay.node_pair), lay.max_n, K,
                                   _ptr(topk_idx), _ptr(topk_dist), _ptr(rank), _stream()), 'sga_match_rank')
    _count(1)
    return topk_idx, topk_dist, rank


def match_anchor_pos(sim: torch.Tensor, lay: PairLayout, e1i: torch.Tensor, e2i: torch.Tensor):
    A = int(e1i.numel())
    pos = torch.empty(A, device=sim.device, dtype=torch.int32)
    check(get_lib().sga_match_anchor_pos(_ptr(sim), _ptr(lay.pair_off), _ptr(lay.sim_off), _ptr(lay.node_pair), _ptr(e1i),
                                         _ptr(e2i), A, _ptr(pos), _stream()), 'sga_match_anchor_pos')
    _count(1)
    return pos


def center_points(pts: torch.Tensor, center: torch.Tensor, node_pair: torch.Tensor):
    """In place ``pts[o] -= center[node_pair[o]]`` (the centring of ``scan3r.py:99-100`` on the device)."""
    _need_cuda(pts, center, node_pair)
    assert pts.dtype == torch.float32 and pts.is_contiguous() and center.dtype == torch.float32 and center.is_contiguous()
    N, P = int(pts.shape[0]), int(pts.shape[1])
    check(get_lib().sga_center_points(_ptr(pts), N, P, _ptr(center), _ptr(node_pair), _stream()), 'sga_center_points')
    _count(1)
    return pts


def match_pair_metrics(sim: torch.Tensor, lay: PairLayout, n_src: torch.Tensor, e1i: torch.Tensor, e2i: torch.Tensor,
                       anchor_off: torch.Tensor):
    """SGAR / alignment score / top-1 correspondences of every pair (one launch).  Returns (top1_idx [N] int32
    pair-local, top1_dist [N], pair_out [B,4] = sgar2, sgar50, sgar100, alignment score)."""
    dev = sim.device
    top1 = torch.empty(lay.N, device=dev, dtype=torch.int32)
    dist = torch.empty(lay.N, device=dev, dtype=torch.float32)
    out = torch.empty(lay.B, 4, device=dev, dtype=torch.float32)
    check(get_lib().sga_match_pair_metrics(_ptr(sim), _ptr(lay.pair_off), _ptr(lay.sim_off), _ptr(n_src), _ptr(e1i), _ptr(e2i),
                                           _ptr(anchor_off), lay.B, _ptr(top1), _ptr(dist), _ptr(out), _stream()),
          'sga_match_pair_metrics')
    _count(1)
    return top1, dist, out


# --------------------------------------------------------------------------------- loss
_WS_CACHE = {}


def _workspace(nbytes: int, device) -> torch.Tensor:
    key = (device.index if device.index is not None else torch.cuda.current_device())
    ws = _WS_CACHE.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(int(nbytes * 1.25) + 256, device=device, dtype=torch.uint8)
        _WS_CACHE[key] = ws
    return ws


def loss_forward_backward(embs: Sequence[torch.Tensor], idx: Sequence[torch.Tensor], lv_ial, lv_icl, zoom: float,
                          want_grad: bool, partition: bool = True):
    """embs: M modal embeddings then the joint (or a single embedding).  idx: e1i,e2i,e1j,e2j int32
    device tensors.  Returns (losses[4], grads or None, g_lv_ial, g_lv_icl).
    ``partition``: the four index sets are disjoint and duplicate-free (true for every collated batch,
    ``scan3r.py:101-107``; ``losses._index_tensors`` checks it on the host arrays) -- required by the packed-image
    Gram kernel; False selects the gather-in-the-loader GEMM for all Grams."""
    lib = get_lib()
    lib.sga_loss_set_gram_path(0 if partition else 1)
    embs = [_f32c(e) for e in embs]
    _need_cuda(*embs)
    n_emb = len(embs)
    dev = embs[0].device
    N = embs[0].shape[0]
    dims = (ctypes.c_int * n_emb)(*[int(e.shape[1]) for e in embs])
    A, J1, J2 = int(idx[0].numel()), int(idx[2].numel()), int(idx[3].numel())
    nbytes = lib.sga_loss_workspace_bytes(n_emb, dims, N, A, J1, J2, 1 if want_grad else 0)
    ws = _workspace(nbytes, dev)
    losses = torch.empty(4, device=dev, dtype=torch.float32)
    grads = [torch.empty_like(e) for e in embs] if want_grad else None
    M = 1 if n_emb == 1 else n_emb - 1
    g_ial = torch.zeros(M, device=dev, dtype=torch.float32) if want_grad else None
    g_icl = torch.zeros(M, device=dev, dtype=torch.float32) if want_grad else None
    e_ptrs = (ctypes.c_void_p * n_emb)(*[e.data_ptr() for e in embs])
    g_ptrs = (ctypes.c_void_p * n_emb)(*[(g.data_ptr() if want_grad else 0) for g in (grads or embs)])
    check(lib.sga_loss_fwd_bwd(e_ptrs, dims, n_emb, N, _ptr(idx[0]), _ptr(idx[1]), _ptr(idx[2]), _ptr(idx[3]), A, J1, J2,
                               _ptr(None if lv_ial is None else _f32c(lv_ial)), _ptr(None if lv_icl is None else _f32c(lv_icl)),
                               float(zoom), _ptr(losses), 1 if want_grad else 0, g_ptrs, _ptr(g_ial), _ptr(g_icl),
                               _ptr(ws), ws.numel(), _stream()), 'sga_loss_fwd_bwd')
    _count(lib.sga_loss_launch_count(n_emb, dims, J1, J2, 1 if want_grad else 0))
    return losses, grads, g_ial, g_icl


def gemm_tf32x3(A, B, M, N, K, a_mn=False, b_mn=False, a_idx=None, b_idx=None, a_div=None, b_div=None, out=None, c_idx=None,
                ksplit=1):
    """C[M,N] = sum_k A(m,k) B(n,k) on the tensor cores (see sga_gemm_tf32x3 in the header)."""
    _need_cuda(A, B)
    if out is None:
        out = torch.zeros((M, N), device=A.device, dtype=torch.float32)
    check(get_lib().sga_gemm_tf32x3(_ptr(A), A.stride(0), 1 if a_mn else 0, _ptr(a_idx), _ptr(a_div), _ptr(B), B.stride(0),
                                    1 if b_mn else 0, _ptr(b_idx), _ptr(b_div), M, N, K, _ptr(out), out.stride(0), _ptr(c_idx),
                                    int(ksplit), _stream()), 'sga_gemm_tf32x3')
    _count(1)
    return out


# --------------------------------------------------------------------------------- EVA baseline (eva.py, gat.py:6-25, losses.py:154-205)
def gcn_aggregate(h, graph: 'BatchGraph', deg, bias=None, relu: bool = False):
    """out_i = sum_{j in CSR row i of ``graph``} h_j / sqrt(deg_i deg_j) (+ bias) (ReLU); ``deg`` = row_cnt of the forward graph."""
    h = _f32c(h)
    N, C = h.shape
    out = torch.empty_like(h)
    check(get_lib().sga_gcn_aggregate(_ptr(h), N, C, _ptr(graph.row_beg), _ptr(graph.row_cnt), _ptr(graph.col), _ptr(deg),
                                      _ptr(None if bias is None else _f32c(bias)), 1 if relu else 0, _ptr(out), _stream()), 'sga_gcn_aggregate')
    _count(1)
    return out


def linear_nobias(x, W):
    """x [N,K] W^T, W [C,K]: the small-K kernel for K <= 8 (GCN layer 1), the tf32x3 tensor-core GEMM otherwise."""
    x, W = _f32c(x), _f32c(W)
    N, K = x.shape
    C = W.shape[0]
    if K <= 8:
        y = torch.empty((N, C), device=x.device, dtype=torch.float32)
        check(get_lib().sga_linear_smallk(_ptr(x), N, K, _ptr(W), C, _ptr(y), _stream()), 'sga_linear_smallk')
        _count(1)
        return y
    return gemm_tf32x3(x, W, N, C, K)


def linear_nobias_backward(x, W, g, need_gx: bool):
    """(gW [C,K], gx [N,K] or None) of y = x W^T."""
    x, W, g = _f32c(x), _f32c(W), _f32c(g)
    N, K = x.shape
    C = W.shape[0]
    if K <= 8:
        gW = torch.zeros((C, K), device=x.device, dtype=torch.float32)
        check(get_lib().sga_wgrad_smallk(_ptr(g), _ptr(x), N, K, C, _ptr(gW), _stream()), 'sga_wgrad_smallk')
        _count(1)
        gx = gemm_tf32x3(g, W, N, K, C, b_mn=True) if need_gx else None
        return gW, gx
    gW = gemm_tf32x3(g, x, C, K, N, a_mn=True, b_mn=True)          # gW[c,k] = sum_i g[i,c] x[i,k]
    gx = gemm_tf32x3(g, W, N, K, C, b_mn=True) if need_gx else None   # gx[i,k] = sum_c g[i,c] W[c,k]
    return gW, gx


def relu_mask(g, y):
    g, y = _f32c(g), _f32c(y)
    out = torch.empty_like(g)
    check(get_lib().sga_relu_mask(_ptr(g), _ptr(y), g.numel(), _ptr(out), _stream()), 'sga_relu_mask')
    _count(1)
    return out


def colsum_rows(x):
    x = _f32c(x)
    N, C = x.shape
    s = torch.zeros(C, device=x.device, dtype=torch.float32)
    check(get_lib().sga_colsum_rows(_ptr(x), N, C, _ptr(s), _stream()), 'sga_colsum_rows')
    _count(1)
    return s


def row_l2norm(x, eps: float = 1e-12):
    x = _f32c(x)
    N, D = x.shape
    n = torch.empty(N, device=x.device, dtype=torch.float32)
    check(get_lib().sga_row_l2norm(_ptr(x), N, D, float(eps), _ptr(n), _stream()), 'sga_row_l2norm')
    _count(1)
    return n


def normalize_backward_rows_(x, norms, g, eps: float = 1e-12):
    """in place on g [N,D]: the backward of F.normalize(x, eps)."""
    N, D = x.shape
    check(get_lib().sga_normalize_bwd_rows(_ptr(_f32c(x)), _ptr(norms), N, D, float(eps), _ptr(g), _stream()), 'sga_normalize_bwd_rows')
    _count(1)
    return g


def fuse_rows(xs, fusion_w):
    """MultiModalFusion over modalities of different widths -> joint [N, sum d_m]."""
    xs = [_f32c(x) for x in xs]
    _need_cuda(*xs)
    M, N = len(xs), xs[0].shape[0]
    dims = [int(x.shape[1]) for x in xs]
    joint = torch.empty((N, sum(dims)), device=xs[0].device, dtype=torch.float32)
    arr_p, arr_i = ctypes.c_void_p * M, ctypes.c_int * M
    check(get_lib().sga_fuse_rows_fwd(arr_p(*[x.data_ptr() for x in xs]), arr_i(*dims), M, _ptr(_f32c(fusion_w).reshape(-1)), N,
                                      _ptr(joint), joint.shape[1], _stream()), 'sga_fuse_rows_fwd')
    _count(1)
    return joint


def fuse_rows_backward(xs, fusion_w, g_joint):
    xs = [_f32c(x) for x in xs]
    M, N = len(xs), xs[0].shape[0]
    dims = [int(x.shape[1]) for x in xs]
    g_joint = _f32c(g_joint)
    gxs = [torch.empty_like(x) for x in xs]
    g_fw = torch.zeros(M, device=xs[0].device, dtype=torch.float32)
    scratch = torch.empty(M, device=xs[0].device, dtype=torch.float32)
    arr_p, arr_i = ctypes.c_void_p * M, ctypes.c_int * M
    check(get_lib().sga_fuse_rows_bwd(arr_p(*[x.data_ptr() for x in xs]), arr_i(*dims), M, _ptr(_f32c(fusion_w).reshape(-1)), N,
                                      _ptr(g_joint), g_joint.shape[1], arr_p(*[g.data_ptr() for g in gxs]), _ptr(g_fw), _ptr(scratch),
                                      _stream()), 'sga_fuse_rows_bwd')
    _count(2)
    return gxs, g_fw


def nca_loss_forward_backward(emb, e1, e2, alpha: float, beta: float, ep: float, want_grad: bool, normalize: bool = True):
    """NCALoss(F.normalize(emb)[e1], F.normalize(emb)[e2]) (losses.py:161-176,189-200) and, if asked, its gradient
    w.r.t. ``emb`` [N,D] (``normalize=False``: the rows are used as they are).  Score matrix and both gradient products on the tf32x3 tensor-core GEMM (row gathers and
    the 1/norm scaling in its loaders, scatter-add of the rows in its epilogue)."""
    emb = _f32c(emb)
    N, D = emb.shape
    A = int(e1.numel())
    dev = emb.device
    norms = row_l2norm(emb) if normalize else None
    S = gemm_tf32x3(emb, emb, A, A, D, a_idx=e1, b_idx=e2, a_div=norms, b_div=norms)
    rs = torch.empty(A, device=dev, dtype=torch.float32)
    cs = torch.empty(A, device=dev, dtype=torch.float32)
    dg = torch.empty(A, device=dev, dtype=torch.float32)
    loss = torch.empty(1, device=dev, dtype=torch.float32)
    lib = get_lib()
    check(lib.sga_nca_forward(_ptr(S), A, float(alpha), float(beta), float(ep), _ptr(rs), _ptr(cs), _ptr(dg), _ptr(loss), _stream()),
          'sga_nca_forward')
    _count(3)
    if not want_grad:
        return loss[0], None
    check(lib.sga_nca_coef(_ptr(S), A, float(alpha), float(beta), float(ep), _ptr(rs), _ptr(cs), _stream()), 'sga_nca_coef')
    _count(1)
    g = torch.zeros((N, D), device=dev, dtype=torch.float32)
    # d/dP = dS Q^ (rows scattered to e1), d/dQ = dS^T P^ (rows scattered to e2)
    gemm_tf32x3(S, emb, A, D, A, b_mn=True, b_idx=e2, b_div=norms, out=g, c_idx=e1)
    gemm_tf32x3(S, emb, A, D, A, a_mn=True, b_mn=True, b_idx=e1, b_div=norms, out=g, c_idx=e2)
    if normalize:
        normalize_backward_rows_(emb, norms, g)
    return loss[0], g


# --------------------------------------------------------------------------------- optimiser
def adam_step(param, grad, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, weight_decay, step, grad_scale=1.0):
    check(get_lib().sga_adam_step(_ptr(param), _ptr(grad), _ptr(exp_avg), _ptr(exp_avg_sq), param.numel(), lr, beta1, beta2,
                                  eps, weight_decay, int(step), float(grad_scale), _stream()), 'sga_adam_step')
    _count(1)


def adam_step_segments(param, grad, exp_avg, exp_avg_sq, seg_off, seg_active, lr, beta1, beta2, eps, weight_decay, step,
                       grad_scale=1.0):
    """Adam over a flat buffer that leaves segments with an all-zero gradient untouched (torch skips grad-None params)."""
    check(get_lib().sga_adam_step_segments(_ptr(param), _ptr(grad), _ptr(exp_avg), _ptr(exp_avg_sq), param.numel(), _ptr(seg_off),
                                           int(seg_off.numel() - 1), _ptr(seg_active), lr, beta1, beta2, eps, weight_decay,
                                           int(step), float(grad_scale), _stream()), 'sga_adam_step_segments')
    _count(2)


def selftest_umma(A: torch.Tensor, B: torch.Tensor, kind: int) -> torch.Tensor:
    D = torch.empty((128, B.shape[0]), device=A.device, dtype=torch.float32)
    check(get_lib().sga_selftest_umma(_ptr(_f32c(A)), _ptr(_f32c(B)), _ptr(D), B.shape[0], A.shape[1], kind, _stream()),
          'sga_selftest_umma')
    return D
