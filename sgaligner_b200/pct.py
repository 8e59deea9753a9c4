"""B200-native ``NaivePCT`` object encoder (``src/aligner/networks/pct.py:275-317``; the ``'pct'`` module of
``MultiModalEncoder``, ``sg_aligner.py:59-60`` -- the point encoder the shipped ground-truth config selects,
``configs/scan3r/scan3r_ground_truth.yaml:5``).

Same sub-module / parameter names as the reference (``embedding.conv1`` ... ``sa4.after_norm``, ``linear.0/1``,
``linear1``, ``linear2``, ``bn1``, ``bn2``; ``q_conv.weight`` IS ``k_conv.weight``, ``pct.py:199``) so a reference
``state_dict`` loads strictly.  The torch modules are parameter containers; the arithmetic runs in the tcgen05 kernels of
``csrc/pct_*.cu``:

    points -> [pct_embed] -> z2 -> 4 x { [pct_pointwise k|v] -> [pct_attn_stats] -> [pct_attn] -> [pct_pointwise trans] }
           -> [pct_cat_linear: concat + 512->1024 conv + max/min pooling] -> head (two small GEMMs)

Every BatchNorm is applied as a folded affine pair by the NEXT kernel's prologue (``ops.bn_fold``), so ``train()``
(batch statistics over the whole batch, running-statistics side effect) and ``eval()`` (running statistics) run the same
kernels; the residual ``x = x + x_s`` (``pct.py:230``) and the ReLUs live in those prologues too.

Forward only: the backward of this encoder is not built (``loss.backward()`` through it raises).  ``P <= 512`` points
per object.  The two ``nn.Dropout(0.5)`` masks are drawn with torch's generator on the device (``dropout_rng = 'cpu'``
draws them with the CPU generator in the reference's order instead -- what the parity tests against CPU goldens use).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops


class Embedding(nn.Module):
    """``pct.py:101-125`` parameter container."""

    def __init__(self, in_channels=3, out_channels=128):
        super().__init__()
        self.conv1 = nn.Conv1d(in_channels, out_channels, kernel_size=1, bias=False)
        self.conv2 = nn.Conv1d(out_channels, out_channels, kernel_size=1, bias=False)
        self.bn1 = nn.BatchNorm1d(out_channels)
        self.bn2 = nn.BatchNorm1d(out_channels)


class SA(nn.Module):
    """``pct.py:187-209`` parameter container; q and k share one weight tensor."""

    def __init__(self, channels):
        super().__init__()
        self.da = channels // 4
        self.q_conv = nn.Conv1d(channels, channels // 4, 1, bias=False)
        self.k_conv = nn.Conv1d(channels, channels // 4, 1, bias=False)
        self.q_conv.weight = self.k_conv.weight
        self.v_conv = nn.Conv1d(channels, channels, 1)
        self.trans_conv = nn.Conv1d(channels, channels, 1)
        self.after_norm = nn.BatchNorm1d(channels)


class _NoBackward(torch.autograd.Function):
    """Marks the encoder output as produced by a non-differentiable (forward-only) path: a backward pass through it
    fails loudly instead of silently training everything but the point encoder."""

    @staticmethod
    def forward(ctx, out, *params):
        return out.clone()

    @staticmethod
    def backward(ctx, g):
        raise NotImplementedError("sgaligner_b200: the backward of the NaivePCT ('pct') object encoder is not implemented; "
                                  "train with modules=['point', ...] or use 'pct' for inference")


class NaivePCT(nn.Module):
    def __init__(self):
        super().__init__()
        self.embedding = Embedding(3, 128)
        self.sa1 = SA(128)
        self.sa2 = SA(128)
        self.sa3 = SA(128)
        self.sa4 = SA(128)
        self.linear = nn.Sequential(nn.Conv1d(512, 1024, kernel_size=1, bias=False), nn.BatchNorm1d(1024), nn.LeakyReLU(negative_slope=0.2))
        self.linear1 = nn.Linear(1024, 512, bias=False)
        self.linear2 = nn.Linear(512, 256)
        self.bn1 = nn.BatchNorm1d(512)
        self.bn2 = nn.BatchNorm1d(256)
        self.dp1 = nn.Dropout(p=0.5)
        self.dp2 = nn.Dropout(p=0.5)
        self.dropout_rng = 'device'

    def _mask(self, n, c, dev):
        if self.dropout_rng == 'cpu':      # the reference's F.dropout on CPU tensors: empty_like(x).bernoulli_(1 - p)
            return torch.empty(n, c).bernoulli_(0.5).to(dev)
        return torch.empty(n, c, device=dev).bernoulli_(0.5)

    def forward(self, pts_npc: torch.Tensor) -> torch.Tensor:
        """``pts_npc``: [N, P, 3] (the collated layout; the reference permutes to [N, 3, P] first).  -> [N, 256]"""
        if not pts_npc.is_cuda:
            raise RuntimeError('sgaligner_b200.NaivePCT needs CUDA tensors (no CPU fallback)')
        N, P, _ = pts_npc.shape
        if P > 512:
            raise NotImplementedError('NaivePCT kernels hold a [128 x P] score block in tensor memory: P <= 512')
        tr = self.training
        if tr and N == 1:
            raise ValueError('Expected more than 1 value per channel when training')      # torch's BatchNorm1d on [1, C]
        grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        with torch.no_grad():
            out = self._forward(pts_npc, N, P, tr)
        if grad:
            out = _NoBackward.apply(out, *[p for p in self.parameters() if p.requires_grad])
        return out

    def _forward(self, pts, N, P, tr):
        emb = self.embedding
        cnt = float(N) * float(P)
        # ---- Embedding (pct.py:120-125)
        st1 = ops.pct_affine_stats(ops.pct_point_moments(pts), emb.conv1.weight) if tr else None
        ab1 = ops.bn_fold(emb.bn1, st1, cnt, tr)
        z2, st2 = ops.pct_embed(pts, emb.conv1.weight, ab1[0], ab1[1], emb.conv2.weight, tr)
        ab2 = ops.bn_fold(emb.bn2, st2, cnt, tr)
        # ---- four self-attention layers (pct.py:211-232); x_l = x_{l-1} + relu(after_norm(t_l)) is formed by the next prologue
        src1, g1 = z2, ab2                 # x0 = relu(bn2(z2))
        src2, g2 = None, None
        xs_saved, t, abt = [], None, None
        for li, sa in enumerate((self.sa1, self.sa2, self.sa3, self.sa4)):
            Wkv = torch.cat([sa.k_conv.weight.reshape(32, 128), sa.v_conv.weight.reshape(128, 128)])
            bkv = torch.cat([torch.zeros(32, device=pts.device), sa.v_conv.bias])
            k, v, x_in, _ = ops.pct_pointwise(src1, g1, src2, g2, Wkv, bkv, 32, want_x=li > 0, want_stats=False)
            if li > 0:
                xs_saved.append(x_in)      # x1, x2, x3
            x_s = ops.pct_attention(k, v)
            t, _, _, stt = ops.pct_pointwise(x_s, None, None, None, sa.trans_conv.weight.reshape(128, 128), sa.trans_conv.bias, 128,
                                             want_x=False, want_stats=tr)
            abt = ops.bn_fold(sa.after_norm, stt, cnt, tr)
            if li == 0:
                src2, g2 = t, abt          # x1 = relu(bn2(z2)) + relu(after_norm(t1))
            else:
                src1, g1, src2, g2 = x_in, None, t, abt
        x1, x2, x3 = xs_saved
        # ---- concat + linear (512 -> 1024) + BN + LeakyReLU + max over points (pct.py:306-310)
        zmax, zmin, stl = ops.pct_cat_linear(x1, x2, x3, t, abt, self.linear[0].weight.reshape(1024, 512))
        abl = ops.bn_fold(self.linear[1], stl if tr else None, cnt, tr)
        pooled = ops.pct_pool_act(zmax, zmin, abl[0], abl[1], P)
        # ---- head (pct.py:311-316)
        y1 = ops.gemm_tf32x3(pooled, self.linear1.weight, N, 512, 1024)
        ab = ops.bn_fold(self.bn1, ops.col_stats(y1) if tr else None, float(N), tr)
        h1 = ops.bn_act_rows(y1, ab[0], ab[1], self._mask(N, 512, pts.device) if tr else None, 2.0)
        y2 = ops.gemm_tf32x3(h1, self.linear2.weight, N, 256, 512)
        ab = ops.bn_fold(self.bn2, ops.col_stats(y2) if tr else None, float(N), tr, lin_bias=self.linear2.bias)
        return ops.bn_act_rows(y2, ab[0], ab[1], self._mask(N, 256, pts.device) if tr else None, 2.0)
