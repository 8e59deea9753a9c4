"""CPU restatement of the ``NaivePCT`` object encoder (the point encoder the shipped
ground-truth config selects: ``configs/scan3r/scan3r_ground_truth.yaml:5``, ``sg_aligner.py:58-59``).
TEST INFRASTRUCTURE ONLY -- groundwork for the CUDA path of a later round; nothing in ``sgaligner_b200/`` imports it.

Every function cites the reference lines it follows (``src/aligner/networks/pct.py``).  Pinned against the unmodified
reference module by ``oracle/make_golden_pct.py`` (``tests/golden/pct_ref.npz``): eval mode, and train mode including
the BatchNorm running-statistics side effect and the two dropouts (same torch RNG consumption order, so the same
seed reproduces the reference's masks).
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


def _bn(x: Tensor, p: Dict[str, Tensor], name: str, training: bool, momentum: float = 0.1, eps: float = 1e-5) -> Tensor:
    """``nn.BatchNorm1d`` on [B, C, N] or [B, C]: batch statistics (biased variance) in training, running statistics
    otherwise; in training the running buffers are updated in place (unbiased variance), as torch does."""
    return F.batch_norm(x, p[name + '.running_mean'], p[name + '.running_var'], p[name + '.weight'], p[name + '.bias'],
                        training, momentum, eps)


def _conv(x: Tensor, p: Dict[str, Tensor], name: str) -> Tensor:
    """``nn.Conv1d(kernel_size=1)`` = a per-point linear map; weight [out, in, 1]."""
    y = torch.einsum('oi,bin->bon', p[name + '.weight'][:, :, 0], x)
    b = p.get(name + '.bias')
    return y if b is None else y + b[None, :, None]


def embedding(x: Tensor, p: Dict[str, Tensor], training: bool, prefix: str = 'embedding') -> Tensor:
    """``Embedding.forward`` (pct.py:101-125): two conv(no bias) + BN + ReLU layers, 3 -> 128 -> 128."""
    x = F.relu(_bn(_conv(x, p, prefix + '.conv1'), p, prefix + '.bn1', training))
    return F.relu(_bn(_conv(x, p, prefix + '.conv2'), p, prefix + '.bn2', training))


def self_attention(x: Tensor, p: Dict[str, Tensor], prefix: str, training: bool) -> Tensor:
    """``SA.forward`` (pct.py:187-232).  q and k share ONE weight (``q_conv.weight = k_conv.weight``, :199), so the
    energy is the Gram matrix k^T k / sqrt(da), da = channels / 4; softmax over the LAST index; the weighted sum
    contracts the FIRST index of the attention map (``torch.bmm(x_v, attention)``, :224); then conv + BN + ReLU and the
    residual."""
    da = p[prefix + '.k_conv.weight'].shape[0]
    x_k = _conv(x, p, prefix + '.k_conv')                     # [B, da, N]   (no bias)
    x_q = x_k.permute(0, 2, 1)                                # [B, N, da]   shared weight
    x_v = _conv(x, p, prefix + '.v_conv')                     # [B, de, N]
    energy = torch.bmm(x_q, x_k) / math.sqrt(da)              # [B, N, N]
    attention = torch.softmax(energy, dim=-1)
    x_s = torch.bmm(x_v, attention)                           # [B, de, N]: x_s[:, j] = sum_i x_v[:, i] attention[i, j]
    x_s = F.relu(_bn(_conv(x_s, p, prefix + '.trans_conv'), p, prefix + '.after_norm', training))
    return x + x_s


def naive_pct(x: Tensor, p: Dict[str, Tensor], training: bool = False) -> Tensor:
    """``NaivePCT.forward`` (pct.py:275-317).  ``x``: [B, 3, N] (the aligner passes ``tot_obj_pts.permute(0, 2, 1)``,
    sg_aligner.py:72).  Returns [B, 256].  In training the two ``nn.Dropout(0.5)`` draw from the global torch RNG in
    this order (dp1 then dp2), exactly as the reference does."""
    x = embedding(x, p, training)
    x1 = self_attention(x, p, 'sa1', training)
    x2 = self_attention(x1, p, 'sa2', training)
    x3 = self_attention(x2, p, 'sa3', training)
    x4 = self_attention(x3, p, 'sa4', training)
    x = torch.cat([x1, x2, x3, x4], dim=1)                    # [B, 512, N]
    x = F.leaky_relu(_bn(_conv(x, p, 'linear.0'), p, 'linear.1', training), negative_slope=0.2)      # 512 -> 1024
    x = torch.max(x, dim=-1)[0]                               # [B, 1024]
    x = F.relu(_bn(x @ p['linear1.weight'].t(), p, 'bn1', training))                                  # 1024 -> 512, no bias
    x = F.dropout(x, 0.5, training)
    x = F.relu(_bn(x @ p['linear2.weight'].t() + p['linear2.bias'], p, 'bn2', training))              # 512 -> 256
    return F.dropout(x, 0.5, training)


def random_params(seed: int = 7) -> Dict[str, Tensor]:
    """A full ``NaivePCT.state_dict()`` (same keys / shapes as the reference module, checked by ``load_state_dict(strict=True)``
    in ``make_golden_pct.py``) drawn from a seeded generator: conv / linear weights ~ N(0, 1/fan_in), non-trivial
    BatchNorm affine parameters and running statistics.  Lets the golden file hold outputs only (the 1.35 M parameters
    are regenerated from the seed)."""
    g = torch.Generator().manual_seed(seed)
    p: Dict[str, Tensor] = {}

    def w(name, shape, fan_in):
        p[name] = torch.randn(*shape, generator=g) / math.sqrt(fan_in)

    def bn(name, c):
        p[name + '.weight'] = 1.0 + 0.2 * torch.randn(c, generator=g)
        p[name + '.bias'] = 0.1 * torch.randn(c, generator=g)
        p[name + '.running_mean'] = 0.1 * torch.randn(c, generator=g)
        p[name + '.running_var'] = 0.5 + torch.rand(c, generator=g)
        p[name + '.num_batches_tracked'] = torch.tensor(0, dtype=torch.long)

    w('embedding.conv1.weight', (128, 3, 1), 3)
    w('embedding.conv2.weight', (128, 128, 1), 128)
    bn('embedding.bn1', 128)
    bn('embedding.bn2', 128)
    for i in (1, 2, 3, 4):
        sa = f'sa{i}'
        w(sa + '.k_conv.weight', (32, 128, 1), 128)
        p[sa + '.q_conv.weight'] = p[sa + '.k_conv.weight']          # one shared tensor (pct.py:199)
        w(sa + '.v_conv.weight', (128, 128, 1), 128)
        p[sa + '.v_conv.bias'] = 0.1 * torch.randn(128, generator=g)
        w(sa + '.trans_conv.weight', (128, 128, 1), 128)
        p[sa + '.trans_conv.bias'] = 0.1 * torch.randn(128, generator=g)
        bn(sa + '.after_norm', 128)
    w('linear.0.weight', (1024, 512, 1), 512)
    bn('linear.1', 1024)
    w('linear1.weight', (512, 1024), 1024)
    w('linear2.weight', (256, 512), 512)
    p['linear2.bias'] = 0.1 * torch.randn(256, generator=g)
    bn('bn1', 512)
    bn('bn2', 256)
    return p


def flops_per_object(n_points: int = 512) -> float:
    """Algorithmic FLOPs of one object (2 per MAC): SURVEY.md 8(f) quotes ~1.06 GFLOP at 512 points."""
    n = n_points
    emb = 2 * n * (3 * 128 + 128 * 128)
    sa = 2 * n * (128 * 32 + 128 * 128 + 128 * 128) + 2 * n * n * (32 + 128)
    lin = 2 * n * 512 * 1024
    head = 2 * (1024 * 512 + 512 * 256)
    return float(emb + 4 * sa + lin + head)
