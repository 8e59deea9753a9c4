"""B200-native mirror of ``src/aligner/losses.py`` (ICL / IAL contrastive losses + uncertainty
weighting).  Same class names, constructor signatures and returned dict keys; the arithmetic runs
in ``sga_loss_fwd_bwd`` (one fused forward+gradient pass instead of 78 matmuls)."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F   # noqa: F401
from torch import nn

from . import autograd as ag

__all__ = ['torch', 'nn', 'F', 'calculate_prob_dist', 'CustomMultiLossLayer', 'ICLLoss', 'IALLoss', 'OverallLoss', 'NCALoss', 'OverallNCALoss']


class _IndexSets(list):
    """[e1i, e2i, e1j, e2j] device tensors + ``partition`` (are the sets disjoint and duplicate-free?)."""
    partition = True
    max_index = -1


def _index_tensors(data_dict, device, n_rows=None):
    """e1i/e2i/e1j/e2j arrive as host int32 numpy arrays (scan3r.py:168-171); cache the device copies
    in the dict so that several loss calls on one batch upload them once.  ``n_rows``: rows of the embedding the
    indices address -- an index outside [0, n_rows) raises ``IndexError`` like the reference's ``emb[e1i]`` gather
    (``losses.py:46-49``) instead of reaching the kernels (negative indices are rejected too, not wrapped)."""
    cached = data_dict.get('_sga_idx')
    if cached is not None and cached[0].device == device:
        if n_rows is not None and cached.max_index >= n_rows:
            raise IndexError(f'index {cached.max_index} is out of bounds for dimension 0 with size {n_rows}')
        return cached
    host = [np.ascontiguousarray(np.asarray(data_dict[k]).astype(np.int32)).reshape(-1) for k in ('e1i', 'e2i', 'e1j', 'e2j')]
    idx = _IndexSets(torch.as_tensor(h).to(device, non_blocking=True) for h in host)
    # the collated sets partition the nodes (scan3r.py:101-107); anything else (hand-made overlapping / repeated
    # indices) is still computed correctly, by the Gram path that gathers rows instead of using packed images
    allidx = np.concatenate(host)
    idx.max_index = int(allidx.max()) if allidx.size else -1
    if allidx.size and int(allidx.min()) < 0:
        raise IndexError(f'negative node index {int(allidx.min())} in e1i/e2i/e1j/e2j')
    if n_rows is not None and idx.max_index >= n_rows:
        raise IndexError(f'index {idx.max_index} is out of bounds for dimension 0 with size {n_rows}')
    idx.partition = bool(allidx.size == 0 or (allidx.min() >= 0 and np.bincount(allidx).max() <= 1))
    try:
        data_dict['_sga_idx'] = idx
    except TypeError:
        pass
    return idx


def calculate_prob_dist(e1i, e2i, e1j, e2j, temp):
    """``losses.py:5-15`` for stand-alone callers (not used by the fused loss, which never
    materialises these matrices per call)."""
    mx = torch.exp(e1i @ e2i.t() / temp)
    sa = torch.exp(e1i @ e1j.t() / temp).sum()
    sb = torch.exp(e1i @ e2j.t() / temp).sum()
    inv = 1.0 + 1.0 / (mx / (sa + 1e-9) + 1e-9) + 1.0 / (mx / (sb + 1e-9) + 1e-9)
    return 1.0 / (inv + 1e-9)


class CustomMultiLossLayer(nn.Module):
    """``losses.py:17-34``: holds the trainable ``log_vars``; the weighting itself is fused into the
    loss kernel (``OverallLoss``)."""

    def __init__(self, loss_num, device=None):
        super().__init__()
        self.loss_num = loss_num
        self.log_vars = nn.Parameter(torch.zeros(self.loss_num, ), requires_grad=True)

    def forward(self, loss_list):
        assert len(loss_list) == self.loss_num
        precision = torch.exp(-self.log_vars)
        loss = 0
        for i in range(self.loss_num):
            loss += precision[i] * loss_list[i] + self.log_vars[i]
        return loss


class ICLLoss(nn.Module):
    """``losses.py:36-58``: single-embedding contrastive loss (temperature fixed to 0.1)."""

    def __init__(self, device, temperature=0.05, alpha=0.5):
        super().__init__()
        self.temp = 0.1
        self.alpha = alpha
        self.device = device

    def forward(self, emb, data_dict):
        idx = _index_tensors(data_dict, emb.device, emb.shape[0])
        return ag.OverallLossFn.apply(idx, 0.1, None, None, emb)[0]


class IALLoss(nn.Module):
    """``losses.py:60-97``.  Stand-alone IAL of one (modal, joint) pair from the fused kernel: with log_vars = 0
    and zoom = 1 the kernel's ``ial`` output for M = 1 modality equals this loss, and because the fused loss is
    affine in ``zoom`` its gradient is the difference of the fused gradients at zoom = 1 and zoom = 0
    (``autograd.StandaloneIAL``).  Differentiable w.r.t. both embeddings like the reference's."""

    def __init__(self, device, temperature=0.05, alpha=0.5):
        super().__init__()
        self.temp = 1.0
        self.alpha = alpha
        self.device = device
        self.zoom = 0.1

    def forward(self, src_emb, ref_emb, data_dict):
        idx = _index_tensors(data_dict, src_emb.device, src_emb.shape[0])
        return ag.StandaloneIAL.apply(idx, src_emb, ref_emb)


class OverallLoss(nn.Module):
    """``losses.py:99-152``."""

    def __init__(self, ial_loss_layer, icl_loss_layer, device, metadata):
        super().__init__()
        self.zoom = metadata['zoom']
        self.device = device
        self.modules = metadata['modules']
        self.weight_align_loss = metadata['wt_align_loss']
        self.weight_contrastive_loss = metadata['wt_contrastive_loss']
        self.align_loss = IALLoss(device)
        self.contrastive_loss = ICLLoss(self.device)
        self.align_multi_loss_layer = ial_loss_layer
        self.contrastive_multi_loss_layer = icl_loss_layer

    def forward(self, output_dict, data_dict):
        mods = self.modules
        dev = output_dict[mods[0]].device
        idx = _index_tensors(data_dict, dev, output_dict[mods[0]].shape[0])
        if len(mods) > 1:
            embs = [output_dict[m] for m in mods] + [output_dict['joint']]
            lv_ial = self.align_multi_loss_layer.log_vars
            lv_icl = self.contrastive_multi_loss_layer.log_vars
            losses = ag.OverallLossFn.apply(idx, float(self.zoom), lv_ial, lv_icl, *embs)
            return {'loss': losses[0], 'icl_loss_unimodal': losses[1].detach(), 'icl_loss_multimodal': losses[2].detach(),
                    'ial_loss': losses[3].detach()}
        losses = ag.OverallLossFn.apply(idx, float(self.zoom), None, None, output_dict[mods[0]])
        return {'loss': losses[0], 'icl_loss_unimodal': losses[1].detach(), 'icl_loss_multimodal': 0.0, 'ial_loss': 0.0}


class NCALoss(nn.Module):
    """``losses.py:154-176``.  ``forward(src_emb, ref_emb)`` takes the two already normalised and gathered row sets, as
    the reference does; ``OverallNCALoss`` goes through the fused form (normalise + gather inside the GEMM loaders)."""

    def __init__(self, alpha, beta, ep):
        super().__init__()
        self.alpha = alpha
        self.beta = beta
        self.ep = ep

    def forward(self, src_emb, ref_emb):
        if not src_emb.is_cuda:
            raise RuntimeError('sgaligner_b200.NCALoss needs CUDA tensors (no CPU fallback)')
        n = src_emb.shape[0]
        both = torch.cat([src_emb, ref_emb])
        ar = torch.arange(2 * n, device=src_emb.device, dtype=torch.int32)
        return ag.NCAPairFn.apply(both, ar[:n].contiguous(), ar[n:].contiguous(), float(self.alpha), float(self.beta), float(self.ep))


class OverallNCALoss(nn.Module):
    """``losses.py:178-205``: one NCA term per entry of the output dict (the joint embedding included), summed."""

    def __init__(self, modules, device):
        super().__init__()
        self.device = device
        self.criterion_dict = {module: NCALoss(alpha=1, beta=1, ep=0.0) for module in modules}
        self.criterion_dict['joint'] = NCALoss(alpha=1, beta=1, ep=0.0)

    def forward(self, output_dict, data_dict):
        loss_dict = {}
        first = next(iter(output_dict.values()))
        idx = _index_tensors(data_dict, first.device, first.shape[0])
        for module in output_dict.keys():
            c = self.criterion_dict[module]
            loss_dict[module] = ag.NCAFn.apply(output_dict[module], idx[0], idx[1], float(c.alpha), float(c.beta), float(c.ep))
        loss_sum = 0
        for module in loss_dict.keys():
            loss_sum = loss_sum + loss_dict[module]
        loss_dict['loss'] = loss_sum
        return loss_dict
