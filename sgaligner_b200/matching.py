"""Matching head + rank metrics.

Mirrors the per-pair loop of ``src/inference/sgaligner/inference_align_reg.py:107-145`` and the
helpers of ``utils/alignment.py`` (same function names / return values), but the similarity
matrices, rankings and anchor positions of ALL pairs of a batch are produced by three kernel
launches; the host only aggregates integer counts.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np
import torch

from . import ops


# ------------------------------------------------------------------------------ device path
def match_batch(embedding: torch.Tensor, data_dict: dict, k: int = 6, full_rank: bool = False, want_sim: bool = True,
                tensor_cores: bool = True, layout: 'ops.PairLayout' = None) -> dict:
    """For every node: its ``k`` best matches among the source+reference nodes of its own pair
    (pair-local indices, the node itself included, exactly like ``rank_list[:, :k]`` of the
    reference).  ``full_rank`` additionally returns the whole ``rank_list`` of every pair.
    Default path: one fused tcgen05 kernel (Gram + top-k, ``k <= 8``); ``tensor_cores=False`` selects
    the fp32 FMA kernels (any k)."""
    lay = layout if layout is not None else ops.PairLayout(np.asarray(data_dict['graph_per_obj_count']), embedding.device)
    if tensor_cores and k <= 8:
        topk_idx, topk_dist, sim = ops.match_topk_tc(embedding.detach(), lay, k, want_sim or full_rank)
        rank = ops.match_rank(sim, lay, 0, True)[2] if full_rank else None
    else:
        sim = ops.match_sim(embedding.detach(), lay)
        topk_idx, topk_dist, rank = ops.match_rank(sim, lay, k, full_rank)
    return {'layout': lay, 'sim': sim, 'topk_idx': topk_idx, 'topk_dist': topk_dist, 'rank': rank}


def rank_lists(result: dict) -> List[torch.Tensor]:
    """Per-pair ``rank_list`` tensors (int64 [n_b, n_b]) as ``torch.argsort(sim, dim=1)`` yields."""
    lay = result['layout']
    out = []
    for b in range(lay.B):
        n = int(lay.n[b])
        o = int(lay.sim_off_host[b])
        out.append(result['rank'][o:o + n * n].view(n, n).long())
    return out


def evaluate_batch(embedding: torch.Tensor, data_dict: dict, ks: Sequence[int] = (1, 2, 3, 4, 5)) -> dict:
    """Hits@k and MRR of a batch without materialising any ranking on the host
    (``alignment.compute_hits_k`` / ``compute_mean_reciprocal_rank`` semantics)."""
    dev = embedding.device
    lay = ops.PairLayout(np.asarray(data_dict['graph_per_obj_count']), dev)
    sim = ops.match_sim(embedding.detach(), lay)
    e1 = torch.as_tensor(np.asarray(data_dict['e1i']).astype(np.int32)).to(dev)
    e2 = torch.as_tensor(np.asarray(data_dict['e2i']).astype(np.int32)).to(dev)
    pos = ops.match_anchor_pos(sim, lay, e1, e2).cpu().numpy()
    hits = {int(k_): int((pos < k_).sum()) for k_ in ks}
    rr = 1.0 / (pos.astype(np.float64) + 1.0)
    return {'hits': hits, 'total': int(pos.shape[0]), 'mrr': float(rr.mean()) if pos.size else 0.0, 'rr': rr, 'pos': pos}


def evaluate_pairs(embedding: torch.Tensor, data_dict: dict, ks: Sequence[int] = (1, 2, 3, 4, 5),
                   recall_modes: Sequence[str] = ('2', '50', '100'), tensor_cores: bool = True) -> dict:
    """Everything the per-pair loop of ``inference_align_reg.py:107-145`` derives from the embedding -- Hits@k,
    MRR, SGAR per recall mode, the alignment score and the top-1 node correspondences (``compute_node_corrs``
    with ``k = 1``) -- for ALL pairs of the batch: similarity + anchor positions + pair metrics are three
    launches and the results come back in one device-to-host copy (the reference moves ``rank_list`` to the host
    seven times per pair).  Pairs without anchors are skipped for Hits/MRR/SGAR exactly as the reference does."""
    dev = embedding.device
    goc = np.asarray(data_dict['graph_per_obj_count']).reshape(-1, 2)
    lay = ops.PairLayout(goc, dev)
    if tensor_cores:
        sim = ops.match_topk_tc(embedding.detach(), lay, 1, True)[2]
    else:
        sim = ops.match_sim(embedding.detach(), lay)
    e1c = np.asarray(data_dict['e1i_count']).reshape(-1).astype(np.int64)
    aoff_h = np.concatenate([[0], np.cumsum(e1c)]).astype(np.int32)
    e1 = torch.as_tensor(np.asarray(data_dict['e1i']).astype(np.int32)).to(dev, non_blocking=True)
    e2 = torch.as_tensor(np.asarray(data_dict['e2i']).astype(np.int32)).to(dev, non_blocking=True)
    n_src = torch.as_tensor(goc[:, 0].astype(np.int32)).to(dev, non_blocking=True)
    aoff = torch.as_tensor(aoff_h).to(dev, non_blocking=True)
    pos = ops.match_anchor_pos(sim, lay, e1, e2)
    top1, dist, pout = ops.match_pair_metrics(sim, lay, n_src, e1, e2, aoff)
    # one transfer: [pos | top1 | pair_out bits]
    packed = torch.cat([pos, top1, pout.reshape(-1).view(torch.int32)]).cpu().numpy()
    A = int(pos.numel())
    pos_h = packed[:A]
    top1_h = packed[A:A + lay.N]
    pout_h = packed[A + lay.N:].view(np.float32).reshape(lay.B, 4)
    has_anchor = e1c > 0
    hits = {int(k_): int((pos_h < k_).sum()) for k_ in ks}
    rr = 1.0 / (pos_h.astype(np.float64) + 1.0)
    mode_col = {'2': 0, '50': 1}
    sgar = {m: [float(v) for v in pout_h[has_anchor, mode_col.get(m, 2)]] for m in recall_modes}
    corrs = []
    for b in range(lay.B):
        o0, ns = int(lay.pair_off_host[b]), int(goc[b, 0])
        t = top1_h[o0:o0 + ns]
        src = np.nonzero(t >= ns)[0]
        corrs.append([(int(i), int(t[i])) for i in src])
    return {'hits': hits, 'total': A, 'mrr': float(rr.mean()) if A else 0.0, 'rr': rr, 'pos': pos_h, 'sgar': sgar,
            'alignment_score': [float(v) for v in pout_h[:, 3]], 'node_corrs': corrs, 'anchors': [int(v) for v in e1c[has_anchor]],
            'top1': top1_h}


# ------------------------------------------------------------------------------ utils/alignment.py API
def _np(x):
    return x.detach().cpu().numpy() if torch.is_tensor(x) else np.asarray(x)


def _row_without_self(rank_list: np.ndarray, i: int) -> np.ndarray:
    r = rank_list[i]
    return r[r != i]


def compute_mean_reciprocal_rank(rank_list, e1i_idxs, e2i_idxs, mrr_arr):
    """``utils/alignment.py:3-11``."""
    rank_list = _np(rank_list)
    for idx, e1 in enumerate(e1i_idxs):
        row = _row_without_self(rank_list, int(e1))
        mrr_arr.append(1.0 / (int(np.nonzero(row == e2i_idxs[idx])[0][0]) + 1))
    return mrr_arr


def compute_hits_k(rank_list, e1i_idxs, e2i_idxs, k=1):
    """``utils/alignment.py:13-25``."""
    rank_list = _np(rank_list)
    correct = 0
    for idx, e1 in enumerate(e1i_idxs):
        if e2i_idxs[idx] in _row_without_self(rank_list, int(e1))[:k]:
            correct += 1
    return correct, e1i_idxs.shape[0]


def compute_sgar(sim, rank_list, e1i_idxs, e2i_idxs, modes):
    """``utils/alignment.py:27-58``."""
    rank_list, sim = _np(rank_list), _np(sim)
    pred, dist = [], []
    for e1 in e1i_idxs:
        row = _row_without_self(rank_list, int(e1))
        pred.append(int(row[0]))
        dist.append(sim[int(e1)][row[0]])
    order = np.argsort(dist)
    vals = {}
    for mode in modes:
        sel = order[:2] if mode == '2' else (order[:len(order) // 2] if mode == '50' else order)
        vals[mode] = 1.0 if all(pred[i] == e2i_idxs[i] for i in sel) else 0.0
    return vals


def compute_node_corrs(rank_list, src_objects_count, k=1):
    """``utils/alignment.py:60-71``."""
    rank_list = _np(rank_list)
    out = []
    for idx in range(src_objects_count):
        for r in _row_without_self(rank_list, idx)[:k]:
            if r >= src_objects_count:
                out.append((idx, int(r)))
    return out


def get_node_corrs_objects_ids(node_corrs, objects_ids, batch_offset):
    """``utils/alignment.py:73-78``."""
    return [(objects_ids[a + batch_offset], objects_ids[b + batch_offset]) for a, b in node_corrs]


def compute_alignment_score(rank_list, src_objects_count, ref_objects_count):
    """``utils/alignment.py:80-89``."""
    rank_list = _np(rank_list)
    aligned = sum(int(_row_without_self(rank_list, i)[0] >= src_objects_count) for i in range(src_objects_count))
    return aligned / ref_objects_count
