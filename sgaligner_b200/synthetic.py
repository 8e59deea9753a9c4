"""Synthetic Scan3R-shaped batches that obey the collate contract of the reference dataloader
(``src/datasets/scan3r.py:60-209``): flat "stack mode" tensors for all objects of all pairs plus
per-pair host-side counts.  Used by ``bench.py``, the tests and the golden generator; there is
no dataset on the GPU box.

Workloads:
  C2  B=32, 64+64 objects, 512 pts, 6 out-edges/node        -> :func:`config_c2`
  C3  B=128, 3RScan-shaped (n~N(28,10) clipped, complete digraph) -> :func:`config_c3`
  C5  256+256 objects, 1024 pts                              -> :func:`config_c5`
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch


def make_batch(n_src: Sequence[int], n_ref: Sequence[int], n_shared: Sequence[int],
               n_anchor: Optional[Sequence[int]] = None, n_points: int = 512,
               edge_mode: str = 'kout', k_out: int = 6, rel_dim: int = 41, attr_dim: int = 164,
               jitter: float = 0.01, seed: int = 0) -> dict:
    """Build one collated batch.

    Pair ``b`` has ``n_src[b]`` source and ``n_ref[b]`` reference objects; the first
    ``n_shared[b]`` objects of both graphs are the same physical objects (reference copy =
    source points + N(0, jitter)); the first ``n_anchor[b]`` of those are the anchors
    (``e1i``/``e2i``); every other object lands in ``e1j``/``e2j`` (``scan3r.py:101-107``).
    Node order inside a pair is source objects then reference objects (``scan3r.py:109``);
    edges are graph-local ``[E,2]`` int64 rows ``(src, dst)``, source graph first
    (``scan3r.py:99,201``).
    """
    rng = np.random.default_rng(seed)
    B = len(n_src)
    if n_anchor is None:
        n_anchor = n_shared
    pts, rel, attr, pose, edges = [], [], [], [], []
    e1i, e2i, e1j, e2j = [], [], [], []
    obj_cnt, edge_cnt = [], []
    base = 0
    for b in range(B):
        ns, nr, sh, na = int(n_src[b]), int(n_ref[b]), int(n_shared[b]), int(n_anchor[b])
        assert na <= sh <= min(ns, nr)
        p_src = (rng.standard_normal((ns, n_points, 3)) + rng.uniform(-2, 2, (ns, 1, 3))).astype(np.float32)
        p_ref = (rng.standard_normal((nr, n_points, 3)) + rng.uniform(-2, 2, (nr, 1, 3))).astype(np.float32)
        p_ref[:sh] = p_src[:sh] + (jitter * rng.standard_normal((sh, n_points, 3))).astype(np.float32)
        pts += [p_src, p_ref]
        a_src = (rng.random((ns, attr_dim)) < 0.05).astype(np.float64)
        a_ref = (rng.random((nr, attr_dim)) < 0.05).astype(np.float64)
        a_ref[:sh] = a_src[:sh]
        attr += [a_src, a_ref]
        r_src = rng.poisson(1.0, (ns, rel_dim)).astype(np.float64)
        r_ref = rng.poisson(1.0, (nr, rel_dim)).astype(np.float64)
        r_ref[:sh] = r_src[:sh]
        rel += [r_src, r_ref]
        t_src = rng.standard_normal((ns, 3))
        t_ref = rng.standard_normal((nr, 3))
        t_ref[:sh] = t_src[:sh] + jitter * rng.standard_normal((sh, 3))
        pose += [t_src, t_ref]
        cnt_e = []
        for n in (ns, nr):
            if edge_mode == 'complete':
                s, d = np.meshgrid(np.arange(n), np.arange(n), indexing='ij')
                keep = s != d
                ed = np.stack([s[keep], d[keep]], 1)
            elif edge_mode == 'kout':
                if n > 1:
                    s = np.repeat(np.arange(n), k_out)
                    d = (s + rng.integers(1, n, s.shape[0])) % n
                    ed = np.stack([s, d], 1)
                else:
                    ed = np.zeros((0, 2), np.int64)
            else:
                raise ValueError(edge_mode)
            edges.append(ed.astype(np.int64))
            cnt_e.append(ed.shape[0])
        edge_cnt.append(cnt_e)
        obj_cnt.append([ns, nr])
        e1i.append(base + np.arange(na))
        e2i.append(base + ns + np.arange(na))
        e1j.append(base + np.arange(na, ns))
        e2j.append(base + ns + np.arange(na, nr))
        base += ns + nr
    cat = np.concatenate
    i32 = lambda xs: cat(xs).astype(np.int32)
    data = {
        'tot_obj_pts': torch.from_numpy(cat(pts)),
        'tot_bow_vec_object_attr_feats': torch.from_numpy(cat(attr)),
        'tot_bow_vec_object_edge_feats': torch.from_numpy(cat(rel)),
        'tot_rel_pose': torch.from_numpy(cat(pose)),
        'edges': torch.from_numpy(cat(edges)),
        'e1i': i32(e1i), 'e2i': i32(e2i), 'e1j': i32(e1j), 'e2j': i32(e2j),
        'e1i_count': np.array([len(x) for x in e1i]), 'e2i_count': np.array([len(x) for x in e2i]),
        'e1j_count': np.array([len(x) for x in e1j]), 'e2j_count': np.array([len(x) for x in e2j]),
        'tot_obj_count': np.array([a + b for a, b in obj_cnt]),
        'graph_per_obj_count': np.array(obj_cnt),
        'graph_per_edge_count': np.array(edge_cnt),
        'global_obj_ids': np.zeros(base, np.int64),
        'obj_ids': cat([cat([np.arange(1, a + 1), np.arange(1, b + 1)]) for a, b in obj_cnt]),
        'scene_ids': np.array([[f'syn{b}_src', f'syn{b}_ref'] for b in range(B)]),
        'pcl_center': np.zeros((B, 3)),
        'overlap': np.full(B, -1.0),
        'batch_size': B,
    }
    return data


def config_c2(batch: int = 32, seed: int = 0, n_obj: int = 64, n_points: int = 512) -> dict:
    """BASELINE.json configs[1]: 64 objects/scene, 512 points/object, ~6 edges/node."""
    n = [n_obj] * batch
    return make_batch(n, n, [n_obj // 2] * batch, n_points=n_points, edge_mode='kout', k_out=6, seed=seed)


def config_c3(batch: int = 128, seed: int = 1, train: bool = True, n_points: int = 512) -> dict:
    """BASELINE.json configs[2]: 3RScan-shaped sub-scans (complete digraph edges as the real
    preprocessor emits, ``preprocessing/scan3r/preprocess.py:176-182``)."""
    rng = np.random.default_rng(seed + 1000)
    ns = np.clip(np.round(rng.normal(28, 10, batch)), 8, 64).astype(int)
    nr = np.clip(np.round(rng.normal(28, 10, batch)), 8, 64).astype(int)
    sh = np.maximum(2, (np.minimum(ns, nr) * rng.uniform(0.3, 0.9, batch)).astype(int))
    if train:   # scan3r.py:89-91: 30 % of the shared objects, at least 2
        na = np.maximum(2, (0.3 * sh).astype(int))
    else:
        na = sh
    return make_batch(ns, nr, sh, na, n_points=n_points, edge_mode='complete', seed=seed)


def config_c5(batch: int = 8, seed: int = 2) -> dict:
    """BASELINE.json configs[4]: 256 objects/scene, 1024 points/object."""
    n = [256] * batch
    return make_batch(n, n, [128] * batch, n_points=1024, edge_mode='kout', k_out=6, seed=seed)


def shard_batch(data: dict, rank: int, world: int) -> dict:
    """Contiguous split of the pairs of a collated batch across ``world`` ranks (never splits a
    pair).  Index arrays are re-based to the shard's first object."""
    B = int(data['batch_size'])
    if B < world:
        raise ValueError(f'cannot shard {B} pairs over {world} ranks: every rank needs at least one pair '
                         '(an empty shard would fail in the loss while the others wait in the all-reduce)')
    # balanced split (np.array_split boundaries): the first B % world ranks get one pair more
    base, extra = divmod(B, world)
    b0 = rank * base + min(rank, extra)
    b1 = b0 + base + (1 if rank < extra else 0)
    return slice_pairs(data, b0, b1)


def slice_pairs(data: dict, b0: int, b1: int) -> dict:
    oc = np.asarray(data['graph_per_obj_count']).reshape(-1, 2)
    ec = np.asarray(data['graph_per_edge_count']).reshape(-1, 2)
    o_off = np.concatenate([[0], np.cumsum(oc.sum(1))])
    e_off = np.concatenate([[0], np.cumsum(ec.sum(1))])
    o0, o1 = int(o_off[b0]), int(o_off[b1])
    e0, e1 = int(e_off[b0]), int(e_off[b1])
    out = {}
    for k in ('tot_obj_pts', 'tot_bow_vec_object_attr_feats', 'tot_bow_vec_object_edge_feats', 'tot_rel_pose'):
        out[k] = data[k][o0:o1]
    out['edges'] = data['edges'][e0:e1]
    for k in ('e1i', 'e2i', 'e1j', 'e2j'):
        c = np.asarray(data[k + '_count'])
        off = np.concatenate([[0], np.cumsum(c)])
        out[k] = (np.asarray(data[k])[off[b0]:off[b1]] - o0).astype(np.int32)
        out[k + '_count'] = c[b0:b1]
    for k in ('tot_obj_count', 'graph_per_obj_count', 'graph_per_edge_count', 'scene_ids', 'pcl_center', 'overlap'):
        out[k] = np.asarray(data[k])[b0:b1]
    for k in ('global_obj_ids', 'obj_ids'):
        out[k] = np.asarray(data[k])[o0:o1]
    out['batch_size'] = b1 - b0
    return out
