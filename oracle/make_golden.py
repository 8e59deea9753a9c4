"""Generate ``tests/golden/*.npz`` by running the UNMODIFIED reference modules (imported from
``/root/reference`` through :mod:`oracle.ref_import`) on seeded inputs, and check the oracle
restatement against them while doing so.  Run in the build container only:

    python -m oracle.make_golden

TEST INFRASTRUCTURE ONLY.  The committed vectors are what the CPU test-suite pins the oracle
with and what the ``-m gpu`` parity tests compare the CUDA path against on the GPU box (where
``/root/reference`` does not exist).
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_import, sgaligner_oracle as O          # noqa: E402
from sgaligner_b200 import synthetic                          # noqa: E402

GOLD = os.path.join(ROOT, 'tests', 'golden')
DATA_KEYS_T = ('tot_obj_pts', 'tot_bow_vec_object_attr_feats', 'tot_bow_vec_object_edge_feats', 'tot_rel_pose', 'edges')
DATA_KEYS_N = ('e1i', 'e2i', 'e1j', 'e2j', 'e1i_count', 'e2i_count', 'e1j_count', 'e2j_count', 'tot_obj_count',
               'graph_per_obj_count', 'graph_per_edge_count')


# ---------------------------------------------------------------------------- example_data (C1)
def _fps(points: np.ndarray, k: int) -> np.ndarray:
    """Farthest-point down-sampling to ``k`` points / re-sampling with replacement when the
    object has fewer (semantics of ``utils/point_cloud.py:61-89``; global numpy RNG)."""
    n = points.shape[0]
    if n < k:
        return points[np.random.choice(n, k)]
    chosen = np.zeros(k, np.int64)
    d = np.full(n, 1e10)
    far = np.random.randint(0, n)
    for t in range(k):
        chosen[t] = far
        d = np.minimum(d, ((points - points[far]) ** 2).sum(-1))
        far = int(np.argmax(d))
    return points[chosen]


def example_pair(n_points: int = 512, min_pts: int = 50, rel_dim: int = 41, attr_dim: int = 164) -> dict:
    """BASELINE.json configs[0]: example_data scene_1 <-> scene_2 as one sub-scan pair.  Only the
    raw point arrays ship with the reference; edges / BoW features / rel_trans are synthesised
    the way ``preprocessing/scan3r/preprocess.py:93-96,164-193`` would shape them (complete
    digraph, rel_trans = root barycentre - object barycentre; barycentre = plain mean here)."""
    np.random.seed(42)
    rng = np.random.default_rng(42)
    scenes = []
    for s in ('scene_1', 'scene_2'):
        d = np.load(os.path.join(ref_import.REFERENCE_ROOT, 'example_data', s, 'data.npy'))
        xyz = np.stack([d['x'], d['y'], d['z']], 1)
        ids = [int(i) for i in np.unique(d['objectId']) if i != 0 and (d['objectId'] == i).sum() >= min_pts]
        objs = {i: xyz[d['objectId'] == i] for i in ids}
        scenes.append((ids, objs, xyz))
    center = scenes[0][2].mean(0)          # scan3r.py:76 (val split: source centre)
    shared = [i for i in scenes[0][0] if i in scenes[1][0]]
    pts, pose, attr, rel, edges, ocnt, ecnt = [], [], [], [], [], [], []
    feats = {}
    for ids, objs, _ in scenes:
        bary = np.stack([objs[i].mean(0) for i in ids])
        pts.append(np.stack([_fps(objs[i], n_points) for i in ids]).astype(np.float32) - center.astype(np.float32))
        pose.append(bary[0][None] - bary)
        a = np.zeros((len(ids), attr_dim)); r = np.zeros((len(ids), rel_dim))
        for k, i in enumerate(ids):
            if i not in feats:
                feats[i] = ((rng.random(attr_dim) < 0.05).astype(np.float64), rng.poisson(1.0, rel_dim).astype(np.float64))
            a[k], r[k] = feats[i]
        attr.append(a); rel.append(r)
        n = len(ids)
        s, o = np.meshgrid(np.arange(n), np.arange(n), indexing='ij')
        edges.append(np.stack([s[s != o], o[s != o]], 1).astype(np.int64))
        ocnt.append(n); ecnt.append(edges[-1].shape[0])
    ids0, ids1 = scenes[0][0], scenes[1][0]
    ns = len(ids0)
    e1i = np.array([ids0.index(i) for i in shared]); e2i = np.array([ids1.index(i) for i in shared]) + ns
    e1j = np.array([k for k, i in enumerate(ids0) if i not in shared]); e2j = np.array([k for k, i in enumerate(ids1) if i not in shared]) + ns
    cat = np.concatenate
    return {
        'tot_obj_pts': torch.from_numpy(cat(pts)), 'tot_bow_vec_object_attr_feats': torch.from_numpy(cat(attr)),
        'tot_bow_vec_object_edge_feats': torch.from_numpy(cat(rel)), 'tot_rel_pose': torch.from_numpy(cat(pose)),
        'edges': torch.from_numpy(cat(edges)),
        'e1i': e1i.astype(np.int32), 'e2i': e2i.astype(np.int32), 'e1j': e1j.astype(np.int32), 'e2j': e2j.astype(np.int32),
        'e1i_count': np.array([len(e1i)]), 'e2i_count': np.array([len(e2i)]), 'e1j_count': np.array([len(e1j)]),
        'e2j_count': np.array([len(e2j)]), 'tot_obj_count': np.array([sum(ocnt)]),
        'graph_per_obj_count': np.array([ocnt]), 'graph_per_edge_count': np.array([ecnt]),
        'global_obj_ids': np.array(ids0 + ids1), 'obj_ids': np.array(ids0 + ids1),
        'scene_ids': np.array([['scene_1', 'scene_2']]), 'pcl_center': center[None], 'overlap': np.array([-1.0]),
        'batch_size': 1,
    }


# ---------------------------------------------------------------------------- cases
def _messy_edges(data: dict, seed: int) -> dict:
    """Inject self loops and duplicated edges into every graph (GATConv must drop the former
    and count the latter separately)."""
    rng = np.random.default_rng(seed)
    oc = np.asarray(data['graph_per_obj_count']).reshape(-1)
    ec = np.asarray(data['graph_per_edge_count']).reshape(-1)
    ed = data['edges'].numpy()
    out, cnt, e0 = [], [], 0
    for n, e in zip(oc, ec):
        g = ed[e0:e0 + e]; e0 += e
        loops = np.stack([rng.integers(0, n, 2)] * 2, 1)
        dup = g[rng.integers(0, max(1, e), 3)] if e else np.zeros((0, 2), np.int64)
        g2 = np.concatenate([g, loops, dup])
        g2 = g2[rng.permutation(len(g2))]
        out.append(g2); cnt.append(len(g2))
    data = dict(data)
    data['edges'] = torch.from_numpy(np.concatenate(out).astype(np.int64))
    data['graph_per_edge_count'] = np.array(cnt).reshape(-1, 2)
    return data


def cases():
    yield 'small4', synthetic.make_batch([7, 6, 9], [5, 9, 4], [4, 5, 3], [3, 2, 3], n_points=64, edge_mode='complete', seed=11), \
        dict(modules=['point', 'gat', 'rel', 'attr'], seed=3)
    yield 'messy_pg', _messy_edges(synthetic.make_batch([12, 10], [11, 13], [6, 7], n_points=128, edge_mode='kout', k_out=3, seed=12), 5), \
        dict(modules=['point', 'gat'], seed=4)
    yield 'point_only', synthetic.make_batch([8, 5], [6, 7], [4, 4], n_points=96, edge_mode='kout', k_out=2, seed=13), \
        dict(modules=['point'], seed=5)
    yield 'mid4', synthetic.make_batch([20, 17, 25, 9], [18, 22, 16, 12], [10, 9, 12, 5], [6, 9, 4, 5], n_points=256, edge_mode='complete', seed=14), \
        dict(modules=['gat', 'point', 'attr', 'rel'], seed=6)
    yield 'c1_example', example_pair(), dict(modules=['point', 'gat', 'rel', 'attr'], seed=42)


def run_reference(data: dict, modules, seed: int):
    sg, ls, al = ref_import.load_reference()
    torch.manual_seed(seed)
    model = sg.MultiModalEncoder(modules=modules, rel_dim=41, attr_dim=164)
    M = len(modules)
    lv_ial = ls.CustomMultiLossLayer(M); lv_icl = ls.CustomMultiLossLayer(M)
    with torch.no_grad():
        lv_ial.log_vars.copy_(0.3 * torch.randn(M)); lv_icl.log_vars.copy_(0.3 * torch.randn(M))
        model.fusion.weight.copy_(1 + 0.5 * torch.randn(M, 1))
        for n_, p_ in model.named_parameters():          # non-zero biases so that they are exercised
            if n_.endswith('bias') and ('conv' in n_ or 'layer_stack' in n_):
                p_.copy_(0.1 * torch.randn_like(p_))
    params0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
    loss_fn = ls.OverallLoss(lv_ial, lv_icl, 'cpu', {'zoom': 0.1, 'wt_align_loss': 1.0, 'wt_contrastive_loss': 1.0, 'modules': modules})
    model.train()
    out = model(data)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        ld = loss_fn(out, data)
    ld['loss'].backward()
    grads = {n_: p_.grad.detach().clone() for n_, p_ in model.named_parameters() if p_.grad is not None}
    grads['__lv_ial'] = lv_ial.log_vars.grad.clone() if lv_ial.log_vars.grad is not None else torch.zeros(M)
    grads['__lv_icl'] = lv_icl.log_vars.grad.clone() if lv_icl.log_vars.grad is not None else torch.zeros(M)
    bn_after = {k: v.detach().clone() for k, v in model.state_dict().items() if 'running' in k or 'num_batches' in k}
    # matching head exactly as inference_align_reg.py:125-128 + alignment.py
    emb = (out['joint'] if M > 1 else out[modules[0]]).detach()
    offs = O.pair_offsets(data)
    ranks, sims, hits, rr = [], [], {k: 0 for k in range(1, 6)}, []
    a0 = 0
    for b in range(data['batch_size']):
        o0, o1 = int(offs[b]), int(offs[b + 1]); na = int(data['e1i_count'][b])
        e1 = data['e1i'][a0:a0 + na] - o0; e2 = data['e2i'][a0:a0 + na] - o0; a0 += na
        e = emb[o0:o1]; e = e / e.norm(dim=1)[:, None]
        sim = 1 - torch.mm(e, e.transpose(0, 1)); rank = torch.argsort(sim, dim=1)
        rr = al.compute_mean_reciprocal_rank(rank, e1, e2, rr)
        for k in hits:
            hits[k] += al.compute_hits_k(rank, e1, e2, k)[0]
        ranks.append(rank.numpy()); sims.append(sim.numpy())
    return params0, lv_ial.log_vars.detach().clone(), lv_icl.log_vars.detach().clone(), out, ld, grads, bn_after, ranks, sims, hits, rr


def rel_err(a, b):
    a = torch.as_tensor(a, dtype=torch.float64); b = torch.as_tensor(b, dtype=torch.float64)
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def main():
    os.makedirs(GOLD, exist_ok=True)
    for name, data, cfg in cases():
        modules = cfg['modules']
        params, lvi, lvc, out, ld, grads, bn_after, ranks, sims, hits, rr = run_reference(data, modules, cfg['seed'])
        # ---- pin the restatement against the reference itself
        o_out = O.encoder_forward(params, data, modules)
        worst = max(rel_err(o_out[k], out[k]) for k in out)
        o_ld = O.overall_loss(o_out, data, modules, lvi, lvc)
        worst_l = max(rel_err(o_ld[k], ld[k]) for k in ld if torch.is_tensor(ld[k]))
        ev = O.evaluate_batch(o_out['joint'] if len(modules) > 1 else o_out[modules[0]], data)
        assert worst < 2e-6 and worst_l < 2e-5, (name, worst, worst_l)
        assert ev['hits'] == hits and abs(ev['mrr'] - float(np.mean(rr))) < 1e-12, (name, ev['hits'], hits)
        print(f'{name:12s} oracle-vs-reference: emb {worst:.2e} loss {worst_l:.2e} hits {hits} mrr {np.mean(rr):.4f}')
        blob = {}
        for k in DATA_KEYS_T:
            blob['in/' + k] = data[k].numpy()
        for k in DATA_KEYS_N:
            blob['in/' + k] = np.asarray(data[k])
        blob['in/batch_size'] = np.array(data['batch_size'])
        blob['cfg/modules'] = np.array(modules)
        for k, v in params.items():
            blob['p/' + k] = v.numpy()
        blob['p/__lv_ial'] = lvi.numpy(); blob['p/__lv_icl'] = lvc.numpy()
        for k, v in out.items():
            blob['out/' + k] = v.detach().numpy()
        for k, v in ld.items():
            blob['loss/' + k] = np.array(float(v))
        for k, v in grads.items():
            blob['grad/' + k] = v.numpy()
        for k, v in bn_after.items():
            blob['bn/' + k] = v.numpy()
        for b, (r, s) in enumerate(zip(ranks, sims)):
            blob[f'rank/{b}'] = r.astype(np.int16); blob[f'sim/{b}'] = s.astype(np.float32)
        blob['metric/hits'] = np.array([hits[k] for k in range(1, 6)]); blob['metric/rr'] = np.array(rr)
        np.savez_compressed(os.path.join(GOLD, name + '.npz'), **blob)
        print('   wrote', name, os.path.getsize(os.path.join(GOLD, name + '.npz')) // 1024, 'KiB')


if __name__ == '__main__':
    main()
