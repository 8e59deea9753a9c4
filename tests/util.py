"""Shared helpers of the test-suite: golden fixture loading, error metrics."""
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, 'tests', 'golden')
CASES = ['small4', 'messy_pg', 'point_only', 'mid4', 'c1_example']

T_KEYS = ('tot_obj_pts', 'tot_bow_vec_object_attr_feats', 'tot_bow_vec_object_edge_feats', 'tot_rel_pose', 'edges')
N_KEYS = ('e1i', 'e2i', 'e1j', 'e2j', 'e1i_count', 'e2i_count', 'e1j_count', 'e2j_count', 'tot_obj_count',
          'graph_per_obj_count', 'graph_per_edge_count')


def load_case(name):
    z = np.load(os.path.join(GOLD, name + '.npz'))
    data = {k: torch.from_numpy(z['in/' + k]) for k in T_KEYS}
    data.update({k: z['in/' + k] for k in N_KEYS})
    data['batch_size'] = int(z['in/batch_size'])
    modules = [str(m) for m in z['cfg/modules']]
    params = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith('p/') and not k.startswith('p/__')}
    lv = (torch.from_numpy(z['p/__lv_ial']), torch.from_numpy(z['p/__lv_icl']))
    out = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith('out/')}
    loss = {k[5:]: float(z[k]) for k in z.files if k.startswith('loss/')}
    grad = {k[5:]: torch.from_numpy(z[k]) for k in z.files if k.startswith('grad/')}
    bn = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith('bn/')}
    rank = [z[f'rank/{b}'].astype(np.int64) for b in range(data['batch_size'])]
    sim = [z[f'sim/{b}'] for b in range(data['batch_size'])]
    return dict(data=data, modules=modules, params=params, lv=lv, out=out, loss=loss, grad=grad, bn=bn, rank=rank, sim=sim,
                hits=z['metric/hits'], rr=z['metric/rr'])


def rel_inf(a, b):
    """||a-b||_inf / ||b||_inf -- the parity metric of BASELINE.json (1e-4 on embeddings)."""
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def grad_close(got, ref, rtol=2e-4, atol=1e-9):
    """max|got-ref| <= rtol*max|ref| + atol (some reference gradients are pure rounding noise, e.g.
    att_dst, which is mathematically almost shift-invariant under the edge softmax)."""
    got = torch.as_tensor(got).double().cpu().reshape(-1)
    ref = torch.as_tensor(ref).double().cpu().reshape(-1)
    return float((got - ref).abs().max()) <= rtol * float(ref.abs().max()) + atol
