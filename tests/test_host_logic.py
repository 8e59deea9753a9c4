"""CPU: host-side logic -- synthetic batches obey the collate contract, sharding, module surface,
state_dict keys, rank-metric helpers vs the oracle."""
import numpy as np
import pytest
import torch

from oracle import sgaligner_oracle as O
from sgaligner_b200 import matching, synthetic
from tests.util import load_case


def test_synthetic_batch_contract():
    d = synthetic.make_batch([5, 7], [6, 4], [3, 2], [2, 2], n_points=16, edge_mode='complete', seed=0)
    N = 5 + 6 + 7 + 4
    assert d['tot_obj_pts'].shape == (N, 16, 3) and d['tot_obj_pts'].dtype == torch.float32
    assert d['tot_bow_vec_object_attr_feats'].dtype == torch.float64 and d['tot_bow_vec_object_attr_feats'].shape == (N, 164)
    assert d['tot_bow_vec_object_edge_feats'].shape == (N, 41) and d['tot_rel_pose'].shape == (N, 3)
    assert d['edges'].dtype == torch.int64 and d['edges'].shape[1] == 2
    assert d['edges'].shape[0] == int(np.asarray(d['graph_per_edge_count']).sum())
    assert d['e1i'].dtype == np.int32
    # every node appears in exactly one of the four index sets (scan3r.py:101-107)
    allidx = np.concatenate([d['e1i'], d['e2i'], d['e1j'], d['e2j']])
    assert sorted(allidx.tolist()) == list(range(N))
    # edges are graph-local
    oc = np.asarray(d['graph_per_obj_count']).reshape(-1)
    ec = np.asarray(d['graph_per_edge_count']).reshape(-1)
    e0 = 0
    for n, e in zip(oc, ec):
        assert int(d['edges'][e0:e0 + e].max()) < n
        e0 += e


def test_shard_batch_is_a_partition():
    d = synthetic.make_batch([5, 7, 3, 6], [6, 4, 5, 5], [3, 2, 2, 4], n_points=8, seed=1)
    parts = [synthetic.shard_batch(d, r, 2) for r in range(2)]
    assert sum(p['batch_size'] for p in parts) == 4
    assert torch.equal(torch.cat([p['tot_obj_pts'] for p in parts]), d['tot_obj_pts'])
    assert torch.equal(torch.cat([p['edges'] for p in parts]), d['edges'])
    off = int(parts[0]['tot_obj_pts'].shape[0])
    assert np.array_equal(np.concatenate([parts[0]['e1i'], parts[1]['e1i'] + off]), d['e1i'])
    # a shard evaluated alone equals the same pairs evaluated inside the big batch (pairs are independent
    # through the encoder and the matching head)
    p = O.init_params(['point', 'gat'], 41, 164, seed=0)
    with torch.no_grad():
        full = O.encoder_forward(p, d, ['point', 'gat'])['joint']
        part = O.encoder_forward(p, parts[1], ['point', 'gat'])['joint']
    assert torch.allclose(full[off:], part, atol=1e-6)


def test_module_surface_and_state_dict_keys():
    from sgaligner_b200 import losses, sg_aligner
    c = load_case('small4')
    m = sg_aligner.MultiModalEncoder(modules=['point', 'gat', 'rel', 'attr'], rel_dim=41, attr_dim=164)
    assert set(m.state_dict().keys()) == set(c['params'].keys())        # reference checkpoint key set
    for k, v in m.state_dict().items():
        assert tuple(v.shape) == tuple(c['params'][k].shape), k
    assert sum(p.numel() for p in m.parameters()) == 182440
    assert isinstance(m.modules, list)                                 # reference quirk kept (sg_aligner.py:42)
    for name in ('ProjectionHead', 'MultiModalFusion', 'MultiModalEncoder', 'torch', 'nn', 'F'):
        assert hasattr(sg_aligner, name)
    for name in ('CustomMultiLossLayer', 'ICLLoss', 'IALLoss', 'OverallLoss', 'calculate_prob_dist', 'torch', 'nn', 'F'):
        assert hasattr(losses, name)
    with pytest.raises(NotImplementedError):
        sg_aligner.MultiModalEncoder(modules=['gat'], rel_dim=41, attr_dim=164)
    with pytest.raises(RuntimeError):                                   # CPU tensors are refused, not silently computed
        m(dict(c['data']))
    lvl = losses.CustomMultiLossLayer(3)
    assert float(lvl([torch.tensor(1.0), torch.tensor(2.0), torch.tensor(3.0)])) == 6.0


def test_compat_shim_resolves_reference_imports():
    """`from aligner.sg_aligner import *` / `from aligner.losses import *` (trainval_sgaligner.py:11-12)."""
    import importlib
    import os
    import sys
    from tests.util import ROOT
    shim = os.path.join(ROOT, 'sgaligner_b200', 'compat')
    sys.path.insert(0, shim)
    try:
        for k in [k for k in sys.modules if k == 'aligner' or k.startswith('aligner.')]:
            del sys.modules[k]
        sg = importlib.import_module('aligner.sg_aligner')
        ls = importlib.import_module('aligner.losses')
        assert sg.MultiModalEncoder.__module__.startswith('sgaligner_b200')
        assert ls.OverallLoss.__module__.startswith('sgaligner_b200')
        assert hasattr(sg, 'torch') and hasattr(ls, 'CustomMultiLossLayer')
    finally:
        sys.path.remove(shim)
        for k in [k for k in sys.modules if k == 'aligner' or k.startswith('aligner.')]:
            del sys.modules[k]


def test_alignment_helpers_match_oracle():
    c = load_case('messy_pg')
    for b, (rank, sim) in enumerate(zip(c['rank'], c['sim'])):
        n = rank.shape[0]
        ns = int(c['data']['graph_per_obj_count'][b][0])
        e1 = np.arange(min(3, ns))
        e2 = ns + np.arange(min(3, ns))
        hits, rr = O.hits_and_rr(rank, e1, e2)
        for k in (1, 3, 5):
            assert matching.compute_hits_k(torch.from_numpy(rank), e1, e2, k)[0] == hits[k]
        assert matching.compute_mean_reciprocal_rank(torch.from_numpy(rank), e1, e2, []) == rr
        assert matching.compute_sgar(torch.from_numpy(sim), torch.from_numpy(rank), e1, e2, ['2', '50', '100']) == O.sgar(sim, rank, e1, e2)
        assert matching.compute_node_corrs(torch.from_numpy(rank), ns, 2) == O.node_corrs(rank, ns, 2)
        assert matching.compute_alignment_score(torch.from_numpy(rank), ns, n - ns) == O.alignment_score(rank, ns, n - ns)


def test_flat_adam_views_on_cpu():
    from sgaligner_b200.trainer import FlatAdam
    ps = [torch.nn.Parameter(torch.randn(3, 5)), torch.nn.Parameter(torch.randn(7))]
    before = [p.detach().clone() for p in ps]
    opt = FlatAdam(ps)
    for p, b in zip(ps, before):
        assert torch.equal(p.detach(), b)
        assert p.data_ptr() >= opt.flat_param.data_ptr()
    (ps[0].sum() * 2 + ps[1].sum()).backward()
    assert float(opt.flat_grad[:15].sum()) == 30.0 and float(opt.flat_grad[64:71].sum()) == 7.0
    opt.zero_grad()
    assert float(opt.flat_grad.abs().sum()) == 0.0


def test_loss_host_side_accounting_and_partition_check():
    """Host-only entry points of the loss (no GPU): launch accounting, workspace sizing, and the partition check that
    decides between the packed-image Gram kernel and the gathering GEMM."""
    import ctypes
    from sgaligner_b200 import _lib
    from sgaligner_b200.losses import _index_tensors
    lib = _lib.get_lib()
    dims = (ctypes.c_int * 3)(100, 100, 200)
    fwd = lib.sga_loss_launch_count(3, dims, 10, 10, 0)
    both = lib.sga_loss_launch_count(3, dims, 10, 10, 1)
    # forward: ridx + finalize + 3 pack + 1 pair + slots + narrow Gram group + wide Gram group
    assert fwd == 9 and both > fwd
    one = (ctypes.c_int * 1)(100)
    assert lib.sga_loss_launch_count(1, one, 0, 0, 0) < fwd
    w_small = lib.sga_loss_workspace_bytes(3, dims, 4096, 1024, 1024, 1024, 0)
    w_grad = lib.sga_loss_workspace_bytes(3, dims, 4096, 1024, 1024, 1024, 1)
    assert 0 < w_small < w_grad
    # two [A, T] fp32 similarity blocks per embedding dominate
    assert w_small >= 3 * 2 * 1024 * 3072 * 4
    d = {'e1i': np.array([0, 1]), 'e2i': np.array([4, 5]), 'e1j': np.array([2, 3]), 'e2j': np.array([6, 7])}
    assert _index_tensors(dict(d), torch.device('cpu')).partition
    d['e1j'] = np.array([2, 1])          # node 1 is both an anchor and a non-anchor
    assert not _index_tensors(dict(d), torch.device('cpu')).partition
    d['e1j'] = np.array([2, 2])          # repeated node
    assert not _index_tensors(dict(d), torch.device('cpu')).partition


def test_eva_module_surface_matches_reference_key_set():
    """EVA: constructor signature of eva.py:10, the reference's exact state_dict key set and
    shapes (frozen in tests/golden/eva_ref.npz from the unmodified module), NCA loss classes exported like losses.py,
    and a CPU batch raises instead of falling back."""
    import os
    import numpy as np
    import pytest
    import torch
    from sgaligner_b200.eva import EVA
    from sgaligner_b200 import losses
    from tests.util import GOLD
    z = np.load(os.path.join(GOLD, 'eva_ref.npz'))
    ref = {k[2:]: tuple(z[k].shape) for k in z.files if k.startswith('p/')}
    m = EVA(modules=['gcn', 'point', 'rel', 'attr'], rel_dim=41, attr_dim=164)
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == ref
    m.load_state_dict({k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith('p/')}, strict=True)
    assert m.modules == ['gcn', 'point', 'rel', 'attr'] and m.n_units == [3, 200, 400]
    assert {'NCALoss', 'OverallNCALoss'} <= set(losses.__all__)
    fn = losses.OverallNCALoss(['gcn', 'point'], 'cpu')
    assert set(fn.criterion_dict) == {'gcn', 'point', 'joint'} and fn.criterion_dict['joint'].alpha == 1
    with pytest.raises(RuntimeError):
        m({'tot_obj_pts': torch.zeros(2, 8, 3)})
