"""GPU parity tests (forward): CUDA kernels through the C ABI vs the oracle / committed golden
vectors generated from the unmodified reference (oracle/make_golden.py)."""
import numpy as np
import pytest
import torch

from oracle import sgaligner_oracle as O
from tests.util import CASES, load_case, rel_inf

pytestmark = pytest.mark.gpu

EMB_TOL = 1e-4      # BASELINE.json: 1e-4 rel on embeddings
LOSS_TOL = 1e-3     # 1e-3 on loss


@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available(), 'GPU tests need a CUDA device'
    return torch.device('cuda:0')


def _model(case, dev):
    from sgaligner_b200.sg_aligner import MultiModalEncoder
    m = MultiModalEncoder(modules=case['modules'], rel_dim=41, attr_dim=164)
    missing = m.load_state_dict(case['params'], strict=True)   # exact key set of the reference state_dict
    assert not missing.missing_keys and not missing.unexpected_keys
    return m.to(dev)


def _cuda_data(case, dev):
    from sgaligner_b200 import to_cuda
    return to_cuda(dict(case['data']), dev)


@pytest.mark.parametrize('name', CASES)
def test_pointnet_simt_vs_oracle(name, dev):
    from sgaligner_b200 import ops
    c = load_case(name)
    p = c['params']
    pts = c['data']['tot_obj_pts']
    ref = O.pointnet_feat(pts, p)
    w = [p[f'object_encoder.conv{i}.{k}'].to(dev) for i in (1, 2, 3) for k in ('weight', 'bias')]
    out, arg = ops.pointnet_forward(pts.to(dev), *w, want_argmax=True, mode=ops.POINTNET_SIMT)
    torch.cuda.synchronize()
    assert rel_inf(out, ref) < 1e-5
    # argmax must point at a point that attains the max (checked through the oracle activations)
    h = pts
    for i in (1, 2, 3):
        h = torch.relu(h @ p[f'object_encoder.conv{i}.weight'].squeeze(-1).t() + p[f'object_encoder.conv{i}.bias'])
    picked = torch.gather(h, 1, arg.cpu().long().unsqueeze(1)).squeeze(1)
    assert (picked - ref).abs().max() < 1e-4 * ref.abs().max()
    assert int(arg.min()) >= 0 and int(arg.max()) < pts.shape[1]


@pytest.mark.parametrize('name', [c for c in CASES if c != 'point_only'])
def test_gat_vs_oracle(name, dev):
    from sgaligner_b200 import ops
    from sgaligner_b200.sg_aligner import MultiGAT
    c = load_case(name)
    d = c['data']
    p = c['params']
    outs, o, e = [], 0, 0
    pose = d['tot_rel_pose'].float()
    for b in range(d['batch_size']):
        for s in (0, 1):
            n, ne = int(d['graph_per_obj_count'][b][s]), int(d['graph_per_edge_count'][b][s])
            outs.append(O.multi_gat(pose[o:o + n], d['edges'][e:e + ne].t(), p, (2, 2)))
            o += n
            e += ne
    ref = torch.cat(outs)
    gat = MultiGAT(n_units=[3, 128, 128], n_heads=[2, 2])
    gat.load_state_dict({k[len('structure_encoder.'):]: v for k, v in p.items() if k.startswith('structure_encoder.')}, strict=True)
    gat = gat.to(dev)
    graph = ops.BatchGraph(d['edges'].to(dev), d['graph_per_obj_count'], d['graph_per_edge_count'])
    with torch.no_grad():
        out = gat(d['tot_rel_pose'].to(dev), graph)
    torch.cuda.synchronize()
    assert rel_inf(out, ref) < 2e-5


def test_csr_rows(dev):
    """CSR content: self loops dropped, duplicates kept, one self loop appended last, input order."""
    from sgaligner_b200 import ops
    c = load_case('messy_pg')
    d = c['data']
    g = ops.BatchGraph(d['edges'].to(dev), d['graph_per_obj_count'], d['graph_per_edge_count'])
    torch.cuda.synchronize()
    rb, rc, col = g.row_beg.cpu().numpy(), g.row_cnt.cpu().numpy(), g.col.cpu().numpy()
    oc = np.asarray(d['graph_per_obj_count']).reshape(-1)
    ec = np.asarray(d['graph_per_edge_count']).reshape(-1)
    ed = d['edges'].numpy()
    n0 = e0 = 0
    for n, e in zip(oc, ec):
        ge = ed[e0:e0 + e]
        for i in range(n):
            want = [n0 + s for s, t in ge if t == i and s != i] + [n0 + i]
            got = col[rb[n0 + i]:rb[n0 + i] + rc[n0 + i]].tolist()
            assert got == want
        n0 += n
        e0 += e


@pytest.mark.parametrize('name', CASES)
def test_encoder_vs_golden(name, dev):
    c = load_case(name)
    model = _model(c, dev).eval()
    with torch.no_grad():
        out = model(_cuda_data(c, dev))
    torch.cuda.synchronize()
    assert set(out.keys()) == set(c['out'].keys())
    for k, ref in c['out'].items():
        assert out[k].shape == ref.shape
        err = rel_inf(out[k], ref)
        assert err < EMB_TOL, (k, err)
        cos = torch.nn.functional.cosine_similarity(out[k].double().cpu(), ref.double(), dim=1)
        assert float(cos.min()) > 1 - 1e-6, (k, float(cos.min()))


@pytest.mark.parametrize('name', CASES)
def test_matching_vs_golden(name, dev):
    """rank_list / Hits@k / MRR against the reference's own matching head (inference_align_reg.py:125-128
    + utils/alignment.py) evaluated on the GOLDEN embedding, so that only the matching kernels are
    under test.  Integer results must be bit-exact wherever the oracle's adjacent-rank gap exceeds
    fp32 noise."""
    from sgaligner_b200 import matching
    c = load_case(name)
    key = 'joint' if len(c['modules']) > 1 else c['modules'][0]
    emb = c['out'][key].to(dev)
    res = matching.match_batch(emb, c['data'], k=6, full_rank=True)
    ev = matching.evaluate_batch(emb, c['data'])
    torch.cuda.synchronize()
    assert [ev['hits'][k] for k in range(1, 6)] == c['hits'].tolist()
    np.testing.assert_allclose(np.sort(ev['rr']), np.sort(c['rr']), rtol=0, atol=0)
    ranks = matching.rank_lists(res)
    lay = res['layout']
    for b, r in enumerate(ranks):
        r = r.cpu().numpy()
        ref_rank, ref_sim = c['rank'][b], c['sim'][b]
        n = r.shape[0]
        assert sorted(r[0].tolist()) == list(range(n))
        srt = np.take_along_axis(ref_sim, ref_rank, 1)
        gap_ok = np.ones_like(ref_rank, dtype=bool)
        gaps = np.diff(srt, axis=1) > 1e-5
        gap_ok[:, 1:] &= gaps
        gap_ok[:, :-1] &= gaps
        assert (r[gap_ok] == ref_rank[gap_ok]).all()
        o0 = int(lay.pair_off_host[b])
        tk = res['topk_idx'][o0:o0 + n].cpu().numpy()
        kk = min(6, n)
        assert (tk[:, :kk] == r[:, :kk]).all()
        sim_dev = res['sim'][int(lay.sim_off_host[b]):int(lay.sim_off_host[b]) + n * n].view(n, n).cpu().numpy()
        assert np.abs(sim_dev - ref_sim).max() < 1e-5   # tensor-core fp32 accumulation truncates: ~-2^-24 per accumulate step, one-signed


@pytest.mark.parametrize('name', CASES)
@pytest.mark.parametrize('tc', [True, False])
def test_pair_metrics_vs_reference(name, tc, dev):
    """Device-side SGAR / alignment score / node correspondences / Hits@k / MRR (one D2H copy for the whole batch)
    against the values of the reference's ``utils/alignment.py`` on the golden embedding."""
    import os
    from sgaligner_b200 import matching
    from tests.util import GOLD
    ref = np.load(os.path.join(GOLD, 'metrics_ref.npz'))
    c = load_case(name)
    key = 'joint' if len(c['modules']) > 1 else c['modules'][0]
    ev = matching.evaluate_pairs(c['out'][key].to(dev), c['data'], tensor_cores=tc)
    assert [ev['hits'][k] for k in range(1, 6)] == c['hits'].tolist()
    np.testing.assert_allclose(np.sort(ev['rr']), np.sort(c['rr']), rtol=0, atol=0)
    si = 0
    for b in range(c['data']['batch_size']):
        assert abs(ev['alignment_score'][b] - float(ref[f'{name}/{b}/alignment_score'])) < 1e-6
        assert ev['node_corrs'][b] == [tuple(int(v) for v in r) for r in ref[f'{name}/{b}/node_corrs']]
        if int(c['data']['e1i_count'][b]):
            assert [ev['sgar'][m][si] for m in ('2', '50', '100')] == ref[f'{name}/{b}/sgar'].tolist()
            si += 1


def test_pair_metrics_random_vs_oracle(dev):
    """Untrained-looking embeddings (many wrong anchors, so every SGAR mode takes both values) at C2 shapes,
    including a pair without anchors and a pair with one anchor."""
    from oracle import sgaligner_oracle as O
    from sgaligner_b200 import matching, synthetic
    data = synthetic.make_batch([20, 33, 8, 64, 5], [25, 30, 9, 64, 7], [6, 9, 0, 32, 1], n_points=8, seed=11)
    g = torch.Generator().manual_seed(3)
    N = int(data['tot_obj_pts'].shape[0])
    emb = torch.randn(N, 48, generator=g)
    # make about half of the anchors easy: copy the source row onto its reference row (+ noise)
    e1, e2 = np.asarray(data['e1i']), np.asarray(data['e2i'])
    for t in range(0, len(e1), 2):
        emb[int(e2[t])] = emb[int(e1[t])] + 0.05 * torch.randn(48, generator=g)
    ev = matching.evaluate_pairs(emb.to(dev), data)
    ref = O.evaluate_batch(emb, data)
    assert ev['hits'] == ref['hits'] and abs(ev['mrr'] - ref['mrr']) < 1e-12
    assert ev['sgar'] == ref['sgar']
    assert ev['node_corrs'] == ref['node_corrs']
    np.testing.assert_allclose(ev['alignment_score'], ref['alignment_score'], atol=1e-6)
    vals = set(v for m in ev['sgar'].values() for v in m)
    assert vals == {0.0, 1.0}


@pytest.mark.parametrize('name', CASES)
def test_loss_forward_vs_golden(name, dev):
    """OverallLoss on the GOLDEN embeddings (isolates the loss kernels)."""
    from sgaligner_b200.losses import CustomMultiLossLayer, OverallLoss
    c = load_case(name)
    M = len(c['modules'])
    li, lc = CustomMultiLossLayer(M), CustomMultiLossLayer(M)
    with torch.no_grad():
        li.log_vars.copy_(c['lv'][0])
        lc.log_vars.copy_(c['lv'][1])
    fn = OverallLoss(li.to(dev), lc.to(dev), dev, {'zoom': 0.1, 'wt_align_loss': 1.0, 'wt_contrastive_loss': 1.0, 'modules': c['modules']})
    out = {k: v.to(dev) for k, v in c['out'].items()}
    with torch.no_grad():
        ld = fn(out, dict(c['data']))
    torch.cuda.synchronize()
    for k, ref in c['loss'].items():
        got = float(ld[k])
        assert abs(got - ref) <= LOSS_TOL * max(abs(ref), 1e-6), (k, got, ref)


def test_no_cpu_fallback(dev):
    """The product path must refuse CPU tensors instead of silently computing elsewhere."""
    c = load_case('small4')
    model = _model(c, dev)
    with pytest.raises(RuntimeError):
        model(dict(c['data']))


def test_streamed_h2d_matches_plain(dev):
    """data.to_cuda_streamed (chunked, overlapped point copy) gives bit-identical embeddings."""
    from sgaligner_b200 import to_cuda
    from sgaligner_b200.data import pin, to_cuda_streamed
    c = load_case('mid4')
    model = _model(c, dev).eval()
    host = pin(dict(c['data']))
    with torch.no_grad():
        a = model(to_cuda(dict(host), dev))
        b = model(to_cuda_streamed(host, dev, n_chunks=3))
    torch.cuda.synchronize()
    for k in a:
        assert torch.equal(a[k], b[k]), k
