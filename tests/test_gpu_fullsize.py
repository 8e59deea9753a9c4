"""GPU: BASELINE.json's full-size configurations through size-independent properties (the CPU oracle
cannot finish them in seconds): kernel-vs-kernel parity at full size, shard invariance, permutation /
self-match properties of the ranking, gradient consistency by finite differences, plus oracle parity
on a sub-sample of the same generator."""
import numpy as np
import pytest
import torch

from oracle import sgaligner_oracle as O
from tests.util import rel_inf

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available()
    return torch.device('cuda:0')


def _model(modules, dev, seed=0, **kw):
    from sgaligner_b200.sg_aligner import MultiModalEncoder
    torch.manual_seed(seed)
    return MultiModalEncoder(modules=modules, rel_dim=41, attr_dim=164, **kw).to(dev)


def test_c2_full_size_tc_vs_fma_pointnet(dev):
    """configs[1] at full size (32 pairs x 128 objects x 512 points): tensor-core PointNet vs the fp32
    FMA kernel on all 4096 objects, and the argmax they report points at a maximum."""
    from sgaligner_b200 import ops, synthetic
    data = synthetic.config_c2(batch=32, seed=0)
    m = _model(['point', 'gat'], dev)
    enc = m.object_encoder
    w = [enc.conv1.weight, enc.conv1.bias, enc.conv2.weight, enc.conv2.bias, enc.conv3.weight, enc.conv3.bias]
    pts = data['tot_obj_pts'].to(dev)
    with torch.no_grad():
        a, arg = ops.pointnet_forward(pts, *w, want_argmax=True, mode=ops.POINTNET_TC)
        b, _ = ops.pointnet_forward(pts, *w, want_argmax=False, mode=ops.POINTNET_SIMT)
    torch.cuda.synchronize()
    assert a.shape == (4096, 256)
    assert rel_inf(a, b) < 3e-5
    assert int(arg.min()) >= 0 and int(arg.max()) < 512


def test_c2_shard_invariance_and_ranking_properties(dev):
    """Pairs are independent through encoder + matching head: any contiguous shard of the batch gives
    bit-identical embeddings and rankings (this is what makes the multi-GPU split exact); every rank
    row is a permutation whose first entry is the node itself."""
    from sgaligner_b200 import matching, synthetic, to_cuda
    data = synthetic.config_c2(batch=8, seed=3)
    m = _model(['point', 'gat'], dev).eval()
    with torch.no_grad():
        full = m(to_cuda(dict(data), dev))
        res = matching.match_batch(full['joint'], data, k=6, full_rank=True)
        parts = [m(to_cuda(synthetic.shard_batch(data, r, 4), dev)) for r in range(4)]
    torch.cuda.synchronize()
    for k in full:
        assert torch.equal(full[k], torch.cat([p[k] for p in parts])), k
    for r in matching.rank_lists(res):
        r = r.cpu().numpy()
        n = r.shape[0]
        assert (np.sort(r, axis=1) == np.arange(n)[None]).all()
        assert (r[:, 0] == np.arange(n)).all()          # sim(i,i) = 0 is the row minimum
    tk = res['topk_idx'].cpu().numpy()
    lay = res['layout']
    for b, r in enumerate(matching.rank_lists(res)):
        o = int(lay.pair_off_host[b])
        assert (tk[o:o + r.shape[0], 0] == np.arange(r.shape[0])).all()


def test_c3_shaped_batch_vs_oracle_subsample(dev):
    """configs[2] generator (3RScan-shaped, complete digraphs, 4 modules, train-style anchors): 6 pairs of
    it against the CPU oracle -- embeddings 1e-4, losses 1e-3, Hits@k identical."""
    from sgaligner_b200 import matching, synthetic, to_cuda
    from sgaligner_b200.losses import CustomMultiLossLayer, OverallLoss
    mods = ['point', 'gat', 'rel', 'attr']
    data = synthetic.slice_pairs(synthetic.config_c3(batch=16, seed=1, n_points=256), 0, 6)
    m = _model(mods, dev).eval()
    params = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    li, lc = CustomMultiLossLayer(4).to(dev), CustomMultiLossLayer(4).to(dev)
    fn = OverallLoss(li, lc, dev, {'zoom': 0.1, 'wt_align_loss': 1.0, 'wt_contrastive_loss': 1.0, 'modules': mods})
    with torch.no_grad():
        out = m(to_cuda(dict(data), dev))
        ld = fn(out, dict(data))
        ev = matching.evaluate_batch(out['joint'], data)
        o_out = O.encoder_forward(params, data, mods)
        o_ld = O.overall_loss(o_out, data, mods, torch.zeros(4), torch.zeros(4))
        o_ev = O.evaluate_batch(o_out['joint'], data)
    torch.cuda.synchronize()
    for k in o_out:
        assert rel_inf(out[k], o_out[k]) < 1e-4, k
    for k in ('loss', 'icl_loss_unimodal', 'icl_loss_multimodal', 'ial_loss'):
        assert abs(float(ld[k]) - float(o_ld[k])) <= 1e-3 * abs(float(o_ld[k])), k
    assert ev['hits'] == o_ev['hits']


def test_c3_full_batch_runs_and_gradient_is_consistent(dev):
    """configs[2] at full batch (B = 128, N ~ 7k nodes, A ~ 800 anchors): one training step's analytic
    gradient agrees with a central finite difference of the loss along a random direction."""
    from sgaligner_b200 import synthetic, to_cuda
    from sgaligner_b200.losses import CustomMultiLossLayer, OverallLoss
    mods = ['point', 'gat', 'rel', 'attr']
    data = to_cuda(synthetic.config_c3(batch=128, seed=1, n_points=128), dev)
    m = _model(mods, dev).eval()
    li, lc = CustomMultiLossLayer(4).to(dev), CustomMultiLossLayer(4).to(dev)
    fn = OverallLoss(li, lc, dev, {'zoom': 0.1, 'wt_align_loss': 1.0, 'wt_contrastive_loss': 1.0, 'modules': mods})
    ps = [m.object_embedding.weight, m.structure_embedding.weight, m.meta_embedding_rel.weight, m.fusion.weight, li.log_vars, lc.log_vars]
    ld = fn(m(data), data)
    loss0 = float(ld['loss'].detach())
    assert np.isfinite(loss0)
    ld['loss'].backward()
    g = [p.grad.detach().clone() for p in ps]
    torch.manual_seed(5)
    dirs = [torch.randn_like(p) for p in ps]
    analytic = sum(float((gi.double() * di.double()).sum()) for gi, di in zip(g, dirs))
    eps = 2e-3
    vals = []
    with torch.no_grad():
        for sgn in (+1, -1):
            for p, d_ in zip(ps, dirs):
                p.add_(sgn * eps * d_)
            vals.append(float(fn(m(data), data)['loss']))
            for p, d_ in zip(ps, dirs):
                p.sub_(sgn * eps * d_)
    numeric = (vals[0] - vals[1]) / (2 * eps)
    assert abs(numeric - analytic) <= 2e-2 * max(abs(analytic), abs(numeric)), (numeric, analytic)


def test_c5_shapes(dev):
    """configs[4] shapes: 256 objects/scene, 1024 points/object, pt_out_dim 512, emb_dim 128 (joint 512-d
    with 4 modules): several M-tiles / channel blocks everywhere; tensor-core kernels vs the FMA kernels."""
    from sgaligner_b200 import matching, ops, synthetic, to_cuda
    mods = ['point', 'gat', 'rel', 'attr']
    data = synthetic.config_c5(batch=2, seed=2)
    m = _model(mods, dev, pt_out_dim=512, emb_dim=128).eval()
    d = to_cuda(dict(data), dev)
    with torch.no_grad():
        out = m(d)
        m.object_encoder.kernel_mode = ops.POINTNET_SIMT
        out2 = m(d)
        a = matching.match_batch(out['joint'], data, k=8, tensor_cores=True)
        b = matching.match_batch(out['joint'], data, k=8, tensor_cores=False)
    torch.cuda.synchronize()
    assert out['joint'].shape == (1024, 512)
    for k in out:
        assert rel_inf(out[k], out2[k]) < 1e-4, k
    assert float((a['sim'] - b['sim']).abs().max()) < 1e-5
    assert float((a['topk_idx'] == b['topk_idx']).float().mean()) > 0.999
    # oracle on one of the two pairs (512 objects x 1024 points is still seconds on the CPU)
    one = synthetic.slice_pairs(data, 0, 1)
    params = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    with torch.no_grad():
        o = O.encoder_forward(params, one, mods)
    n1 = o['joint'].shape[0]
    assert rel_inf(out['joint'][:n1], o['joint']) < 1e-4


@pytest.mark.parametrize('name', ['full_c2', 'full_c3'])
def test_full_size_training_step_vs_reference_golden(name, dev):
    """BASELINE.json configs[1] (C2: 32 pairs, 4096 objects x 512 points, PointNet+GAT) and configs[2] (C3: 128
    3RScan-shaped pairs, 7285 objects, P+S+R+A) at FULL size against the UNMODIFIED reference run in the build
    container (oracle/make_golden_fullsize.py -> tests/golden/full_c{2,3}.npz): embeddings 1e-4 (every 8th row +
    whole-tensor checksums), the four losses 1e-3, EVERY parameter gradient 1e-3 in the default tensor-core mode,
    BatchNorm running statistics 1e-4, Hits@1..5 identical."""
    import os
    from oracle.make_golden_fullsize import CONFIGS, input_checksum
    from sgaligner_b200 import matching, ops, to_cuda
    from sgaligner_b200.losses import CustomMultiLossLayer, OverallLoss
    from sgaligner_b200.sg_aligner import MultiModalEncoder
    from tests.util import GOLD, grad_close
    z = np.load(os.path.join(GOLD, name + '.npz'))
    gen, modules, _ = CONFIGS[name]
    data = gen()
    assert np.allclose(input_checksum(data), z['in/checksum'], rtol=1e-12), 'synthetic generator drifted from the golden inputs'
    M = len(modules)
    model = MultiModalEncoder(modules=modules, rel_dim=41, attr_dim=164)
    model.load_state_dict({k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith('p/')}, strict=True)
    model = model.to(dev).train()
    li, lc = CustomMultiLossLayer(M).to(dev), CustomMultiLossLayer(M).to(dev)
    fn = OverallLoss(li, lc, dev, {'zoom': 0.1, 'wt_align_loss': 1.0, 'wt_contrastive_loss': 1.0, 'modules': modules})
    d = to_cuda(dict(data), dev)
    ops.pointnet_tie_stats(reset=True)
    out = model(d)
    ld = fn(out, d)
    ld['loss'].backward()
    torch.cuda.synchronize()
    n_tie, n_flip = ops.pointnet_tie_stats()
    print(f'[{name}] max-pool near-ties re-evaluated in fp32: {n_tie} of {out[modules[0]].shape[0] * 256}, reordered {n_flip}')
    stride = int(z['cfg/stride'])
    for k in out:
        ref = torch.from_numpy(z['out/' + k])
        assert rel_inf(out[k][::stride], ref) < 1e-4, k
        s = out[k].detach().double()
        got = np.array([float(s.sum()), float(s.pow(2).sum())])
        scale = np.array([float(s.abs().sum()), float(s.pow(2).sum())])
        assert (np.abs(got - z['sum/' + k]) <= 1e-5 * scale).all(), (k, got, z['sum/' + k])
    for k in ('loss', 'icl_loss_unimodal', 'icl_loss_multimodal', 'ial_loss'):
        assert abs(float(ld[k]) - float(z['loss/' + k])) <= 1e-3 * abs(float(z['loss/' + k])), k
    named = dict(model.named_parameters())
    gkeys = [k[5:] for k in z.files if k.startswith('grad/')]
    atol = 1e-6 * max(float(np.abs(z['grad/' + k]).max()) for k in gkeys)
    worst = ('', 0.0)
    for k in gkeys:
        ref = torch.from_numpy(z['grad/' + k])
        got = li.log_vars.grad if k == '__lv_ial' else lc.log_vars.grad if k == '__lv_icl' else named[k].grad
        if got is None:
            assert float(ref.abs().max()) == 0.0, k
            continue
        e = rel_inf(got, ref)
        worst = max(worst, (k, e), key=lambda t: t[1])
        assert grad_close(got, ref, rtol=1e-3, atol=atol), (k, e)
    print(f'[{name}] worst parameter-gradient error {worst[1]:.2e} ({worst[0]})')
    sd = model.state_dict()
    for k in z.files:
        if k.startswith('bn/'):
            if 'num_batches' in k:
                assert int(sd[k[3:]]) == int(z[k])
            else:
                assert rel_inf(sd[k[3:]], torch.from_numpy(z[k])) < 1e-4, k
    ev = matching.evaluate_batch(out['joint'].detach(), data)
    assert [ev['hits'][i] for i in range(1, 6)] == [int(v) for v in z['metric/hits']]
