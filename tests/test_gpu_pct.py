"""GPU parity of the NaivePCT object encoder (src/aligner/networks/pct.py:275-317):
every tensor-core stage against an fp64 torch evaluation of the same formula, the whole encoder against the outputs of the
UNMODIFIED reference module (tests/golden/pct_ref.npz: eval mode; train mode with the BatchNorm running-statistics side
effect and the dropout masks of the recorded seed), and MultiModalEncoder(['pct','gat','rel','attr']) -- the module list
of the shipped config -- against the reference encoder (tests/golden/pct_encoder.npz)."""
import json
import math
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import pct_oracle
from tests.util import GOLD, rel_inf

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available()
    return torch.device('cuda:0')


def _rand(shape, dev, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(dev)


@pytest.mark.parametrize('N,P', [(3, 96), (5, 128), (7, 300), (150, 512)])
def test_pointwise_conv_prologues_and_stats(N, P, dev):
    """Y = (g1(src1) + g2(src2)) W^T + b for the prologue forms the encoder uses, the k|v split output, the stored X
    and the BatchNorm sums, vs fp64."""
    from sgaligner_b200 import ops
    s1, s2 = _rand((N, P, 128), dev, 1), _rand((N, P, 128), dev, 2)
    a1, b1, a2, b2 = [_rand((128,), dev, 3 + i, 0.7) for i in range(4)]
    W = _rand((160, 128), dev, 9, 1 / math.sqrt(128))
    bias = _rand((160,), dev, 10, 0.1)
    X = torch.relu(a1.double() * s1.double() + b1.double()) + torch.relu(a2.double() * s2.double() + b2.double())
    Y = X @ W.double().t() + bias.double()
    k, v, x, st = ops.pct_pointwise(s1, (a1, b1), s2, (a2, b2), W, bias, 32, want_x=True, want_stats=True)
    torch.cuda.synchronize()
    assert rel_inf(x, X) < 1e-6
    assert rel_inf(k, Y[..., :32]) < 2e-5 and rel_inf(v, Y[..., 32:]) < 2e-5
    flat = Y.reshape(-1, 160)
    assert rel_inf(st[:160], flat.sum(0)) < 1e-5 and rel_inf(st[160:], (flat * flat).sum(0)) < 1e-5
    # the same k | v product without statistics (what the encoder launches: the two-CTA kernel with the points on the lanes
    # for k), two sources and one source
    k2, v2, x2, _ = ops.pct_pointwise(s1, (a1, b1), s2, (a2, b2), W, bias, 32, want_x=True, want_stats=False)
    torch.cuda.synchronize()
    assert rel_inf(x2, X) < 1e-6
    assert rel_inf(k2, Y[..., :32]) < 2e-5 and rel_inf(v2, Y[..., 32:]) < 2e-5
    k3, v3, _, _ = ops.pct_pointwise(s1, None, None, None, W, None, 32, want_x=False, want_stats=False)
    Y3 = s1.double() @ W.double().t()
    torch.cuda.synchronize()
    assert rel_inf(k3, Y3[..., :32]) < 2e-5 and rel_inf(v3, Y3[..., 32:]) < 2e-5
    # ... and with the per-object maxima the backward's operand scales are derived from (recorded in the epilogues)
    k4, v4, x4, vmax = ops.pct_pointwise_kv(s1, (a1, b1), s2, (a2, b2), W, bias, want_x=True)
    gsrc = _rand((N, P, 128), dev, 77, 1e-3)
    dx, dxmax = ops.pct_pointwise_grad(gsrc, W[32:].t().contiguous(), want_absmax=True)
    torch.cuda.synchronize()
    assert torch.equal(k4, k2) and torch.equal(v4, v2) and torch.equal(x4, x2)
    assert torch.equal(vmax, v4.abs().amax(dim=(1, 2))) and torch.equal(dxmax, dx.abs().amax(dim=(1, 2)))
    assert rel_inf(dx, gsrc.double() @ W[32:].double()) < 2e-5
    # identity prologue, one source, Cout = 128
    W2 = W[:128].contiguous()
    y, _, _, st2 = ops.pct_pointwise(s1, None, None, None, W2, bias[:128].contiguous(), 128, want_x=False, want_stats=True)
    Y2 = s1.double() @ W2.double().t() + bias[:128].double()
    torch.cuda.synchronize()
    assert rel_inf(y, Y2) < 2e-5
    assert rel_inf(st2[:128], Y2.reshape(-1, 128).sum(0)) < 1e-5
    # identity + relu-affine (the x_l = x_{l-1} + relu(bn(t_l)) residual)
    y3, _, x3, _ = ops.pct_pointwise(s1, None, s2, (a2, b2), W2, None, 128, want_x=True, want_stats=False)
    X3 = s1.double() + torch.relu(a2.double() * s2.double() + b2.double())
    torch.cuda.synchronize()
    assert rel_inf(x3, X3) < 1e-6 and rel_inf(y3, X3 @ W2.double().t()) < 2e-5


@pytest.mark.parametrize('N,P', [(4, 96), (6, 256), (150, 512)])
def test_embed_stage(N, P, dev):
    from sgaligner_b200 import ops
    pts = _rand((N, P, 3), dev, 1)
    W1, W2 = _rand((128, 3), dev, 2, 0.6), _rand((128, 128), dev, 3, 1 / math.sqrt(128))
    a1, b1 = _rand((128,), dev, 4, 0.7), _rand((128,), dev, 5, 0.3)
    z2, st = ops.pct_embed(pts, W1, a1, b1, W2, True)
    h = torch.relu(a1.double() * (pts.double() @ W1.double().t()) + b1.double())
    Z = h @ W2.double().t()
    torch.cuda.synchronize()
    assert rel_inf(z2, Z) < 2e-5
    assert rel_inf(st[:128], Z.reshape(-1, 128).sum(0)) < 1e-5 and rel_inf(st[128:], (Z * Z).reshape(-1, 128).sum(0)) < 1e-5
    # the statistics of conv1's output from the point moments
    mom = ops.pct_point_moments(pts)
    s1 = ops.pct_affine_stats(mom, W1)
    z1 = (pts.double() @ W1.double().t()).reshape(-1, 128)
    torch.cuda.synchronize()
    assert rel_inf(s1[:128], z1.sum(0)) < 1e-9 + 1e-6 and rel_inf(s1[128:], (z1 * z1).sum(0)) < 1e-6


@pytest.mark.parametrize('N,P', [(3, 96), (4, 128), (5, 300), (9, 512), (160, 512), (6, 40), (5, 200), (450, 512)])
def test_attention_vs_fp64(N, P, dev):
    """x_s = bmm(x_v, softmax(x_k^T x_k / sqrt(32), -1)) (pct.py:217-224) vs fp64, ragged point counts included."""
    from sgaligner_b200 import ops
    k = _rand((N, P, 32), dev, 1, 1.3)
    v = _rand((N, P, 128), dev, 2)
    xs = ops.pct_attention(k, v)
    kd, vd = k.double(), v.double()
    att = torch.softmax(kd @ kd.transpose(1, 2) / math.sqrt(32), dim=-1)          # [N, i, j]
    ref = att.transpose(1, 2) @ vd                                                # xs[j, c] = sum_i att[i, j] v[i, c]
    torch.cuda.synchronize()
    assert torch.isfinite(xs).all()
    assert rel_inf(xs, ref) < 3e-5


@pytest.mark.parametrize('N,P,scale', [(4, 128, 12.0), (6, 512, 40.0)])
def test_attention_large_energies(N, P, scale, dev):
    """Energies of 10^3..10^5 (what an untrained / badly scaled network produces, and what the golden input reaches in
    sa3/sa4): the softmax normaliser must not be rounded at the magnitude of the row maximum.  Judged against fp64
    next to torch's own fp32 evaluation of the reference formula on the same inputs."""
    from sgaligner_b200 import ops
    k = _rand((N, P, 32), dev, 1, scale)
    v = _rand((N, P, 128), dev, 2, 5.0)
    xs = ops.pct_attention(k, v)
    kd, vd = k.double(), v.double()
    ref = torch.softmax(kd @ kd.transpose(1, 2) / math.sqrt(32), dim=-1).transpose(1, 2) @ vd
    f32 = torch.softmax(torch.bmm(k, k.transpose(1, 2)) / math.sqrt(32), dim=-1).transpose(1, 2) @ v
    torch.cuda.synchronize()
    e_ours, e_f32 = rel_inf(xs, ref), rel_inf(f32, ref)
    print('attention at |energy| ~ %.0f: ours %.2e, torch fp32 %.2e vs fp64' % (float((kd * kd).sum(-1).max() / math.sqrt(32)), e_ours, e_f32))
    assert torch.isfinite(xs).all()
    assert e_ours < max(3e-5, 4 * e_f32)


@pytest.mark.parametrize('N,P', [(3, 40), (4, 96), (5, 300), (40, 512)])
def test_cat_linear_pooling(N, P, dev):
    from sgaligner_b200 import ops
    x1, x2, x3, t4 = [_rand((N, P, 128), dev, 1 + i) for i in range(4)]
    a4, b4 = _rand((128,), dev, 7, 0.7), _rand((128,), dev, 8, 0.3)
    WL = _rand((1024, 512), dev, 9, 1 / math.sqrt(512))
    zmax, zmin, st, _, _ = ops.pct_cat_linear(x1, x2, x3, t4, (a4, b4), WL)
    x4 = x3.double() + torch.relu(a4.double() * t4.double() + b4.double())
    Z = torch.cat([x1.double(), x2.double(), x3.double(), x4], dim=2) @ WL.double().t()      # [N, P, 1024]
    aL, bL = _rand((1024,), dev, 11, 0.8), _rand((1024,), dev, 12, 0.3)
    pooled = ops.pct_pool_act(zmax, zmin, aL, bL, P)
    ref = F.leaky_relu(aL.double() * Z + bL.double(), 0.2).max(dim=1).values
    torch.cuda.synchronize()
    assert rel_inf(pooled, ref) < 3e-5
    flat = Z.reshape(-1, 1024)
    assert rel_inf(st[:1024], flat.sum(0)) < 1e-5 and rel_inf(st[1024:], (flat * flat).sum(0)) < 1e-5
    # training forward: the same values plus the arg-max point of every (object, channel) -- torch.max's index
    zmax2, zmin2, st2, imax, imin = ops.pct_cat_linear(x1, x2, x3, t4, (a4, b4), WL, track=True)
    pooled2, pstar, zsel = ops.pct_pool_act(zmax2, zmin2, aL, bL, P, imax, imin)
    torch.cuda.synchronize()
    assert torch.equal(pooled2, pooled) and torch.equal(zmax2, zmax) and torch.equal(zmin2, zmin)
    Y = F.leaky_relu(aL.double() * Z + bL.double(), 0.2)
    ref_idx = Y.max(dim=1).indices
    picked = torch.gather(Y, 1, pstar.long()[:, None, :])[:, 0]
    assert rel_inf(picked, ref) < 3e-5                                          # the tracked point attains the maximum
    assert float((pstar.long() == ref_idx).float().mean()) > 0.999              # (near-ties may resolve differently)
    assert rel_inf(zsel, torch.gather(Z, 1, pstar.long()[:, None, :])[:, 0]) < 3e-5


def _gold():
    return np.load(os.path.join(GOLD, 'pct_ref.npz'))


def test_naive_pct_eval_vs_reference(dev):
    from sgaligner_b200.pct import NaivePCT
    z = _gold()
    m = NaivePCT()
    m.load_state_dict(pct_oracle.random_params(int(z['param_seed'])), strict=True)
    m = m.to(dev).eval()
    x = torch.from_numpy(z['x']).permute(0, 2, 1).contiguous().to(dev)        # golden input is [B, 3, N]
    with torch.no_grad():
        y = m(x)
    torch.cuda.synchronize()
    ref = torch.from_numpy(z['y_eval'])
    assert y.shape == ref.shape
    print('NaivePCT eval rel error vs the reference: %.2e' % rel_inf(y, ref))
    assert rel_inf(y, ref) < 1e-4
    assert int(m.bn1.num_batches_tracked) == 0


def test_naive_pct_train_mode_vs_reference(dev):
    """Batch statistics through all nine BatchNorm layers, the running-statistics side effect and the two dropouts
    (masks of the recorded seed, drawn like the reference's F.dropout on CPU)."""
    from sgaligner_b200.pct import NaivePCT
    z = _gold()
    m = NaivePCT()
    m.load_state_dict(pct_oracle.random_params(int(z['param_seed'])), strict=True)
    m = m.to(dev).train()
    m.dropout_rng = 'cpu'
    x = torch.from_numpy(z['x']).permute(0, 2, 1).contiguous().to(dev)
    torch.manual_seed(int(z['train_seed']))
    with torch.no_grad():
        y = m(x)
    torch.cuda.synchronize()
    ref = torch.from_numpy(z['y_train'])
    print('NaivePCT train rel error vs the reference: %.2e' % rel_inf(y, ref))
    assert rel_inf(y, ref) < 1e-4
    assert torch.equal(y.cpu() == 0, ref == 0)                                  # same dropout / ReLU pattern
    sd = m.state_dict()
    for key in z.files:
        if key.startswith('after/'):
            r = torch.from_numpy(z[key])
            if 'num_batches' in key:
                assert int(sd[key[6:]]) == int(r), key
            else:
                assert rel_inf(sd[key[6:]], r) < 1e-4, key


def test_naive_pct_full_size_eval_vs_oracle(dev):
    """512 points per object (4 point tiles, every pipeline stage in steady state): against the fp32 oracle."""
    from sgaligner_b200.pct import NaivePCT
    p = pct_oracle.random_params(7)
    m = NaivePCT()
    m.load_state_dict(p, strict=True)
    m = m.to(dev).eval()
    g = torch.Generator().manual_seed(3)
    x = torch.randn(10, 512, 3, generator=g) + torch.rand(10, 1, 3, generator=g) * 2 - 1
    with torch.no_grad():
        y = m(x.to(dev))
        ref = pct_oracle.naive_pct(x.permute(0, 2, 1).double(), {k: (v.double() if v.is_floating_point() else v) for k, v in p.items()})
    torch.cuda.synchronize()
    print('NaivePCT eval, P = 512: rel error vs the fp64 oracle %.2e' % rel_inf(y, ref))
    assert rel_inf(y, ref) < 1e-4


def test_encoder_with_pct_vs_reference(dev):
    """MultiModalEncoder(['pct','gat','rel','attr']): strict state_dict compatibility with the reference key set, eval
    embeddings 1e-4, Hits identical, train-mode running statistics of the PCT BatchNorm layers."""
    from oracle.make_golden_pct_encoder import MODULES, batch, seeded_product_state_dict
    from sgaligner_b200 import matching, to_cuda
    from sgaligner_b200.sg_aligner import MultiModalEncoder
    z = np.load(os.path.join(GOLD, 'pct_encoder.npz'))
    keys = json.loads(str(z['keys']))
    sd = seeded_product_state_dict(int(z['seed']))
    assert {k: list(v.shape) for k, v in sd.items()} == keys                   # the reference's exact key set and shapes
    model = MultiModalEncoder(modules=MODULES, rel_dim=41, attr_dim=164)
    model.load_state_dict(sd, strict=True)
    model = model.to(dev).eval()
    data = batch()
    d = to_cuda(dict(data), dev)
    with torch.no_grad():
        out = model(d)
    torch.cuda.synchronize()
    for k in ('pct', 'gat', 'rel', 'attr', 'joint'):
        assert rel_inf(out[k], torch.from_numpy(z['out/' + k])) < 1e-4, k
    ev = matching.evaluate_batch(out['joint'], data)
    assert [ev['hits'][i] for i in range(1, 6)] == [int(v) for v in z['metric/hits']]
    model.train()
    model.object_encoder.dropout_rng = 'cpu'
    torch.manual_seed(5)
    with torch.no_grad():
        out_t = model(d)
    torch.cuda.synchronize()
    assert rel_inf(out_t['pct'], torch.from_numpy(z['train/pct'])) < 1e-4
    after = model.state_dict()
    for key in z.files:
        if key.startswith('after/') and 'running' in key:
            assert rel_inf(after[key[6:]], torch.from_numpy(z[key])) < 1e-4, key
