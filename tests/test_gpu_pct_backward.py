"""GPU parity of the NaivePCT BACKWARD (autograd of src/aligner/networks/pct.py:275-317): the building
blocks against fp64 autograd of the same formulas, and every parameter gradient of the encoder against (i) the gradients
of the UNMODIFIED reference module (tests/golden/pct_ref.npz: train mode, recorded dropout seed) and (ii) fp64 autograd of
the oracle at shapes with several point tiles, in train() and eval().  Gate: 1e-3 of the tensor's largest gradient
(tensors whose gradient is mathematically zero -- a bias in front of a train-mode BatchNorm -- against 1e-3 of the largest
gradient of the model instead: both sides hold rounding noise there)."""
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


@pytest.mark.parametrize('rows,C,masked,slope', [(1000, 128, False, 0.0), (77, 512, True, 0.0), (33, 1024, False, 0.2), (5000, 256, True, 0.0)])
@pytest.mark.parametrize('training', [True, False])
def test_bn_backward_vs_autograd(rows, C, masked, slope, training, dev):
    from sgaligner_b200 import ops
    y = _rand((rows, C), dev, 1) * 2 + 0.3
    g = _rand((rows, C), dev, 2, 1e-4)                      # gradient-sized values
    bn = torch.nn.BatchNorm1d(C).to(dev)
    with torch.no_grad():
        bn.weight.copy_(_rand((C,), dev, 3, 0.3) + 1)
        bn.bias.copy_(_rand((C,), dev, 4, 0.2))
        bn.running_mean.copy_(_rand((C,), dev, 5, 0.2))
        bn.running_var.copy_(torch.rand(C, device=dev) + 0.5)
    lin_bias = _rand((C,), dev, 6, 0.1)
    mask = (torch.rand(rows, C, device=dev) < 0.5).float() if masked else None
    # reference: fp64 autograd
    yd = y.double().requires_grad_(True)
    w = bn.weight.detach().double().requires_grad_(True)
    b = bn.bias.detach().double().requires_grad_(True)
    z = F.batch_norm(yd + lin_bias.double(), bn.running_mean.double().clone(), bn.running_var.double().clone(), w, b, training, 0.1, bn.eps)
    act = F.leaky_relu(z, slope) if slope else F.relu(z)
    if masked:
        act = act * mask.double() * 2.0
    (act * g.double()).sum().backward()
    # ours
    stats = ops.col_stats(y) if training else None
    ab = ops.bn_fold(bn, stats, float(rows), training, lin_bias=lin_bias)
    dy, dga, dbe, _ = ops.bn_backward(g, y, ab, bn, stats, float(rows), training, mask=mask, scale=2.0, slope=slope, lin_bias=lin_bias)
    torch.cuda.synchronize()
    assert rel_inf(dy, yd.grad) < 2e-5
    assert rel_inf(dga, w.grad) < 2e-5 and rel_inf(dbe, b.grad) < 2e-5


@pytest.mark.parametrize('N,P', [(5, 128), (7, 333), (3, 7), (64, 512)])
def test_bn_backward_records_per_object_maximum(N, P, dev):
    """The variant that hands the next tensor-core product its operand scale: max |dy| per object, bit-identical dy; P * 128 / 4
    is not always a multiple of the CTA's 256 float4 (a CTA may straddle two objects)."""
    from sgaligner_b200 import ops
    C = 128
    y, g = _rand((N, P, C), dev, 1), _rand((N, P, C), dev, 2)
    bn = torch.nn.BatchNorm1d(C).to(dev)
    with torch.no_grad():
        bn.weight.copy_(_rand((C,), dev, 3, 0.3) + 1)
        bn.bias.copy_(_rand((C,), dev, 4, 0.2))
    stats = ops.col_stats(y.reshape(-1, C))
    ab = ops.bn_fold(bn, stats, float(N * P), True)
    dy0, _, _, _ = ops.bn_backward(g, y, ab, bn, stats, float(N * P), True)
    dy1, _, _, ex = ops.bn_backward(g, y, ab, bn, stats, float(N * P), True, want_absmax=True)
    torch.cuda.synchronize()
    assert torch.equal(dy0, dy1)
    assert torch.equal(ex[3], dy1.abs().amax(dim=(1, 2)))


@pytest.mark.parametrize('N,P', [(3, 96), (5, 128), (4, 300), (20, 512), (7, 40), (6, 200), (400, 512)])
def test_attention_backward_vs_autograd(N, P, dev):
    from sgaligner_b200 import ops
    k = _rand((N, P, 32), dev, 1, 1.2)
    v = _rand((N, P, 128), dev, 2)
    dxs = _rand((N, P, 128), dev, 3, 1e-5)                  # far below the fp16 range on purpose
    xs, c2 = ops.pct_attention(k, v, want_c2=True)
    dk1, dk2, dv, dv_colsum, dv_absmax = ops.pct_attention_backward(k, v, c2, dxs)
    torch.cuda.synchronize()
    kd = k.double().requires_grad_(True)
    vd = v.double().requires_grad_(True)
    A = torch.softmax(kd @ kd.transpose(1, 2) / math.sqrt(32), dim=-1)
    ref = A.transpose(1, 2) @ vd                             # xs[j] = sum_i A[i, j] v[i]
    assert rel_inf(xs, ref) < 2e-5
    (ref * dxs.double()).sum().backward()
    e_dv, e_dk = rel_inf(dv, vd.grad), rel_inf(dk1 + dk2, kd.grad)
    print('attention backward N=%d P=%d: dv %.2e  dk %.2e' % (N, P, e_dv, e_dk))
    assert e_dv < 1e-4 and e_dk < 1e-4


@pytest.mark.parametrize('N,P', [(2, 96), (3, 300), (150, 512)])
def test_cat_stage_backward_pieces(N, P, dev):
    from sgaligner_b200 import ops
    xs4 = [_rand((N, P, 128), dev, 10 + i) for i in range(4)]
    M = _rand((512, 512), dev, 20, 1e-6)
    M = (M + M.t()).contiguous()
    u = _rand((512,), dev, 21, 1e-6)
    xbar = _rand((512,), dev, 24, 0.5)
    gs = ops.pct_cat_dense_backward(*xs4, M, u, xbar)
    xcat = torch.cat(xs4, dim=-1).double()
    ref = -((xcat - xbar.double()) @ M.double().t()) - u.double()
    torch.cuda.synchronize()
    got = torch.cat(gs, dim=-1)
    assert rel_inf(got, ref) < 1e-5, rel_inf(got, ref)
    # sparse half
    WL = _rand((1024, 512), dev, 22, 1 / math.sqrt(512))
    coef = _rand((N, 1024), dev, 23, 1e-5)
    coef[:, ::7] = 0
    pstar = torch.randint(0, P, (N, 1024), device=dev, dtype=torch.int32)
    dWL = torch.zeros_like(WL)
    base = [g.clone() for g in gs]
    ops.pct_cat_sparse_backward(coef, pstar, WL, xs4, gs, dWL)
    torch.cuda.synchronize()
    D = torch.zeros(N, P, 1024, dtype=torch.float64, device=dev)
    D.scatter_(1, pstar.long()[:, None, :], coef.double()[:, None, :])        # D[n, pstar[n,c], c] = coef[n,c]
    ref_g = torch.cat(base, dim=-1).double() + D @ WL.double()
    ref_w = torch.einsum('npc,npk->ck', D, xcat)
    assert rel_inf(torch.cat(gs, dim=-1), ref_g) < 1e-5
    assert rel_inf(dWL, ref_w) < 1e-5
    # Gram matrix of the concatenated activations through the grouped weight-gradient GEMM
    G = torch.zeros((512, 512), device=dev)
    ops.wgrad_group([(xs4[a].reshape(-1, 128), xs4[b].reshape(-1, 128), G[128 * a:128 * a + 128, 128 * b:128 * b + 128])
                     for a in range(4) for b in range(a, 4)])
    dWk = torch.zeros((32, 128), device=dev)
    small = _rand((N, P, 32), dev, 30, 1e-6)
    ops.wgrad_group([(small.reshape(-1, 32), xs4[0].reshape(-1, 128), dWk)])
    # ... and through the dedicated kernel (accumulators resident in tensor memory, bf16 pairs), plus the transposed
    # [128 | 32]-wide form of d W_v, d W_k
    G2 = torch.zeros((512, 512), device=dev)
    for a in range(4):
        ops.pct_wgrad(xs4[a].reshape(-1, 128), [xs4[b].reshape(-1, 128) for b in range(a, 4)],
                      [G2[128 * a:128 * a + 128, 128 * b:128 * b + 128] for b in range(a, 4)])
    dWv2, dWk2 = torch.zeros((128, 128), device=dev), torch.zeros((32, 128), device=dev)
    gv = _rand((N, P, 128), dev, 31, 1e-6)
    ops.pct_wgrad(xs4[0].reshape(-1, 128), [gv.reshape(-1, 128), small.reshape(-1, 32)], [dWv2, dWk2], transpose=True)
    torch.cuda.synchronize()
    Gr = (xcat.reshape(-1, 512).t() @ xcat.reshape(-1, 512))
    for a in range(4):
        for b in range(a, 4):
            blk = (slice(128 * a, 128 * a + 128), slice(128 * b, 128 * b + 128))
            assert rel_inf(G[blk], Gr[blk]) < 5e-5
            assert rel_inf(G2[blk], Gr[blk]) < 5e-5, (a, b, rel_inf(G2[blk], Gr[blk]))
    assert rel_inf(dWk2, small.double().reshape(-1, 32).t() @ xs4[0].double().reshape(-1, 128)) < 5e-5
    assert rel_inf(dWv2, gv.double().reshape(-1, 128).t() @ xs4[0].double().reshape(-1, 128)) < 5e-5
    assert rel_inf(dWk, small.double().reshape(-1, 32).t() @ xs4[0].double().reshape(-1, 128)) < 5e-5


def _grad_report(named, ref_of, gmax, tol=1e-3):
    worst, lines = 0.0, []
    for name, p in named:
        ref = ref_of(name)
        if ref is None:
            continue
        got = p.grad
        assert got is not None, name
        got = got.detach().double().cpu().reshape(-1)
        if ref.numel() != got.numel():
            got = got[::17]
        ref = ref.detach().double().cpu().reshape(-1)
        err = float((got - ref).abs().max() / max(float(ref.abs().max()), 1e-3 * gmax))
        lines.append('%-28s %.2e' % (name, err))
        worst = max(worst, err)
    return worst, lines


def test_param_grads_vs_reference_golden(dev):
    """Train mode, the reference's recorded dropout seed: gradients of the unmodified reference module."""
    from sgaligner_b200.pct import NaivePCT
    z = np.load(os.path.join(GOLD, 'pct_ref.npz'))
    m = NaivePCT()
    m.load_state_dict(pct_oracle.random_params(int(z['param_seed'])), strict=True)
    m = m.to(dev).train()
    m.dropout_rng = 'cpu'
    x = torch.from_numpy(z['x']).permute(0, 2, 1).contiguous().to(dev)
    R = torch.from_numpy(z['grad_R']).to(dev)
    torch.manual_seed(int(z['train_seed']))
    y = m(x)
    assert rel_inf(y, torch.from_numpy(z['y_train'])) < 1e-4
    (y * R).sum().backward()
    torch.cuda.synchronize()

    def ref_of(name):
        key = 'grad/' + name.replace('q_conv', 'k_conv')
        return torch.from_numpy(z[key]) if key in z.files else None

    worst, lines = _grad_report(list(m.named_parameters()), ref_of, float(z['grad_max']))
    print('\n'.join(lines))
    print('NaivePCT parameter gradients vs the reference module: worst %.2e' % worst)
    assert worst < 1e-3


def _oracle_case(N, P, training, seed, dev):
    from sgaligner_b200.pct import NaivePCT
    p = pct_oracle.random_params(13)
    m = NaivePCT()
    m.load_state_dict(p, strict=True)
    m = m.to(dev).train(training)
    m.dropout_rng = 'cpu'
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(N, P, 3, generator=g) * 0.7 + torch.rand(N, 1, 3, generator=g) * 2 - 1
    R = torch.randn(N, 256, generator=g)
    if not training:
        # eval() with CALIBRATED running statistics (one train-mode pass at momentum 1 copies the batch statistics): the
        # seeded recipe's random running_var would leave the activations un-normalised -- attention energies of 3e5,
        # a regime no trained network is in and where fp32 autograd itself is only good to a few per cent
        for mod in m.modules():
            if isinstance(mod, torch.nn.BatchNorm1d):
                mod.momentum = 1.0
        m.train()
        with torch.no_grad():
            m(x.to(dev))
        m.eval()
        p = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    torch.manual_seed(77)
    y = m(x.to(dev))
    (y * R.to(dev)).sum().backward()
    torch.cuda.synchronize()

    def oracle_grads(dt):
        po = {k: (v.clone().to(dt).requires_grad_('running' not in k) if v.is_floating_point() else v.clone()) for k, v in p.items()}
        for sa in ('sa1', 'sa2', 'sa3', 'sa4'):
            po[sa + '.q_conv.weight'] = po[sa + '.k_conv.weight']
        torch.manual_seed(77)
        yo = pct_oracle.naive_pct(x.permute(0, 2, 1).to(dt), po, training=training)
        (yo * R.to(dt)).sum().backward()
        return yo.detach(), po

    yo, po = oracle_grads(torch.float64)
    assert rel_inf(y, yo) < 2e-4
    gmax = max(float(v.grad.abs().max()) for k, v in po.items() if v.is_floating_point() and v.grad is not None)
    worst, lines = _grad_report(list(m.named_parameters()), lambda name: po[name.replace('q_conv', 'k_conv')].grad, gmax)
    return worst, lines


@pytest.mark.parametrize('N,P,training', [(5, 300, True), (4, 200, False), (9, 512, True)])
def test_param_grads_vs_oracle_autograd(N, P, training, dev):
    """(N = 400: several work items per persistent CTA of the two-CTA-per-SM attention kernels, see test_attention_backward.)
    Every parameter gradient against fp64 autograd of the oracle, three input seeds per shape.  The network has ~10^6
    ReLU kinks per seed (five ReLU layers on [N, P, 128]) and a pre-activation within the forward's 1e-5 of zero takes the
    other branch than the reference's: ONE such element moves single gradient entries by a per cent (measured with
    tools/dbg_pct_bwd_chain.py: seed 3 at N=5, P=300 flips one element of sa1's BatchNorm output -> d beta off by 1.4e-3,
    everything upstream of it agrees to 2e-5).  Hence: the median over the seeds meets the 1e-3 gate and at least two of the
    three seeds are within 3e-3 (a flipped max-pool winner can move one weight row by tens of per cent: printed, not gated)."""
    worsts = []
    for seed in (3, 4, 5):
        worst, lines = _oracle_case(N, P, training, seed, dev)
        bad = [ln for ln in lines if float(ln.split()[-1]) > 3e-4]
        print('seed %d: worst %.2e%s' % (seed, worst, ('   [' + '; '.join(' '.join(b.split()) for b in bad) + ']') if bad else ''))
        worsts.append(worst)
    print('NaivePCT parameter gradients vs fp64 oracle autograd (N=%d P=%d train=%s): worst per seed %s'
          % (N, P, training, ['%.2e' % w for w in worsts]))
    assert sorted(worsts)[1] < 1e-3 and sorted(worsts)[1] < 3e-3


@pytest.mark.parametrize('init', ['torch_default', 'stress'])
def test_param_grads_mid_size_vs_eager_cuda(init, dev):
    """512 objects x 512 points (1024 attention work items, 262 k rows in every batch-wide reduction, several work items per
    persistent CTA in every kernel): against the oracle's op sequence evaluated in FP64 on the same GPU (torch eager) -- the
    size the CPU oracle cannot reach in a test -- with the same sequence in fp32 (TF32 off) beside it to show what fp32
    autograd itself loses at this size.
      torch_default: the initialisation training starts from (reference: NaivePCT() as constructed, pct.py:276-297);
                     gate: every tensor within max(5e-3, three times fp32 eager's own worst) of fp64 (measured: ours 1.1e-3 on one
                     BatchNorm bias -- a ReLU decision --, fp32 eager 9.4e-4 on another), median tensor < 2e-4 (measured 3e-5).
      stress:        the seeded recipe of the goldens, whose BatchNorm affine parameters and unnormalised attention output
                     (columns of the softmax do not sum to one, pct.py:224) drive the layer-4 energies to 4.5e4: the forward's
                     7e-6 on k (22-bit operand pairs, compounding over four layers; fp32: 1e-6) is a 2e-4 error on x_s there
                     (`tools/dbg_pct_bwd_chain.py 512 512 1 gpu 21`), and the chain through the peaked softmax multiplies it:
                     the gradients are REPORTED (fp32 eager beside them) with a loose gate -- worst tensor 1e-1, median 1e-2
                     (measured 2.1e-2 / 2.2e-3; fp32 eager 2.9e-3 / 3e-4)."""
    import torch.nn.functional as F_
    from sgaligner_b200.pct import NaivePCT
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    orig_dropout = F_.dropout
    try:
        N, P = 512, 512
        m = NaivePCT()
        if init == 'stress':
            m.load_state_dict(pct_oracle.random_params(21), strict=True)
        else:
            torch.manual_seed(4)
            m = NaivePCT()
        p = {k: v.detach().clone() for k, v in m.state_dict().items()}
        m = m.to(dev).train()
        g = torch.Generator().manual_seed(8)
        x = (torch.randn(N, P, 3, generator=g) * 0.7 + torch.rand(N, 1, 3, generator=g) * 2 - 1).to(dev)
        R = torch.randn(N, 256, generator=g).to(dev)
        masks = [(torch.rand(N, 512, generator=g) < 0.5).float().to(dev), (torch.rand(N, 256, generator=g) < 0.5).float().to(dev)]
        it = iter(masks)
        m._mask = lambda n, c, d: next(it)                     # the same dropout masks on both sides
        y = m(x)
        (y * R).sum().backward()
        torch.cuda.synchronize()

        def eager(dt):
            po = {k: (v.clone().to(dev).to(dt).requires_grad_('running' not in k) if v.is_floating_point() else v.clone().to(dev))
                  for k, v in p.items()}
            for sa in ('sa1', 'sa2', 'sa3', 'sa4'):
                po[sa + '.q_conv.weight'] = po[sa + '.k_conv.weight']
            it2 = iter(masks)
            F_.dropout = lambda t, pr, training: t * next(it2).to(t.dtype) * 2.0
            yo = pct_oracle.naive_pct(x.permute(0, 2, 1).to(dt), po, training=True)
            (yo * R.to(dt)).sum().backward()
            return yo.detach(), po

        yo, po = eager(torch.float64)
        assert rel_inf(y, yo) < 2e-4
        _, p32 = eager(torch.float32)
        gmax = max(float(v.grad.abs().max()) for k, v in po.items() if v.is_floating_point() and v.grad is not None)
        ref_of = lambda name: po[name.replace('q_conv', 'k_conv')].grad      # noqa: E731
        worst, lines = _grad_report(list(m.named_parameters()), ref_of, gmax)

        class _P:
            def __init__(self, g_):
                self.grad = g_
        worst32, lines32 = _grad_report([(n, _P(p32[n.replace('q_conv', 'k_conv')].grad)) for n, _ in m.named_parameters()], ref_of, gmax)
        errs = sorted(float(ln.split()[-1]) for ln in lines)
        print('\n'.join('%s   (fp32 eager: %s)' % (a, b.split()[-1]) for a, b in zip(lines, lines32) if float(a.split()[-1]) > 5e-4))
        print('NaivePCT 512 x 512 (%s) parameter gradients vs fp64 eager CUDA autograd: worst %.2e, median tensor %.2e  (fp32 eager: worst %.2e)'
              % (init, worst, errs[len(errs) // 2], worst32))
        if init == 'stress':
            assert worst < 1e-1 and errs[len(errs) // 2] < 1e-2
        else:
            # atomics make the sums differ in the last bits from run to run, and a max-pool / ReLU decision that flips moves one
            # tensor by ~2e-3 at this batch size: the worst tensor gets head-room, the median is the stable statistic
            assert worst < max(5e-3, 3 * worst32) and errs[len(errs) // 2] < 2e-4
    finally:
        F_.dropout = orig_dropout
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old


def test_backward_twice_with_retain_graph(dev):
    """The reference's trainer calls ``loss.backward(retain_graph=True)``: a second pass over the same graph must find the
    saved activations untouched and accumulate the same gradients again; frozen parameters receive none."""
    from sgaligner_b200.pct import NaivePCT
    torch.manual_seed(2)
    m = NaivePCT().to(dev).train()
    m.linear1.weight.requires_grad_(False)
    x = torch.randn(6, 130, 3, device=dev)
    y = m(x)
    loss = (y * y).sum()
    loss.backward(retain_graph=True)
    g1 = {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None}
    loss.backward()
    torch.cuda.synchronize()
    assert 'linear1.weight' not in g1 and m.linear1.weight.grad is None
    for n, p in m.named_parameters():
        if n in g1:
            assert rel_inf(p.grad, 2 * g1[n]) < 1e-5 or float(g1[n].abs().max()) == 0.0, n      # atomics: not bit-identical


def test_encoder_with_pct_trains(dev):
    """MultiModalEncoder(['pct','gat','rel','attr']) -- the module list of the shipped config
    (configs/scan3r/scan3r_ground_truth.yaml:5) -- through OverallLoss, backward and Adam: finite gradients on every
    parameter of the point encoder and a decreasing loss."""
    from sgaligner_b200 import synthetic, to_cuda
    from sgaligner_b200.losses import CustomMultiLossLayer, OverallLoss
    from sgaligner_b200.sg_aligner import MultiModalEncoder
    from sgaligner_b200.trainer import FlatAdam, train_step
    torch.manual_seed(1)
    modules = ['pct', 'gat', 'rel', 'attr']
    data = to_cuda(synthetic.make_batch([10] * 4, [12] * 4, [6] * 4, n_points=160, edge_mode='complete', seed=3), dev)
    model = MultiModalEncoder(modules=modules, rel_dim=41, attr_dim=164).to(dev).train()
    li, lc = CustomMultiLossLayer(4).to(dev), CustomMultiLossLayer(4).to(dev)
    fn = OverallLoss(li, lc, dev, {'zoom': 0.1, 'wt_align_loss': 1.0, 'wt_contrastive_loss': 1.0, 'modules': modules})
    opt = FlatAdam(list(model.parameters()) + list(li.parameters()) + list(lc.parameters()), lr=1e-3, weight_decay=1e-6)
    losses = [float(train_step(model, fn, opt, data)['loss']) for _ in range(10)]
    for name, p in model.object_encoder.named_parameters():
        assert p.grad is not None and bool(torch.isfinite(p.grad).all()), name
    print('pct training losses', ['%.3f' % l for l in losses])
    assert all(np.isfinite(losses))
    assert min(losses[-3:]) < losses[0]
