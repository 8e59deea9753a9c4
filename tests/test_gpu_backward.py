"""GPU parity tests (backward / training step): analytic CUDA backward kernels vs the parameter
gradients the unmodified reference produced through autograd (tests/golden/*.npz)."""
import numpy as np
import pytest
import torch

from tests.util import CASES, grad_close, load_case, rel_inf

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available()
    return torch.device('cuda:0')


def _setup(case, dev, pointnet_mode=None):
    from sgaligner_b200 import to_cuda
    from sgaligner_b200.losses import CustomMultiLossLayer, OverallLoss
    from sgaligner_b200.sg_aligner import MultiModalEncoder
    model = MultiModalEncoder(modules=case['modules'], rel_dim=41, attr_dim=164)
    model.load_state_dict(case['params'], strict=True)
    model = model.to(dev)
    if pointnet_mode is not None:
        model.object_encoder.kernel_mode = pointnet_mode
    M = len(case['modules'])
    li, lc = CustomMultiLossLayer(M), CustomMultiLossLayer(M)
    with torch.no_grad():
        li.log_vars.copy_(case['lv'][0])
        lc.log_vars.copy_(case['lv'][1])
    li, lc = li.to(dev), lc.to(dev)
    fn = OverallLoss(li, lc, dev, {'zoom': 0.1, 'wt_align_loss': 1.0, 'wt_contrastive_loss': 1.0, 'modules': case['modules']})
    return model, li, lc, fn, to_cuda(dict(case['data']), dev)


@pytest.mark.parametrize('mode', ['simt', 'tc'])
@pytest.mark.parametrize('name', CASES)
def test_param_grads_vs_golden(name, mode, dev):
    from sgaligner_b200 import ops
    c = load_case(name)
    model, li, lc, fn, data = _setup(c, dev, ops.POINTNET_SIMT if mode == 'simt' else ops.POINTNET_TC)
    model.train()
    ops.pointnet_tie_stats(reset=True)
    out = model(data)
    ld = fn(out, data)
    ld['loss'].backward()
    torch.cuda.synchronize()
    if mode == 'tc' and 'point' in c['modules']:
        n_tie, n_flip = ops.pointnet_tie_stats()
        n_entries = data['tot_obj_pts'].shape[0] * model.pt_out_dim
        print(f'[{name}] max-pool near-ties re-evaluated in fp32: {n_tie} of {n_entries} entries, {n_flip} reordered')
    assert abs(float(ld['loss'].detach()) - c['loss']['loss']) <= 1e-3 * abs(c['loss']['loss'])
    named = dict(model.named_parameters())
    # tensor-core mode: near-ties of the max-pool are re-evaluated in fp32 (pointnet_tie_fix_kernel) and the ReLU
    # kink of the recomputed conv2 likewise (pointnet_bwd_tc.cu), so the gradient is routed exactly as in fp32:
    # the SAME tolerance as the fp32 FMA kernels, for every parameter
    tol = 1e-3
    atol = 1e-6 * max(float(g.abs().max()) for g in c['grad'].values())
    for k, ref in c['grad'].items():
        if k == '__lv_ial':
            got = li.log_vars.grad
        elif k == '__lv_icl':
            got = lc.log_vars.grad
        else:
            got = named[k].grad
        if got is None:
            assert float(ref.abs().max()) == 0.0, k
            continue
        assert grad_close(got, ref, rtol=tol, atol=atol), (k, rel_inf(got, ref))
    # BatchNorm parameters receive no gradient in the reference either (outputs discarded)
    for i in (1, 2, 3):
        assert named[f'object_encoder.bn{i}.weight'].grad is None


@pytest.mark.parametrize('name', ['mid4', 'c1_example'])
def test_tc_argmax_is_the_fp32_argmax(name, dev):
    """The argmax the tensor-core forward hands to the backward equals the fp32 FMA kernel's wherever the channel is
    active (dead channels carry no gradient); the measured number of near-ties / reorderings is printed."""
    from sgaligner_b200 import ops
    c = load_case(name)
    pts = c['data']['tot_obj_pts'].to(dev)
    w = [c['params'][f'object_encoder.conv{i}.{k}'].to(dev) for i in (1, 2, 3) for k in ('weight', 'bias')]
    ops.pointnet_tie_stats(reset=True)
    o_tc, a_tc = ops.pointnet_forward(pts, *w, want_argmax=True, mode=ops.POINTNET_TC)
    n_tie, n_flip = ops.pointnet_tie_stats()
    o_fp, a_fp = ops.pointnet_forward(pts, *w, want_argmax=True, mode=ops.POINTNET_SIMT)
    torch.cuda.synchronize()
    active = o_fp > 0
    diff = (a_tc != a_fp) & active
    # a point that differs must be an exact duplicate (resampled objects) or an fp32-level tie
    bad = 0
    if bool(diff.any()):
        nn_, cc_ = diff.nonzero(as_tuple=True)
        pa, pb = pts[nn_, a_tc[nn_, cc_].long()], pts[nn_, a_fp[nn_, cc_].long()]
        bad = int(((pa - pb).abs().max(1).values > 0).sum())
    print(f'[{name}] near-ties {n_tie}, reordered {n_flip}, argmax differing from fp32 on distinct points: {bad} of {int(active.sum())}')
    assert bad <= max(1, int(1e-6 * active.sum()))
    assert float((o_tc - o_fp).abs().max()) <= 1e-4 * float(o_fp.abs().max())


@pytest.mark.parametrize('name', ['small4', 'mid4'])
def test_bn_running_stats_side_effect(name, dev):
    c = load_case(name)
    model, li, lc, fn, data = _setup(c, dev)
    model.train()
    with torch.no_grad():
        model(data)
    torch.cuda.synchronize()
    sd = model.state_dict()
    for k, ref in c['bn'].items():
        if 'num_batches' in k:
            assert int(sd[k]) == int(ref)
        else:
            assert rel_inf(sd[k], ref) < 1e-4, k


def test_eval_mode_leaves_bn_untouched(dev):
    c = load_case('small4')
    model, *_, data = _setup(c, dev)
    model.eval()
    with torch.no_grad():
        model(data)
    sd = model.state_dict()
    assert int(sd['object_encoder.bn1.num_batches_tracked']) == 0


def test_flat_adam_matches_torch_adam(dev):
    from sgaligner_b200.trainer import FlatAdam
    torch.manual_seed(0)
    shapes = [(64, 3), (64,), (100, 256), (1, 2, 128), (4, 1)]
    ps_ref = [torch.randn(s).requires_grad_(True) for s in shapes]
    ps = [torch.nn.Parameter(p.detach().clone().to(dev)) for p in ps_ref]
    ref_opt = torch.optim.Adam(ps_ref, lr=1e-3, weight_decay=1e-6)
    opt = FlatAdam(ps, lr=1e-3, weight_decay=1e-6)
    for step in range(5):
        gs = [torch.randn(s) * (0.1 + step) for s in shapes]
        for p, g in zip(ps_ref, gs):
            p.grad = g.clone()
        opt.zero_grad()
        for p, g in zip(ps, gs):
            p.grad.copy_(g.to(dev))
        ref_opt.step()
        opt.step()
    torch.cuda.synchronize()
    for p, r in zip(ps, ps_ref):
        assert rel_inf(p.detach(), r.detach()) < 1e-5


def test_train_steps_reduce_loss(dev):
    """Whole step through the public API: forward, loss, backward, Adam -- loss must go down."""
    from sgaligner_b200 import synthetic, to_cuda
    from sgaligner_b200.losses import CustomMultiLossLayer, OverallLoss
    from sgaligner_b200.sg_aligner import MultiModalEncoder
    from sgaligner_b200.trainer import FlatAdam, train_step
    torch.manual_seed(1)
    modules = ['point', 'gat', 'rel', 'attr']
    data = to_cuda(synthetic.make_batch([10] * 4, [12] * 4, [6] * 4, n_points=128, edge_mode='complete', seed=3), dev)
    model = MultiModalEncoder(modules=modules, rel_dim=41, attr_dim=164).to(dev)
    li, lc = CustomMultiLossLayer(4).to(dev), CustomMultiLossLayer(4).to(dev)
    fn = OverallLoss(li, lc, dev, {'zoom': 0.1, 'wt_align_loss': 1.0, 'wt_contrastive_loss': 1.0, 'modules': modules})
    opt = FlatAdam(list(model.parameters()) + list(li.parameters()) + list(lc.parameters()), lr=1e-3, weight_decay=1e-6)
    losses = [float(train_step(model, fn, opt, data)['loss']) for _ in range(12)]
    assert all(np.isfinite(losses))
    assert losses[-1] < losses[0]


def test_direct_grad_accumulation_matches_autograd(dev):
    """For parameters opted in with ``ops.enable_direct_grad`` (FlatAdam does it for its flat buffer) the backward
    kernels accumulate parameter gradients straight into the kept-allocated ``.grad`` tensors instead of returning
    fresh tensors for autograd to add: same values, and a second backward accumulates on top exactly like autograd
    would.  Without the opt-in, kept-allocated ``.grad`` tensors are left to autograd (torch.autograd.grad works)."""
    from sgaligner_b200 import synthetic, to_cuda
    from sgaligner_b200.losses import CustomMultiLossLayer, OverallLoss
    from sgaligner_b200.sg_aligner import MultiModalEncoder
    mods = ['point', 'gat', 'rel', 'attr']
    data = to_cuda(synthetic.make_batch([9, 12, 7], [11, 8, 10], [5, 6, 4], n_points=128, edge_mode='complete', seed=5), dev)
    torch.manual_seed(0)
    model = MultiModalEncoder(modules=mods, rel_dim=41, attr_dim=164).to(dev).eval()
    li, lc = CustomMultiLossLayer(4).to(dev), CustomMultiLossLayer(4).to(dev)
    fn = OverallLoss(li, lc, dev, {'zoom': 0.1, 'wt_align_loss': 1.0, 'wt_contrastive_loss': 1.0, 'modules': mods})
    params = [p for p in model.parameters() if p.requires_grad]

    def backward():
        fn(model(data), data)['loss'].backward()
        torch.cuda.synchronize()

    assert all(p.grad is None for p in params)
    backward()                                          # fallback: autograd receives tensors
    ref = [None if p.grad is None else p.grad.detach().clone() for p in params]
    for p in params:
        if p.grad is not None:
            p.grad.zero_()                              # kept allocated, but NOT opted in: autograd still sees gradients
    live = [p for p in params if p.grad is not None]
    gs = torch.autograd.grad(fn(model(data), data)['loss'], live, allow_unused=True)
    for p, g_ in zip(live, gs):
        assert g_ is not None and float(p.grad.abs().max()) == 0.0
    from sgaligner_b200 import ops
    ops.enable_direct_grad(params)                      # explicit opt-in: the next backward goes direct
    ptrs = [None if p.grad is None else p.grad.data_ptr() for p in params]
    backward()
    for p, r, q in zip(params, ref, ptrs):
        if r is None:
            continue
        assert p.grad.data_ptr() == q                   # accumulated in place
        assert grad_close(p.grad, r, rtol=1e-4, atol=1e-7 * float(r.abs().max()) + 1e-12)
    backward()                                          # no zeroing: gradients accumulate
    for p, r in zip(params, ref):
        if r is not None:
            assert grad_close(p.grad, 2 * r, rtol=1e-4, atol=2e-7 * float(r.abs().max()) + 1e-12)
    ops.enable_direct_grad(params, on=False)


def test_flat_adam_skips_gradless_parameters_and_speaks_torch_state_dict(dev):
    """torch.optim.Adam leaves a parameter whose ``.grad`` is None alone (modules that are not selected, the
    BatchNorm affine parameters); FlatAdam recognises it as an all-zero gradient segment and skips it too -- no
    weight-decay drift.  Its state_dict has torch's layout: a torch.optim.Adam over the same parameters loads it
    and continues identically, and vice versa; ``param_groups[0]['lr']`` is live (lr schedulers)."""
    from sgaligner_b200.trainer import FlatAdam
    torch.manual_seed(3)
    shapes = [(64, 3), (7,), (100, 41), (5, 5)]
    ps_ref = [torch.randn(s).requires_grad_(True) for s in shapes]
    ps = [torch.nn.Parameter(p.detach().clone().to(dev)) for p in ps_ref]
    ref_opt = torch.optim.Adam(ps_ref, lr=1e-2, weight_decay=1e-2)
    opt = FlatAdam(ps, lr=1e-2, weight_decay=1e-2)
    for step in range(3):
        gs = [torch.randn(s) for s in shapes]
        opt.zero_grad()
        for i, (p, pr, g) in enumerate(zip(ps, ps_ref, gs)):
            if i == 2:                       # this one never receives a gradient
                pr.grad = None
                continue
            pr.grad = g.clone()
            p.grad.copy_(g.to(dev))
        ref_opt.step()
        opt.step()
    torch.cuda.synchronize()
    assert torch.equal(ps[2].detach().cpu(), ps_ref[2].detach())          # untouched, bit for bit
    for p, r in zip(ps, ps_ref):
        assert rel_inf(p.detach(), r.detach()) < 1e-5
    # state_dict round trips in both directions
    sd = opt.state_dict()
    assert set(sd) == {'state', 'param_groups'} and sd['param_groups'][0]['params'] == [0, 1, 2, 3]
    sd_ref = ref_opt.state_dict()
    for i in (0, 1, 3):
        assert rel_inf(sd['state'][i]['exp_avg'], sd_ref['state'][i]['exp_avg']) < 1e-5
    opt2 = FlatAdam([torch.nn.Parameter(p.detach().clone()) for p in ps], lr=5.0)
    opt2.load_state_dict(sd_ref)             # torch's own checkpoint (no entry for the grad-less parameter)
    assert opt2.step_count == 3 and abs(opt2.param_groups[0]['lr'] - 1e-2) < 1e-12
    assert rel_inf(opt2.exp_avg_sq[opt2.offsets[3]:opt2.offsets[3] + 25], sd_ref['state'][3]['exp_avg_sq'].reshape(-1)) < 1e-6
    sched = torch.optim.lr_scheduler.StepLR(ref_opt, step_size=1, gamma=0.5)   # lr is read from param_groups
    opt.param_groups[0]['lr'] *= 0.5
    sched.step()
    assert abs(opt.lr - ref_opt.param_groups[0]['lr']) < 1e-12
