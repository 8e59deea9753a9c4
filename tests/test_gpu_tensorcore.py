"""tcgen05 bring-up and the tensor-core PointNet kernel vs the oracle and vs the fp32 FMA kernel."""
import pytest
import torch

from oracle import sgaligner_oracle as O
from tests.util import CASES, load_case, rel_inf

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available()
    return torch.device('cuda:0')


@pytest.mark.parametrize('kind,K,ncols', [(0, 64, 128), (0, 128, 128), (0, 128, 64), (1, 32, 128), (1, 64, 128), (1, 64, 48)])
def test_umma_selftest(kind, K, ncols, dev):
    """One split-operand UMMA tile (bf16x3 / tf32x3) against an fp64 matmul: validates the
    shared-memory descriptors, the swizzled operand layout, the instruction descriptor and the
    TMEM load shape that every tensor-core kernel of the library shares."""
    from sgaligner_b200 import ops
    g = torch.Generator().manual_seed(kind * 100 + K + ncols)
    A = torch.randn(128, K, generator=g)
    B = torch.randn(ncols, K, generator=g)
    D = ops.selftest_umma(A.to(dev), B.to(dev), kind)
    torch.cuda.synchronize()
    ref = A.double() @ B.double().t()
    err = rel_inf(D, ref)
    assert err < (3e-5 if kind == 0 else 3e-6), err


@pytest.mark.parametrize('name', CASES)
def test_pointnet_tc_vs_oracle(name, dev):
    from sgaligner_b200 import ops
    c = load_case(name)
    p = c['params']
    pts = c['data']['tot_obj_pts']
    ref64 = O.pointnet_feat(pts.double(), {k: v.double() for k, v in p.items() if v.is_floating_point()})
    w = [p[f'object_encoder.conv{i}.{k}'].to(dev) for i in (1, 2, 3) for k in ('weight', 'bias')]
    out, arg = ops.pointnet_forward(pts.to(dev), *w, want_argmax=True, mode=ops.POINTNET_TC)
    out2, _ = ops.pointnet_forward(pts.to(dev), *w, want_argmax=False, mode=ops.POINTNET_TC)
    simt, arg_s = ops.pointnet_forward(pts.to(dev), *w, want_argmax=True, mode=ops.POINTNET_SIMT)
    torch.cuda.synchronize()
    assert rel_inf(out, ref64) < 3e-5
    # the argmax-tracking variant re-evaluates near-ties in fp32; it only touches the value at a ReLU-kink flip
    assert float((out - out2).abs().max()) <= 1e-4 * float(out2.abs().max())
    assert float((out != out2).float().mean()) < 1e-3
    assert rel_inf(out, simt) < 3e-5
    live = (simt > 1e-3 * simt.max()).cpu()
    assert int(arg.min()) >= 0 and int(arg.max()) < pts.shape[1]
    # same argmax as the fp32 kernel wherever the channel is alive (ties / near-ties may differ)
    agree = (arg.cpu() == arg_s.cpu())[live].float().mean()
    assert float(agree) > 0.97


def test_pointnet_tc_ragged_sizes(dev):
    """P not a multiple of the 128-point tile, N not a multiple of the SM count, C3 = 128 / 512."""
    from sgaligner_b200 import ops
    for (N, P, C3, seed) in [(5, 100, 256, 0), (301, 130, 128, 1), (17, 1, 256, 2), (40, 513, 512, 3)]:
        p = O.init_params(['point'], 41, 164, pt_out_dim=C3, seed=seed)
        for i in (1, 2, 3):
            p[f'object_encoder.conv{i}.bias'] = 0.1 * torch.randn(p[f'object_encoder.conv{i}.bias'].shape)
        pts = torch.randn(N, P, 3) + torch.rand(N, 1, 3) * 4 - 2
        ref = O.pointnet_feat(pts, p)
        w = [p[f'object_encoder.conv{i}.{k}'].to(dev) for i in (1, 2, 3) for k in ('weight', 'bias')]
        out, _ = ops.pointnet_forward(pts.to(dev), *w, want_argmax=False, mode=ops.POINTNET_TC)
        torch.cuda.synchronize()
        assert rel_inf(out, ref) < 3e-5, (N, P, C3)


def test_pointnet_tc_fused_bn_statistics(dev):
    """Train-mode forward: the statistics accumulated inside the tensor-core launch (conv3 thread-local,
    conv2 warp transpose-reduce, conv1 analytic from the point moments) against the oracle's fp64 batch
    statistics and the stand-alone fp32 statistics kernel; pooled feature / argmax identical to the
    launch without statistics.  Includes ragged P (padded tile columns must not be counted) and C3 = 512
    (two channel-block CTAs per SM column: conv1/conv2 statistics counted once)."""
    from sgaligner_b200 import ops
    for (N, P, C3, seed) in [(64, 512, 256, 0), (301, 130, 128, 1), (9, 100, 512, 2), (200, 257, 256, 3)]:
        p = O.init_params(['point'], 41, 164, pt_out_dim=C3, seed=seed)
        for i in (1, 2, 3):
            p[f'object_encoder.conv{i}.bias'] = 0.1 * torch.randn(p[f'object_encoder.conv{i}.bias'].shape)
        pts = torch.randn(N, P, 3) + torch.rand(N, 1, 3) * 4 - 2
        w = [p[f'object_encoder.conv{i}.{k}'].to(dev) for i in (1, 2, 3) for k in ('weight', 'bias')]
        out, arg, mom = ops.pointnet_forward_stats(pts.to(dev), *w, want_argmax=True)
        out0, arg0 = ops.pointnet_forward(pts.to(dev), *w, want_argmax=True, mode=ops.POINTNET_TC)
        mom_simt = ops.pointnet_bn_moments(pts.to(dev), *w)
        torch.cuda.synchronize()
        assert torch.equal(out, out0) and torch.equal(arg, arg0)
        stats = O.pointnet_bn_batch_stats(pts.double(), {k: v.double() for k, v in p.items() if v.is_floating_point()})
        cnt = float(N * P)
        o = 0
        for i, c in ((1, 64), (2, 128), (3, C3)):
            for m in (mom, mom_simt):
                s, sq = m[o:o + c].cpu(), m[o + c:o + 2 * c].cpu()
                mean = s / cnt
                var = (sq - s * mean) / (cnt - 1.0)
                rmean, rvar = stats[i - 1]
                assert rel_inf(mean, rmean) < 2e-5, (N, P, C3, i)
                assert rel_inf(var, rvar) < 2e-5, (N, P, C3, i)
            o += 2 * c


def test_pointnet_bwd_tc_vs_fma_kernel(dev):
    """Tensor-core backward (tcgen05 GEMMs over 128-instance tiles) against the fp32 FMA backward on the SAME
    forward results (same argmax), so the two differ by bf16x3 rounding only."""
    from sgaligner_b200 import ops
    for (N, P, C3, seed) in [(7, 64, 128, 0), (301, 130, 256, 1), (600, 512, 256, 2), (40, 257, 512, 3)]:
        p = O.init_params(['point'], 41, 164, pt_out_dim=C3, seed=seed)
        g = torch.Generator().manual_seed(seed)
        for i in (1, 2, 3):
            p[f'object_encoder.conv{i}.bias'] = 0.1 * torch.randn(p[f'object_encoder.conv{i}.bias'].shape, generator=g)
        pts = (torch.randn(N, P, 3, generator=g) + torch.rand(N, 1, 3, generator=g) * 4 - 2).to(dev)
        w = [p[f'object_encoder.conv{i}.{k}'].to(dev) for i in (1, 2, 3) for k in ('weight', 'bias')]
        out, arg = ops.pointnet_forward(pts, *w, want_argmax=True, mode=ops.POINTNET_TC)
        gout = torch.randn(N, C3, generator=g).to(dev)
        a = ops.pointnet_backward(pts, *w, out, arg, gout, mode=ops.POINTNET_TC)
        b = ops.pointnet_backward(pts, *w, out, arg, gout, mode=ops.POINTNET_SIMT)
        torch.cuda.synchronize()
        for name, x, y in zip(('gW1', 'gb1', 'gW2', 'gb2', 'gW3', 'gb3'), a, b):
            assert rel_inf(x, y) < 5e-5, (N, P, C3, name, rel_inf(x, y))


@pytest.mark.parametrize('name', CASES)
def test_match_topk_tc_vs_fma_path(name, dev):
    """Fused tcgen05 Gram + top-k vs the fp32 FMA kernels on the golden embedding: similarity within
    2e-6, top-k columns identical wherever adjacent similarities differ by more than 1e-5."""
    import numpy as np
    from sgaligner_b200 import matching
    c = load_case(name)
    key = 'joint' if len(c['modules']) > 1 else c['modules'][0]
    emb = c['out'][key].to(dev)
    a = matching.match_batch(emb, c['data'], k=6, full_rank=True, tensor_cores=True)
    b = matching.match_batch(emb, c['data'], k=6, full_rank=True, tensor_cores=False)
    torch.cuda.synchronize()
    assert float((a['sim'] - b['sim']).abs().max()) < 1e-5
    lay = a['layout']
    for p in range(lay.B):
        n = int(lay.n[p])
        o = int(lay.pair_off_host[p])
        kk = min(6, n)
        ta, tb = a['topk_idx'][o:o + n, :kk].cpu().numpy(), b['topk_idx'][o:o + n, :kk].cpu().numpy()
        da = b['topk_dist'][o:o + n, :kk].cpu().numpy()
        gaps = np.ones_like(ta, dtype=bool)
        g = np.diff(da, axis=1) > 1e-5
        gaps[:, 1:] &= g
        gaps[:, :-1] &= g
        assert (ta[gaps] == tb[gaps]).all()
        ref = c['rank'][p][:, :kk]
        srt = np.take_along_axis(c['sim'][p], c['rank'][p], 1)[:, :kk + 1]
        g2 = np.diff(srt, axis=1) > 1e-5
        ok = np.ones_like(ref, dtype=bool)
        ok &= g2[:, :kk] if g2.shape[1] >= kk else True
        ok[:, 1:] &= g2[:, :kk - 1]
        assert (ta[ok] == ref[ok]).all()


def test_match_topk_tc_large_pairs(dev):
    """Pairs with more than 128 nodes (several row / column tiles) and D not a multiple of 32."""
    import numpy as np
    from sgaligner_b200 import matching
    g = torch.Generator().manual_seed(3)
    counts = np.array([[150, 133], [64, 70], [200, 190]])
    N = int(counts.sum())
    emb = torch.randn(N, 200, generator=g).to(dev)
    data = {'graph_per_obj_count': counts}
    a = matching.match_batch(emb, data, k=8, full_rank=False, want_sim=True, tensor_cores=True)
    b = matching.match_batch(emb, data, k=8, full_rank=False, tensor_cores=False)
    torch.cuda.synchronize()
    assert float((a['sim'] - b['sim']).abs().max()) < 1e-5
    agree = (a['topk_idx'] == b['topk_idx']).float().mean()
    assert float(agree) > 0.999
    assert float((a['topk_dist'] - b['topk_dist']).abs().max()) < 1e-5


@pytest.mark.parametrize('layout', ['kk', 'kmn', 'mnmn'])
@pytest.mark.parametrize('shape', [(128, 128, 32), (200, 300, 100), (1024, 700, 400), (77, 130, 33)])
def test_gemm_tf32x3_layouts(layout, shape, dev):
    """The generic tcgen05 GEMM in all three operand-layout pairs, with row gathers, divisors, ragged
    sizes; plus the scatter-add / split-K epilogue -- against an fp64 matmul."""
    from sgaligner_b200 import ops
    M, N, K = shape
    g = torch.Generator().manual_seed(M + N + K)
    nsrc = 1500
    Z = torch.randn(nsrc, max(M, N, K) + 8, generator=g)
    div = torch.rand(nsrc, generator=g) + 0.5
    Zd = Z.double() / div.double()[:, None]
    if layout == 'kk':        # A(m,k) = Z[ia[m], k]/div ; B(n,k) = Z[ib[n], k]/div
        ia = torch.randint(0, nsrc, (M,), generator=g)
        ib = torch.randint(0, nsrc, (N,), generator=g)
        ref = Zd[ia][:, :K] @ Zd[ib][:, :K].t()
        out = ops.gemm_tf32x3(Z.to(dev), Z.to(dev), M, N, K, a_idx=ia.int().to(dev), b_idx=ib.int().to(dev), a_div=div.to(dev), b_div=div.to(dev))
    elif layout == 'kmn':     # A(m,k) = D[m,k] ; B(n,k) = Z[ib[k], n]/div
        D = torch.randn(M, K, generator=g)
        ib = torch.randint(0, nsrc, (K,), generator=g)
        ref = D.double() @ Zd[ib][:, :N]
        out = ops.gemm_tf32x3(D.to(dev), Z.to(dev), M, N, K, b_mn=True, b_idx=ib.int().to(dev), b_div=div.to(dev))
    else:                     # A(m,k) = D[k,m] ; B(n,k) = Z[ib[k], n]/div
        D = torch.randn(K, M, generator=g)
        ib = torch.randint(0, nsrc, (K,), generator=g)
        ref = D.double().t() @ Zd[ib][:, :N]
        out = ops.gemm_tf32x3(D.to(dev), Z.to(dev), M, N, K, a_mn=True, b_mn=True, b_idx=ib.int().to(dev), b_div=div.to(dev))
    torch.cuda.synchronize()
    assert rel_inf(out, ref) < 5e-6, rel_inf(out, ref)
    if layout == 'mnmn':      # scatter-add rows + split-K
        ci = torch.randint(0, 50, (M,), generator=g)
        acc = torch.zeros(50, N, device=dev)
        ops.gemm_tf32x3(D.to(dev), Z.to(dev), M, N, K, a_mn=True, b_mn=True, b_idx=ib.int().to(dev), b_div=div.to(dev), out=acc,
                        c_idx=ci.int().to(dev), ksplit=3)
        torch.cuda.synchronize()
        want = torch.zeros(50, N, dtype=torch.float64).index_add_(0, ci, ref)
        assert rel_inf(acc, want) < 1e-5


@pytest.mark.parametrize('N,P', [(37, 512), (300, 200), (5, 1000)])
def test_pointnet_bn_moments_from_gram_matrices(N, P, dev):
    """ops.pointnet_bn_moments_gram (tensor-core Gram matrices + fp64 finalize) against the fp32 FMA statistics pass
    (ops.pointnet_bn_moments): per-channel mean and (biased) variance of the three pre-ReLU conv outputs."""
    from sgaligner_b200 import ops
    g = torch.Generator().manual_seed(N)
    pts = (torch.randn(N, P, 3, generator=g) + torch.rand(N, 1, 3, generator=g) * 4 - 2).to(dev)
    C3 = 256
    W1, b1 = torch.randn(64, 3, generator=g) * 0.5, torch.randn(64, generator=g) * 0.1
    W2, b2 = torch.randn(128, 64, generator=g) * 0.15, torch.randn(128, generator=g) * 0.1
    W3, b3 = torch.randn(C3, 128, generator=g) * 0.1, torch.randn(C3, generator=g) * 0.1
    args = [t.to(dev) for t in (W1, b1, W2, b2, W3, b3)]
    ref = ops.pointnet_bn_moments(pts, *args).cpu()
    got = ops.pointnet_bn_moments_gram(pts, *args).cpu()
    n = float(N * P)
    o = 0
    for c in (64, 128, C3):
        mr, mg = ref[o:o + c] / n, got[o:o + c] / n
        vr, vg = ref[o + c:o + 2 * c] / n - mr * mr, got[o + c:o + 2 * c] / n - mg * mg
        o += 2 * c
        assert float((mg - mr).abs().max()) <= 2e-4 * float(mr.abs().max() + vr.sqrt().max()), (c, float((mg - mr).abs().max()))
        assert float(((vg - vr).abs() / vr.clamp_min(1e-12)).max()) <= 2e-3, (c, float(((vg - vr).abs() / vr.clamp_min(1e-12)).max()))
