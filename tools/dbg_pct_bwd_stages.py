"""Debug: run the NaivePCT backward on a test case and check every building-block call against fp64 on ITS OWN inputs."""
import math
import sys

import torch

sys.path.insert(0, '.')
from oracle import pct_oracle  # noqa: E402
from sgaligner_b200 import ops  # noqa: E402
from sgaligner_b200.pct import NaivePCT  # noqa: E402


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-300))


N, P, training = (int(sys.argv[1]), int(sys.argv[2]), sys.argv[3] == '1') if len(sys.argv) > 3 else (4, 200, False)
PSEED = int(sys.argv[4]) if len(sys.argv) > 4 else 13
dev = torch.device('cuda:0')
p = pct_oracle.random_params(PSEED)
m = NaivePCT()
m.load_state_dict(p, strict=True)
m = m.to(dev).train(training)
m.dropout_rng = 'cpu'
g = torch.Generator().manual_seed(3)
x = torch.randn(N, P, 3, generator=g) * 0.7 + torch.rand(N, 1, 3, generator=g) * 2 - 1
R = torch.randn(N, 256, generator=g)

orig_attn, orig_pw, orig_wg = ops.pct_attention_backward, ops.pct_pointwise_grad, ops.wgrad_group


def attn(k, v, c2, dxs):
    dk1, dk2, dv, dvc, dvm = orig_attn(k, v, c2, dxs)
    torch.cuda.synchronize()
    with torch.enable_grad():
        kd = k.double().requires_grad_(True)
        vd = v.double().requires_grad_(True)
        A = torch.softmax(kd @ kd.transpose(1, 2) / math.sqrt(32), dim=-1)
        ((A.transpose(1, 2) @ vd) * dxs.double()).sum().backward()
    print('attn_bwd: |k|max %.1f energy max %.0f  |dxs|max %.2e  dv %.1e dk %.1e  (|dk|max %.2e)' % (
        float(k.abs().max()), float((k.double() ** 2).sum(-1).max() / math.sqrt(32)), float(dxs.abs().max()), rel(dv, vd.grad),
        rel(dk1 + dk2, kd.grad), float(kd.grad.abs().max())))
    return dk1, dk2, dv, dvc, dvm


def pw(src, Wt, absmax=None):
    out = orig_pw(src, Wt, absmax)
    torch.cuda.synchronize()
    print('pointwise_grad: %.1e' % rel(out, src.double() @ Wt.double().t()))
    return out


def wg(problems):
    before = [c.clone() for _, _, c in problems]
    orig_wg(problems)
    torch.cuda.synchronize()
    for (a, b, c), c0 in zip(problems, before):
        print('wgrad [%d x %d]: %.1e' % (a.shape[1], b.shape[1], rel(c, c0.double() + a.double().t() @ b.double())))


orig_ig, orig_bn, orig_dense = ops.pct_sa_input_grad, ops.bn_backward, ops.pct_cat_dense_backward


def ig(gx, gcat, dxv, dk1, dk2, Wk):
    ref = gx.double() + dxv.double() + (dk1.double() + dk2.double()) @ Wk.double()
    if gcat is not None:
        ref = ref + gcat.double()
    out = orig_ig(gx, gcat, dxv, dk1, dk2, Wk)
    torch.cuda.synchronize()
    print('sa_input_grad: %.1e   column sums %.1e' % (rel(out, ref), rel(out.double().sum((0, 1)), ref.sum((0, 1)))))
    return out


def dense(x1, x2, x3, x4, M, u, xbar):
    gs = orig_dense(x1, x2, x3, x4, M, u, xbar)
    torch.cuda.synchronize()
    xc = torch.cat([x1, x2, x3, x4], -1).double() - xbar.double()
    ref = -(xc @ M.double().t()) - u.double()
    got = torch.cat(gs, -1).double()
    print('cat dense: %.1e   column sums %.1e' % (rel(got, ref), rel(got.sum((0, 1)), ref.sum((0, 1)))))
    return gs


ops.pct_attention_backward, ops.pct_pointwise_grad, ops.wgrad_group = attn, pw, wg
ops.pct_sa_input_grad, ops.pct_cat_dense_backward = ig, dense
torch.manual_seed(77)
y = m(x.to(dev))
(y * R.to(dev)).sum().backward()
torch.cuda.synchronize()
