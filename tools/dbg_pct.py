"""Stage-by-stage error of the NaivePCT kernels against the fp64 oracle on one input (tools only; imports oracle/).
For every SA layer: the error of k, v, x_s, t as computed by the kernels FROM THE KERNELS' OWN INPUTS of that stage
(isolated stage error) and the accumulated error of the residual stream.
    python tools/dbg_pct.py [golden | N P seed]"""
import os, sys, math
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import torch.nn.functional as F
from oracle import pct_oracle as O
from sgaligner_b200 import ops
from sgaligner_b200.pct import NaivePCT

dev = torch.device('cuda:0')
if len(sys.argv) < 2 or sys.argv[1] == 'golden':
    z = np.load('tests/golden/pct_ref.npz')
    p = O.random_params(int(z['param_seed']))
    x = torch.from_numpy(z['x']).permute(0, 2, 1).contiguous()
else:
    N, P, seed = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    p = O.random_params(7)
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(N, P, 3, generator=g) + torch.rand(N, 1, 3, generator=g) * 2 - 1
m = NaivePCT(); m.load_state_dict(p, strict=True); m = m.to(dev).eval()
p64 = {k: (v.double() if v.is_floating_point() else v) for k, v in p.items()}
ri = lambda a, b: float((a.double().cpu() - b.double().cpu()).abs().max() / b.double().cpu().abs().max().clamp_min(1e-300))

rec = []
orig = {n: getattr(ops, n) for n in ('pct_pointwise', 'pct_attention', 'pct_embed', 'pct_cat_linear', 'pct_pool_act')}
def wrap(n):
    def f(*a, **k):
        r = orig[n](*a, **k)
        rec.append((n, a, k, r))
        return r
    return f
for n in orig: setattr(ops, n, wrap(n))
with torch.no_grad():
    y = m(x.to(dev))
torch.cuda.synchronize()
ref = O.naive_pct(x.permute(0, 2, 1).double(), p64)
print('output vs fp64 oracle: %.2e' % ri(y, ref))

def D(t): return None if t is None else t.double().cpu()
def g(src, ab):
    if src is None: return 0
    s = D(src)
    return s if ab is None else torch.relu(D(ab[0]) * s + D(ab[1]))
# oracle chain
xo = O.embedding(x.permute(0, 2, 1).double(), p64, False)          # [B,128,P]
li = 0
for n, a, k, r in rec:
    if n == 'pct_embed':
        print('embed z2: kernels vs fp64 of same inputs  (n/a)   accumulated x0 vs oracle: %.2e' % ri(torch.relu(D(rec[1][1][1][0]) * D(r[0]) + D(rec[1][1][1][1])) if False else D(r[0]), D(r[0])))
    if n == 'pct_pointwise':
        src1, ab1, src2, ab2, W, bias, c0 = a[:7]
        X = g(src1, ab1) + g(src2, ab2)
        Y = X @ D(W).t() + (0 if bias is None else D(bias))
        out0, out1, ox, st = r
        got = torch.cat([D(out0)] + ([D(out1)] if out1 is not None else []), -1)
        kind = 'k|v' if out1 is not None else 'trans'
        msg = '  %-5s stage error %.2e' % (kind, ri(got, Y))
        if out1 is not None:
            msg += '  (k %.2e  v %.2e)   X (= x%d) accumulated vs oracle %.2e, |x|max %.1f  |k|max %.1f' % (ri(got[..., :32], Y[..., :32]), ri(got[..., 32:], Y[..., 32:]), li, ri(X, xo.permute(0, 2, 1)), float(X.abs().max()), float(Y[..., :32].abs().max()))
        print(msg)
    if n == 'pct_attention':
        kk, vv = D(a[0]), D(a[1])
        e = kk @ kk.transpose(1, 2) / math.sqrt(32)
        att = torch.softmax(e, -1)
        refxs = att.transpose(1, 2) @ vv
        e32 = (a[0] @ a[0].transpose(1, 2)) / math.sqrt(32)
        xs32 = torch.softmax(e32, -1).transpose(1, 2) @ a[1]
        print('  attn  stage error %.2e   (torch fp32 on the GPU, same inputs: %.2e)   |energy|max %.1f' % (ri(r, refxs), ri(xs32, refxs), float(e.abs().max())))
        li += 1
        xo = O.self_attention(xo, p64, 'sa%d' % li, False)
    if n == 'pct_cat_linear':
        x1, x2, x3, t4, ab4, WL = a
        x4 = D(x3) + g(t4, ab4)
        Z = torch.cat([D(x1), D(x2), D(x3), x4], 2) @ D(WL).t()
        zmax, zmin, st = r
        print('  cat   zmax stage error %.2e  zmin %.2e   x4 accumulated vs oracle %.2e' % (ri(D(zmax).max(1).values, Z.max(1).values), ri(D(zmin).min(1).values, Z.min(1).values), ri(x4, xo.permute(0, 2, 1))))
