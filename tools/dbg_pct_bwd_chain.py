"""Debug: the residual-stream gradients G(x_l) of the NaivePCT backward against fp64 autograd of the oracle."""
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, '.')
from oracle import pct_oracle as PO  # noqa: E402
from sgaligner_b200 import ops  # noqa: E402
from sgaligner_b200.pct import NaivePCT  # noqa: E402


def rel(a, b):
    return float((a.double().cpu() - b.double().cpu()).abs().max() / b.double().abs().max().clamp_min(1e-300))


N, P, training = (int(sys.argv[1]), int(sys.argv[2]), sys.argv[3] == '1') if len(sys.argv) > 3 else (5, 300, True)
dev = torch.device('cuda:0')
odev = dev if (len(sys.argv) > 4 and sys.argv[4] == 'gpu') else torch.device('cpu')      # where the fp64 oracle runs
PSEED = int(sys.argv[5]) if len(sys.argv) > 5 else 13
p = PO.random_params(PSEED)
m = NaivePCT()
m.load_state_dict(p, strict=True)
m = m.to(dev).train(training)
m.dropout_rng = 'cpu'
g = torch.Generator().manual_seed(3)
x = torch.randn(N, P, 3, generator=g) * 0.7 + torch.rand(N, 1, 3, generator=g) * 2 - 1
R = torch.randn(N, 256, generator=g)

rec = []
orig_bn = ops.bn_backward


def bn(g_, y, ab, bnm, stats, cnt, tr, **kw):
    out = orig_bn(g_, y, ab, bnm, stats, cnt, tr, **kw)
    if y.dim() == 3:
        rec.append((g_.detach().clone(), out[0].detach().clone() if out[0] is not None else None, out[1].clone(), out[2].clone()))
    return out


ops.bn_backward = bn
SAVED = {}
_orig_bwd = NaivePCT._backward


def _bwd(self, S, g_out, tr):
    SAVED['layers'] = [dict(L) for L in S['layers']]
    SAVED['z2'] = S['z2']
    return _orig_bwd(self, S, g_out, tr)


NaivePCT._backward = _bwd
MASKS = [(torch.rand(N, 512, generator=g) < 0.5).float().to(dev), (torch.rand(N, 256, generator=g) < 0.5).float().to(dev)]
_it = iter(MASKS)
m._mask = lambda n_, c_, d_: next(_it)
y = m(x.to(dev))
(y * R.to(dev)).sum().backward()
torch.cuda.synchronize()

# oracle, layer by layer, fp64
po = {k: (v.clone().double().to(odev).requires_grad_('running' not in k) if v.is_floating_point() else v.clone().to(odev)) for k, v in p.items()}
for sa in ('sa1', 'sa2', 'sa3', 'sa4'):
    po[sa + '.q_conv.weight'] = po[sa + '.k_conv.weight']
xin = x.permute(0, 2, 1).double().to(odev)
x0 = PO.embedding(xin, po, training)
x0.retain_grad()
xs = [x0]
ts = []
FWD = []
for i in (1, 2, 3, 4):
    pre = 'sa%d' % i
    xp = xs[-1]
    x_k = PO._conv(xp, po, pre + '.k_conv')
    x_v = PO._conv(xp, po, pre + '.v_conv')
    A = torch.softmax(torch.bmm(x_k.permute(0, 2, 1), x_k) / (32 ** 0.5), dim=-1)
    x_s = torch.bmm(x_v, A)
    t = PO._conv(x_s, po, pre + '.trans_conv')
    FWD.append((xp.detach(), x_k.detach(), x_v.detach(), x_s.detach(), t.detach()))
    t.retain_grad()
    ts.append(t)
    xn = xp + F.relu(PO._bn(t, po, pre + '.after_norm', training))
    xn.retain_grad()
    xs.append(xn)
xc = torch.cat(xs[1:], dim=1)
z = F.leaky_relu(PO._bn(PO._conv(xc, po, 'linear.0'), po, 'linear.1', training), 0.2)
h = torch.max(z, dim=-1)[0]
h = F.relu(PO._bn(h @ po['linear1.weight'].t(), po, 'bn1', training))
h = h * MASKS[0].double().to(odev) * 2.0 if training else h
h = F.relu(PO._bn(h @ po['linear2.weight'].t() + po['linear2.bias'], po, 'bn2', training))
yo = h * MASKS[1].double().to(odev) * 2.0 if training else h
print('forward', rel(y, yo))
(yo * R.double().to(odev)).sum().backward()
# rec order: layer 4, 3, 2, 1 (G(x4) .. G(x1)), then embedding bn2 (G(x0))
for idx, l in enumerate((4, 3, 2, 1, 0)):
    g_ours, dt_ours, dga, dbe = rec[idx]
    g_ref = xs[l].grad.permute(0, 2, 1)
    line = 'G(x%d): %.2e  (|G|max %.2e)' % (l, rel(g_ours, g_ref), float(g_ref.abs().max()))
    if l >= 1:
        line += '   dt%d: %.2e   colsum(G) %.2e' % (l, rel(dt_ours, ts[l - 1].grad.permute(0, 2, 1)),
                                                   rel(g_ours.double().sum((0, 1)).cpu(), g_ref.sum((0, 1))))
        pre = 'sa%d' % l
        line += '  dbeta %.2e dgamma %.2e' % (rel(dbe, po[pre + '.after_norm.bias'].grad), rel(dga, po[pre + '.after_norm.weight'].grad))
    print(line)

print('forward intermediates, ours vs fp64 (rel. to the tensor max / rel. to each row norm, worst row):')
for li in range(4):
    L = SAVED['layers'][li]
    xp, x_k, x_v, x_s, t = FWD[li]
    def both(a, b):
        a = a.double()
        b = b.permute(0, 2, 1).to(a.device)
        d = (a - b)
        return '%.1e / %.1e' % (float(d.abs().max() / b.abs().max()), float((d.norm(dim=-1) / b.norm(dim=-1).clamp_min(1e-30)).max()))
    print('layer %d: x_in %s   k %s   v %s   x_s %s   t %s   |k|max %.0f' % (li + 1, both(L['x_in'], xp), both(L['k'], x_k), both(L['v'], x_v),
                                                                        both(L['x_s'], x_s), both(L['t'], t), float(x_k.abs().max())))
