"""CPU emulation of the NaivePCT kernels' split-operand numerics (fp64 arithmetic, operands quantised to bf16 hi+lo, the
lo*lo partial product dropped where the kernel drops it), with the quantisation switchable per stage: shows which stage
dominates the end-to-end error against the fp64 oracle on the golden input.  Test infrastructure (imports oracle/)."""
import math, sys
import numpy as np, torch
import torch.nn.functional as F
from oracle import pct_oracle as O

FMT=[torch.bfloat16]
def bf(x):  return x.float().to(FMT[0]).double()
def split(x, parts=2):
    out = []; r = x.clone()
    for _ in range(parts):
        h = bf(r); out.append(h); r = r - h
    return out
def mm3(a, b, drop_lolo=True, parts=2):
    """a[...,m,k] @ b[...,k,n] with split operands"""
    A = split(a, parts); B = split(b, parts)
    acc = 0
    for i, x in enumerate(A):
        for j, y in enumerate(B):
            if drop_lolo and parts == 2 and i == 1 and j == 1: continue
            if parts == 3 and i + j > 2: continue
            acc = acc + x @ y
    return acc
def f32(x): return x.float().double()

def run(x, p, q):           # q: dict stage -> mode
    def conv(h, w, stage, bias=None):   # h [B,C,N]
        W = p[w][:, :, 0]
        if q.get(stage):
            y = mm3(W, h, parts=q[stage]) if q[stage] != 'f32' else f32(W @ h)
        else:
            y = W @ h
        if bias is not None: y = y + p[bias][None, :, None]
        return f32(y) if q.get('store') else y
    def bn(h, name):
        return F.batch_norm(h, p[name+'.running_mean'], p[name+'.running_var'], p[name+'.weight'], p[name+'.bias'], False, 0.1, 1e-5)
    h = F.relu(bn(p['embedding.conv1.weight'][:, :, 0] @ x, 'embedding.bn1'))
    h = F.relu(bn(conv(h, 'embedding.conv2.weight', 'emb'), 'embedding.bn2'))
    xs = []
    for i in (1, 2, 3, 4):
        s = f'sa{i}'
        k = conv(h, s+'.k_conv.weight', 'kv'); v = conv(h, s+'.v_conv.weight', 'kv', s+'.v_conv.bias')
        kt = k.permute(0, 2, 1)
        mode = q.get('energy')
        if mode == 'f32': e = f32(kt.float() @ k.float()).double()
        elif mode: e = mm3(kt, k, drop_lolo=False, parts=mode)
        else: e = kt @ k
        e = e / math.sqrt(32)
        att = torch.softmax(e, -1)
        if q.get('pv'): o = mm3(v, att, parts=q['pv'])
        else: o = v @ att
        if q.get('store'): o = f32(o)
        t = conv(o, s+'.trans_conv.weight', 'trans', s+'.trans_conv.bias')
        h = h + F.relu(bn(t, s+'.after_norm'))
        if q.get('store'): h = f32(h)
        xs.append(h)
    c = torch.cat(xs, 1)
    z = conv(c, 'linear.0.weight', 'lin')
    z = F.leaky_relu(bn(z, 'linear.1'), 0.2).max(-1)[0]
    z = F.relu(bn(z @ p['linear1.weight'].t(), 'bn1'))
    return F.relu(bn(z @ p['linear2.weight'].t() + p['linear2.bias'], 'bn2'))

if __name__ == '__main__':
    z = np.load('tests/golden/pct_ref.npz')
    p = {k: (v.double() if v.is_floating_point() else v) for k, v in O.random_params(int(z['param_seed'])).items()}
    x = torch.from_numpy(z['x']).double()
    ref = O.naive_pct(x, p)
    ri = lambda a, b: float((a - b).abs().max() / b.abs().max())
    print('emul(no quant) vs oracle', ri(run(x, p, {}), ref))
    for name, q in [('all2', dict(emb=2, kv=2, energy=2, pv=2, trans=2, lin=2, store=1)), ('energy2', dict(energy=2)), ('energy3', dict(energy=3)),
                    ('energy_f32', dict(energy='f32')), ('pv2', dict(pv=2)), ('kv2', dict(kv=2)), ('trans2', dict(trans=2)), ('lin2', dict(lin=2)), ('emb2', dict(emb=2)),
                    ('store', dict(store=1)), ('all2_energy3', dict(emb=2, kv=2, energy=3, pv=2, trans=2, lin=2, store=1)),
                    ('all3', dict(emb=3, kv=3, energy=3, pv=3, trans=3, lin=3, store=1)),
                    ('kv3_energy3', dict(emb=2, kv=3, energy=3, pv=2, trans=2, lin=2, store=1))]:
        print('%-14s %.2e' % (name, ri(run(x, p, q), ref)))
    print('--- fp16 split'); FMT[0]=torch.float16
    for name, q in [('all2', dict(emb=2, kv=2, energy=2, pv=2, trans=2, lin=2, store=1)), ('energy2', dict(energy=2)), ('kv2', dict(kv=2)), ('emb2', dict(emb=2)), ('pv2', dict(pv=2)), ('lin2',dict(lin=2))]:
        print('%-22s %.2e' % (name, ri(run(x, p, q), ref)))
    FMT[0]=torch.bfloat16
    print('--- second set')
    for name, q in [('emb3_kv3_en3', dict(emb=3, kv=3, energy=3, pv=2, trans=2, lin=2, store=1)),
                    ('emb3_kv3_en3_tr3', dict(emb=3, kv=3, energy=3, pv=2, trans=3, lin=2, store=1)),
                    ('emb3_kv3_en3_tr3_pv3', dict(emb=3, kv=3, energy=3, pv=3, trans=3, lin=2, store=1)),
                    ('emb3_kv3_en2', dict(emb=3, kv=3, energy=2, pv=2, trans=2, lin=2, store=1)),
                    ('emb3_kv3_enf32', dict(emb=3, kv=3, energy='f32', pv=2, trans=2, lin=2, store=1)),
                    ('allf32', dict(emb='f32', kv='f32', energy='f32', trans='f32', lin='f32', store=1)),
                    ]:
        print('%-22s %.2e' % (name, ri(run(x, p, q), ref)))
