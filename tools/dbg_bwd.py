import os, sys
sys.path.insert(0, os.getcwd())
import torch
from oracle import sgaligner_oracle as O
from sgaligner_b200 import ops
dev = torch.device('cuda:0')
def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))
for (N, P, C3, seed) in [(7, 64, 128, 0), (7, 64, 256, 0), (74, 64, 128, 0), (75, 64, 128, 0), (301, 130, 128, 1), (301, 130, 256, 1), (600, 512, 256, 2)]:
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
    # fp64 autograd reference through the same argmax
    wd = [t.double().detach().requires_grad_(True) for t in w]
    x = pts.double()
    h1 = torch.relu(x @ wd[0].squeeze(-1).t() + wd[1])
    h2 = torch.relu(h1 @ wd[2].squeeze(-1).t() + wd[3])
    z3 = h2 @ wd[4].squeeze(-1).t() + wd[5]
    sel = torch.gather(z3, 1, arg.long().unsqueeze(1)).squeeze(1)
    o = torch.relu(sel)
    (o * gout.double()).sum().backward()
    ref = [t.grad.reshape(s.shape) for t, s in zip(wd, b)]
    ref = [ref[0], ref[1], ref[2], ref[3], ref[4], ref[5]]
    torch.cuda.synchronize()
    print((N, P, C3), 'tc-vs-simt', ' '.join(f'{rel(x_, y_):.1e}' for x_, y_ in zip(a, b)),
          '| tc-vs-f64', ' '.join(f'{rel(x_, y_):.1e}' for x_, y_ in zip(a, ref)),
          '| simt-vs-f64', ' '.join(f'{rel(x_, y_):.1e}' for x_, y_ in zip(b, ref)))
