"""Debug: attention backward (dv, dk) vs fp64 autograd at growing energy scales, next to fp32 torch autograd."""
import math
import sys

import torch

sys.path.insert(0, '.')
from sgaligner_b200 import ops  # noqa: E402


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max())


dev = torch.device('cuda:0')
for P in (200, 512):
    for ks in (1.2, 3.0, 6.0, 12.0):
        for off in (0.0, 5.0):
            g = torch.Generator().manual_seed(1)
            N = 4
            k = (torch.randn(N, P, 32, generator=g) * ks + off).to(dev)
            v = torch.randn(N, P, 128, generator=g).to(dev)
            dxs = (torch.randn(N, P, 128, generator=g) * 1e-5).to(dev)
            xs, c2 = ops.pct_attention(k, v, want_c2=True)
            dk1, dk2, dv, dv_colsum, dv_absmax = ops.pct_attention_backward(k, v, c2, dxs)
            torch.cuda.synchronize()

            def ref(dt):
                kd = k.to(dt).requires_grad_(True)
                vd = v.to(dt).requires_grad_(True)
                A = torch.softmax(kd @ kd.transpose(1, 2) / math.sqrt(32), dim=-1)
                o = A.transpose(1, 2) @ vd
                (o * dxs.to(dt)).sum().backward()
                return o.detach(), kd.grad, vd.grad
            o64, gk64, gv64 = ref(torch.float64)
            o32, gk32, gv32 = ref(torch.float32)
            print('P=%d kscale=%.1f off=%.0f | fwd %.1e (f32 %.1e) | dv %.1e (f32 %.1e) | dk %.1e (f32 %.1e) | dk_row %.1e dk_col %.1e' % (
                P, ks, off, rel(xs, o64), rel(o32, o64), rel(dv, gv64), rel(gv32, gv64), rel(dk1 + dk2, gk64), rel(gk32, gk64),
                float(dk1.abs().max()), float(dk2.abs().max())))
