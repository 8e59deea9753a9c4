"""One forward + backward attention call of the NaivePCT SA layer (for ncu captures and timing).
python tools/pct_attn_bench.py [N] [P] [reps]"""
import sys

import torch

sys.path.insert(0, '.')
from sgaligner_b200 import ops  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 592
P = int(sys.argv[2]) if len(sys.argv) > 2 else 512
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dev = torch.device('cuda:0')
g = torch.Generator().manual_seed(0)
k = (torch.randn(N, P, 32, generator=g) * 1.5).to(dev)
v = torch.randn(N, P, 128, generator=g).to(dev)
dxs = (torch.randn(N, P, 128, generator=g) * 1e-3).to(dev)
for _ in range(reps):
    xs, c2 = ops.pct_attention(k, v, want_c2=True)
    dk1, dk2, dv, dv_colsum, dv_absmax = ops.pct_attention_backward(k, v, c2, dxs)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ops.KERNEL_EVENTS = []
xs, c2 = ops.pct_attention(k, v, want_c2=True)
dk1, dk2, dv, dv_colsum, dv_absmax = ops.pct_attention_backward(k, v, c2, dxs)
torch.cuda.synchronize()
for nme, e0, e1 in ops.KERNEL_EVENTS:
    print('%-20s %.3f ms' % (nme, e0.elapsed_time(e1)))
