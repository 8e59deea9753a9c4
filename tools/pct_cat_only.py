import sys, math, torch
sys.path.insert(0, '.')
from sgaligner_b200 import ops
dev = torch.device('cuda:0')
N, P = 592, 512
g = torch.Generator().manual_seed(0)
xs = [torch.randn(N, P, 128, generator=g).to(dev) for _ in range(4)]
a4, b4 = torch.rand(128, generator=g).to(dev), torch.randn(128, generator=g).to(dev)
WL = (torch.randn(1024, 512, generator=g) / math.sqrt(512)).to(dev)
for _ in range(3):
    ops.pct_cat_linear(xs[0], xs[1], xs[2], xs[3], (a4, b4), WL, track=True)
    ops.pct_cat_linear(xs[0], xs[1], xs[2], xs[3], (a4, b4), WL, track=False)
torch.cuda.synchronize()
