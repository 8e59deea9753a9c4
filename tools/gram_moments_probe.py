import os, sys
sys.path.insert(0, os.getcwd())
import torch
from sgaligner_b200 import ops
dev = torch.device('cuda:0')
g = torch.Generator().manual_seed(0)
pts = torch.randn(4096, 512, 3, generator=g).to(dev)
args = [t.to(dev) for t in (torch.randn(64, 3, generator=g), torch.randn(64, generator=g), torch.randn(128, 64, generator=g) * 0.1,
                            torch.randn(128, generator=g), torch.randn(256, 128, generator=g) * 0.1, torch.randn(256, generator=g))]
for _ in range(3):
    ops.pointnet_bn_moments_gram(pts, *args)
torch.cuda.synchronize()
a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    ops.pointnet_bn_moments_gram(pts, *args)
b.record(); torch.cuda.synchronize()
print('gram moments: %.3f ms per call' % (a.elapsed_time(b) / 5))
