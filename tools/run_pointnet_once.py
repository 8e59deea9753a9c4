"""A few launches of the tensor-core PointNet kernel at the C2 shape (for ncu; tools only)."""
import os, sys
sys.path.insert(0, os.getcwd())
import torch
from sgaligner_b200 import ops
from sgaligner_b200.sg_aligner import PointNetfeat
dev = torch.device('cuda:0')
torch.manual_seed(0)
net = PointNetfeat(out_size=256).to(dev)
pts = torch.randn(4096, 512, 3, device=dev)
if 'c2' in sys.argv[1:]:
    from sgaligner_b200 import synthetic
    pts = synthetic.config_c2(batch=32, seed=100)['tot_obj_pts'].to(dev)
if 'offset' in sys.argv[1:]:
    pts = pts + (torch.rand(4096, 1, 3, device=dev) * 4 - 2)
w = [net.conv1.weight, net.conv1.bias, net.conv2.weight, net.conv2.bias, net.conv3.weight, net.conv3.bias]
stats = 'stats' in sys.argv[1:]
argmax = 'argmax' in sys.argv[1:]      # the training forward of the default path (arg-max tracking, no moments)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev) if 'flush' in sys.argv[1:] else None
with torch.no_grad():
    for _ in range(3):
        if stats:
            ops.pointnet_forward_stats(pts, *w, want_argmax=True)
        else:
            ops.pointnet_forward(pts, *w, want_argmax=argmax)
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        if stats:
            ops.pointnet_forward_stats(pts, *w, want_argmax=True)
        else:
            ops.pointnet_forward(pts, *w, want_argmax=argmax)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
print('pointnet_fwd', 'stats' if stats else 'plain', 'flushed' if flush is not None else 'warm', ' '.join(sys.argv[1:]), 'ms:', min(ts), sorted(ts)[len(ts) // 2])
