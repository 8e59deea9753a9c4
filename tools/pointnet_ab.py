"""PointNet tensor-core forward at C2: per-launch time (CUDA events, L2 flushed) for a CTA cap of 132 / none (tools only).
Run under SGA_LIB_PATH=... / SGA_POINTNET_SCHED=static|dynamic for A/B comparisons on one box."""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from sgaligner_b200 import ops, synthetic
from sgaligner_b200.sg_aligner import PointNetfeat
dev = torch.device('cuda:0')
torch.manual_seed(0)
net = PointNetfeat(out_size=256).to(dev)
pts = synthetic.config_c2(batch=32, seed=100)['tot_obj_pts'].to(dev)
w = [net.conv1.weight, net.conv1.bias, net.conv2.weight, net.conv2.bias, net.conv3.weight, net.conv3.bias]
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
for cap in (132, 0):
    ops.pointnet_set_max_ctas(cap)
    for want in (False, True):
        with torch.no_grad():
            for _ in range(3):
                ops.pointnet_forward(pts, *w, want_argmax=want)
            ts = []
            for _ in range(12):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); ops.pointnet_forward(pts, *w, want_argmax=want); b.record(); torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
        print('lib=%s sched=%s cap=%3d argmax=%d : median %.4f ms  min %.4f' % (os.path.basename(os.environ.get('SGA_LIB_PATH', 'current')),
              os.environ.get('SGA_POINTNET_SCHED', 'dynamic'), cap, want, float(np.median(ts)), min(ts)))
ops.pointnet_set_max_ctas(0)
