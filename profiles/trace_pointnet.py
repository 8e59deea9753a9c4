"""Pipeline timeline of the tensor-core PointNet kernel (CTA 0): run on the GPU box.
    python profiles/trace_pointnet.py > gpurun_out/trace.txt"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sgaligner_b200 import ops, _lib
from sgaligner_b200.sg_aligner import PointNetfeat
dev = torch.device('cuda:0')
lib = _lib.get_lib()
lib.sga_debug_set_trace.argtypes = [ctypes.c_void_p]
torch.manual_seed(0)
net = PointNetfeat(out_size=256).to(dev)
pts = torch.randn(4096, 512, 3, device=dev)
w = [net.conv1.weight, net.conv1.bias, net.conv2.weight, net.conv2.bias, net.conv3.weight, net.conv3.bias]
with torch.no_grad():
    for _ in range(3):
        ops.pointnet_forward(pts, *w, want_argmax=False)
    trace = torch.zeros(2048, dtype=torch.int64, device=dev)
    lib.sga_debug_set_trace(ctypes.c_void_p(trace.data_ptr()))
    ops.pointnet_forward(pts, *w, want_argmax=False)
    torch.cuda.synchronize()
    lib.sga_debug_set_trace(ctypes.c_void_p(0))
t = trace.cpu().numpy()
c, m = t[:1024].reshape(64, 16), t[1024:].reshape(64, 16)
base = c[8, 0]
names_c = ['d2_full', 'D2inregs', 'converted', 'S1next', 'H2stored', '-', '-', 'd3_full', 'E3done']
names_m = ['MMA2beg', 'MMA2iss', 'h2c0', 'h2c1', 'h2c2', 'h2c3', 'MMA3iss']
for g in range(8, 16):
    print(f'tile {g}: compute ' + ' '.join(f'{n}={c[g, k] - base}' for k, n in enumerate(names_c) if n != '-'))
    print(f'          mma     ' + ' '.join(f'{n}={m[g, k] - base}' for k, n in enumerate(names_m)))
per = np.diff(c[8:60, 0])
print('period (cycles/tile): mean', per.mean(), 'min', per.min(), 'max', per.max())
