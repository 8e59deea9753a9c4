import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from sgaligner_b200 import synthetic, to_cuda
from sgaligner_b200.serving import CapturedInference
from sgaligner_b200.sg_aligner import MultiModalEncoder
dev = torch.device('cuda:0')
torch.manual_seed(0)
model = MultiModalEncoder(modules=['point', 'gat'], rel_dim=41, attr_dim=164).to(dev).eval()
host = synthetic.config_c2(batch=32, seed=100)
data = to_cuda(dict(host), dev)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
FLUSH = 'flush' in sys.argv[1:]
def timed(fn, n=20, w=5):
    for _ in range(w): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        if FLUSH: flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts))
# pure H2D graph
hp = host['tot_obj_pts'].pin_memory()
dst = torch.empty_like(data['tot_obj_pts'])
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    dst.copy_(hp, non_blocking=True)
torch.cuda.synchronize()
with torch.cuda.graph(g):
    dst.copy_(hp, non_blocking=True)
print('graph H2D 25MB ms', timed(g.replay), ' eager H2D ms', timed(lambda: dst.copy_(hp, non_blocking=True)))
for nc in (4, 8, 4, 8, 2):
    cap = CapturedInference(model, data, k=6)
    cap.capture_host_step(host, n_chunks=nc)
    print('chunks', nc, 'run_host ms', timed(cap.run_host), 'again', timed(cap.run_host))
    del cap
