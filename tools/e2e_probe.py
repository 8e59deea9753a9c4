"""e2e step anatomy: pure H2D time of one batch, and the e2e step for several chunk counts (tools only)."""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from sgaligner_b200 import matching, ops, synthetic, to_cuda
from sgaligner_b200.data import h2d_bytes, pin, to_cuda_streamed
from sgaligner_b200.sg_aligner import MultiModalEncoder
dev = torch.device('cuda:0')
torch.manual_seed(0)
model = MultiModalEncoder(modules=['point', 'gat'], rel_dim=41, attr_dim=164).to(dev).eval()
host = synthetic.config_c2(batch=32, seed=100)
hp = pin(host)
e1 = torch.as_tensor(host['e1i']).pin_memory(); e2 = torch.as_tensor(host['e2i']).pin_memory()
def timed(fn, n=30, w=5):
    for _ in range(w): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts))
def h2d_only():
    d = to_cuda(dict(hp), dev)
    return d
print('bytes', h2d_bytes(host), 'H2D only ms', timed(h2d_only))
for nc in (1, 2, 4, 8, 16):
    def step():
        d = to_cuda_streamed(hp, dev, n_chunks=nc)
        with torch.no_grad():
            out = model(d)
            res = matching.match_batch(out['joint'], d, k=6, full_rank=False)
            pos = ops.match_anchor_pos(res['sim'], res['layout'], e1.to(dev, non_blocking=True), e2.to(dev, non_blocking=True))
        return res['topk_idx'].cpu(), pos.cpu()
    print('chunks', nc, 'e2e ms', timed(step))
