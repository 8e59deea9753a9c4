"""Graph-replay time of the serving step vs the number of SMs left to the graph branch (tools only)."""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from sgaligner_b200 import synthetic, to_cuda
from sgaligner_b200.serving import CapturedInference
from sgaligner_b200.sg_aligner import MultiModalEncoder
dev = torch.device('cuda:0')
torch.manual_seed(0)
model = MultiModalEncoder(modules=['point', 'gat'], rel_dim=41, attr_dim=164).to(dev).eval()
data = to_cuda(synthetic.config_c2(batch=32, seed=100), dev)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
ref = None
for sms in [0, 4, 8, 12, 16, 24]:
    cap = CapturedInference(model, data, k=6, graph_branch_sms=sms)
    for _ in range(5): cap.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(30):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); out = cap.replay(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    if ref is None:
        ref = out['topk_idx'].clone()
    same = bool(torch.equal(ref, out['topk_idx']))
    print(f'graph_branch_sms={sms:3d} pointnet_ctas={cap.pointnet_ctas:4d} median {np.median(ts):.4f} ms  min {min(ts):.4f}  same_topk={same}')
    del cap
