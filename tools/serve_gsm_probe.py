"""Device-resident serving throughput at C2 against the number of SMs left to the graph branch (tools only)."""
import os, sys
sys.path.insert(0, os.getcwd())
import torch
from sgaligner_b200 import synthetic, to_cuda
from sgaligner_b200.serving import PipelinedServing
from sgaligner_b200.sg_aligner import MultiModalEncoder

dev = torch.device('cuda:0')
torch.manual_seed(0)
model = MultiModalEncoder(modules=['point', 'gat'], rel_dim=41, attr_dim=164).to(dev).eval()
data = to_cuda(dict(synthetic.config_c2(batch=32, seed=100)), dev)
STEPS, SLOTS = 60, 7
for gsm in (16, 12, 10, 8, 6, 4, 2, 0):
    pipe = PipelinedServing(model, data, k=6, n_slots=SLOTS, compute_streams=3, graph_branch_sms=gsm)
    for s in range(SLOTS):
        pipe.load_resident(s, data)
    def loop(n):
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record(); pipe.fork_resident()
        for k in range(n):
            pipe.submit_resident(k % SLOTS)
        pipe.sync_resident(); t1.record(); torch.cuda.synchronize()
        return t0.elapsed_time(t1) / n
    loop(SLOTS)
    ms = min(loop(STEPS) for _ in range(3))
    print('graph-branch SMs %2d (point CTAs %3d): %.4f ms/step  %.0f pairs/s' % (gsm, pipe.slots[0].pointnet_ctas, ms, 32 / ms * 1e3), flush=True)
    del pipe
    torch.cuda.empty_cache()
