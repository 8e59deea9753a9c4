"""Device-resident serving throughput at C2 for combinations of compute streams, point-encoder launches per step and
SMs left to the graph branch (tools only).   python tools/serve_overlap_probe.py"""
import os, sys, itertools
sys.path.insert(0, os.getcwd())
import torch
from sgaligner_b200 import synthetic, to_cuda
from sgaligner_b200.serving import PipelinedServing
from sgaligner_b200.sg_aligner import MultiModalEncoder

dev = torch.device('cuda:0')
torch.manual_seed(0)
mods = (sys.argv[1] if len(sys.argv) > 1 else 'point,gat').split(',')
model = MultiModalEncoder(modules=mods, rel_dim=41, attr_dim=164).to(dev).eval()
host = synthetic.config_c2(batch=32, seed=100)
data = to_cuda(dict(host), dev)
STEPS, SLOTS = 40, 7

def run(streams, chunks, gsm):
    pipe = PipelinedServing(model, data, k=6, n_slots=SLOTS, compute_streams=streams, point_chunks=chunks, graph_branch_sms=gsm)
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
    best = min(loop(STEPS) for _ in range(3))
    del pipe
    torch.cuda.empty_cache()
    return best

for streams, chunks, gsm in itertools.product((1, 2, 3), (1, 2, 4, 8), (16, 0)):
    if gsm == 0 and chunks == 1 and streams == 1:
        pass
    ms = run(streams, chunks, gsm)
    print('streams %d  point launches %d  graph-branch SMs %2d :  %.4f ms/step  %.0f pairs/s' % (streams, chunks, gsm, ms, 32 / ms * 1e3), flush=True)
