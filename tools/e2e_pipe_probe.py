"""Host-to-host pipelined serving at C2: steady-state ms/step for combinations of compute streams and steps in flight,
next to the copy-only floor of the same loop (tools only).   python tools/e2e_pipe_probe.py"""
import os, sys
sys.path.insert(0, os.getcwd())
import torch
from sgaligner_b200 import synthetic, to_cuda
from sgaligner_b200.data import pin
from sgaligner_b200.serving import PipelinedServing
from sgaligner_b200.sg_aligner import MultiModalEncoder

dev = torch.device('cuda:0')
torch.manual_seed(0)
model = MultiModalEncoder(modules=['point', 'gat'], rel_dim=41, attr_dim=164).to(dev).eval()
host = synthetic.config_c2(batch=32, seed=100)
data = to_cuda(dict(pin(host)), dev)
SLOTS, STEPS = 7, 42


def loop(pipe, in_flight, steps, copy_only=False):
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for k in range(steps):
        s = k % SLOTS
        if k >= in_flight:
            pipe.wait((k - in_flight) % SLOTS)
        if copy_only:
            c = pipe.slots[s]
            with torch.cuda.stream(pipe.copy_stream):
                for k_ in pipe.keys:
                    c.static[k_].copy_(c.p_in[k_], non_blocking=True)
                c.e1.copy_(c.p_e1, non_blocking=True); c.e2.copy_(c.p_e2, non_blocking=True)
                c.ev_done.record(pipe.copy_stream)
            c.in_flight = True
        else:
            pipe.submit(s)
    for k in range(max(0, steps - in_flight), steps):
        pipe.wait(k % SLOTS)
    torch.cuda.current_stream().wait_stream(pipe.copy_stream)
    t1.record(); torch.cuda.synchronize()
    return t0.elapsed_time(t1) / steps

for streams in (1, 2, 3):
    pipe = PipelinedServing(model, data, k=6, n_slots=SLOTS, compute_streams=streams)
    for s in range(SLOTS):
        pipe.fill(s, host)
    if streams == 1:
        loop(pipe, 3, SLOTS, copy_only=True)
        print('copy-only loop (all H2D bytes of a step, %d B): %.4f ms/step' % (pipe.h2d_bytes, min(loop(pipe, 3, STEPS, True) for _ in range(3))), flush=True)
    for inf in (2, 3, 4, 6):
        loop(pipe, inf, SLOTS)
        ms = min(loop(pipe, inf, STEPS) for _ in range(3))
        print('compute streams %d  in flight %d : %.4f ms/step  %.0f pairs/s' % (streams, inf, ms, 32 / ms * 1e3), flush=True)
    del pipe
    torch.cuda.empty_cache()
