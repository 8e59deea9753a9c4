"""The loss kernel sequence alone on random embeddings (for ncu).  python tools/loss_only.py [batch=128] [M=4] [reps=2]"""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
import torch
from sgaligner_b200 import ops
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 128
M = int(sys.argv[2]) if len(sys.argv) > 2 else 4
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dev = torch.device('cuda:0')
n = 64
N = batch * 2 * n
g = torch.Generator(device='cpu').manual_seed(0)
embs = [torch.randn(N, 100, generator=g).to(dev) for _ in range(M)]
embs.append(torch.cat([torch.nn.functional.normalize(e, dim=1) / M for e in embs], dim=1).contiguous())
e1i, e2i, e1j, e2j = [], [], [], []
for b in range(batch):
    o = b * 2 * n
    e1i += list(range(o, o + 32)); e2i += list(range(o + n, o + n + 32))
    e1j += list(range(o + 32, o + n)); e2j += list(range(o + n + 32, o + 2 * n))
idx = [torch.tensor(x, dtype=torch.int32, device=dev) for x in (e1i, e2i, e1j, e2j)]
lv = torch.zeros(M, device=dev)
for want_grad in (False, True):
    for _ in range(reps):
        ops.loss_forward_backward(embs, idx, lv, lv, 0.1, want_grad)
    a = torch.cuda.Event(enable_timing=True); b_ = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        out = ops.loss_forward_backward(embs, idx, lv, lv, 0.1, want_grad)
    b_.record()
    torch.cuda.synchronize()
    print('batch %d M %d want_grad=%s: %.3f ms  losses %s' % (batch, M, want_grad, a.elapsed_time(b_) / reps, out[0].tolist()))
