"""Timing of the NaivePCT object encoder at the C2 object count (tools only; CUDA events, per-kernel via ops.KERNEL_EVENTS).
    python tools/pct_bench.py [N=4096] [P=512]"""
import os, sys
sys.path.insert(0, os.getcwd())
import collections
import numpy as np
import torch
from sgaligner_b200 import ops
from sgaligner_b200.pct import NaivePCT

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
P = int(sys.argv[2]) if len(sys.argv) > 2 else 512
dev = torch.device('cuda:0')
torch.manual_seed(0)
m = NaivePCT().to(dev)
pts = torch.randn(N, P, 3, device=dev) + torch.rand(N, 1, 3, device=dev) * 4 - 2
flop = N * (2 * P * (3 * 128 + 128 * 128) + 4 * (2 * P * (128 * 32 + 2 * 128 * 128) + 2 * P * P * 160) + 2 * P * 512 * 1024 + 2 * (1024 * 512 + 512 * 256))
for mode in ('eval', 'train'):
    m.train(mode == 'train')
    with torch.no_grad():
        for _ in range(3):
            y = m(pts)
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); y = m(pts); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ops.KERNEL_EVENTS = []
        y = m(pts)
        torch.cuda.synchronize()
        agg = collections.OrderedDict()
        for name, a, b in ops.KERNEL_EVENTS:
            agg.setdefault(name, []).append(a.elapsed_time(b))
        ops.KERNEL_EVENTS = None
    t = float(np.median(ts))
    print('%s: N=%d P=%d  %.3f ms / forward  = %.1f objects/ms, %.1f TFLOP/s algorithmic (%.2f GFLOP/object)  finite=%s' %
          (mode, N, P, t, N / t, flop / t / 1e9, flop / N / 1e9, bool(torch.isfinite(y).all())))
    for name, v in agg.items():
        print('   %-16s x%-2d  total %.3f ms   (%s)' % (name, len(v), sum(v), ' '.join('%.3f' % x for x in v)))
