"""Timing of the three shapes of sga_pct_wgrad a NaivePCT training step launches (R = N * P rows):
   A + [128, 32] (d W_v | d W_k), A + [128] (d W_t), and the four launches of the upper-triangular 512 x 512 Gram."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sgaligner_b200 import ops

N, P = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (4096, 512)
dev = torch.device('cuda:0')
xs = [torch.randn(N * P, 128, device=dev) for _ in range(4)]
small = torch.randn(N * P, 32, device=dev)


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


C1, C2 = torch.zeros(128, 128, device=dev), torch.zeros(32, 128, device=dev)
G = torch.zeros(512, 512, device=dev)
print('A + [128, 32]: %.3f ms' % timed(lambda: ops.pct_wgrad(xs[0], [xs[1], small], [C1, C2], transpose=True)))
print('A + [128]    : %.3f ms' % timed(lambda: ops.pct_wgrad(xs[0], [xs[1]], [C1])))
for a in range(4):
    print('Gram row %d   : %.3f ms' % (a, timed(lambda: ops.pct_wgrad(xs[a], [xs[b] for b in range(a, 4)],
                                                                      [G[128 * a:128 * a + 128, 128 * b:128 * b + 128] for b in range(a, 4)]))))
