"""Per-kernel time of one NaivePCT training step (forward + backward of the encoder alone) at the C2 object count,
from torch.profiler's CUDA activity records (no nsys in the image).  python tools/pct_train_profile.py [N] [P]"""
import sys

import torch

sys.path.insert(0, '.')
from sgaligner_b200.pct import NaivePCT  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
P = int(sys.argv[2]) if len(sys.argv) > 2 else 512
dev = torch.device('cuda:0')
torch.manual_seed(0)
m = NaivePCT().to(dev).train()
x = torch.randn(N, P, 3, device=dev)
R = torch.randn(N, 256, device=dev)


def step():
    for p in m.parameters():
        p.grad = None
    (m(x) * R).sum().backward()


for _ in range(2):
    step()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(3):
    step()
b.record()
torch.cuda.synchronize()
print('encoder fwd+bwd: %.2f ms per step (N=%d P=%d), peak memory %.1f GiB' % (a.elapsed_time(b) / 3, N, P, torch.cuda.max_memory_allocated() / 2 ** 30))
with torch.no_grad():
    m(x)
torch.cuda.synchronize()
a.record()
with torch.no_grad():
    for _ in range(3):
        m(x)
b.record()
torch.cuda.synchronize()
print('encoder fwd (train mode, no grad): %.2f ms' % (a.elapsed_time(b) / 3))
from torch.profiler import ProfilerActivity, profile  # noqa: E402
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
rows = [(e.key, e.device_time_total if hasattr(e, 'device_time_total') else e.cuda_time_total, e.count) for e in prof.key_averages()]
rows = [r for r in rows if r[1] > 0]
rows.sort(key=lambda r: -r[1])
tot = sum(r[1] for r in rows)
print('%-90s %10s %6s' % ('kernel', 'ms', 'calls'))
for k, t, c in rows[:40]:
    print('%-90s %10.3f %6d' % (k[:90], t / 1e3, c))
print('total device time %.2f ms' % (tot / 1e3))
