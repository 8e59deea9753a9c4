"""Direct check of the loss forward (F values via losses) for the gram_ts path vs the legacy GEMM path and timing."""
import os, sys, subprocess
sys.path.insert(0, os.getcwd())
import torch
from sgaligner_b200 import ops
dev = torch.device('cuda:0')
def run(batch, M, n=64, na=32):
    N = batch * 2 * n
    g = torch.Generator().manual_seed(0)
    embs = [torch.randn(N, 100, generator=g).to(dev) for _ in range(M)]
    embs.append(torch.cat([torch.nn.functional.normalize(e, dim=1) / M for e in embs], dim=1).contiguous())
    e1i, e2i, e1j, e2j = [], [], [], []
    for b in range(batch):
        o = b * 2 * n
        e1i += list(range(o, o + na)); e2i += list(range(o + n, o + n + na))
        e1j += list(range(o + na, o + n)); e2j += list(range(o + n + na, o + 2 * n))
    idx = [torch.tensor(x, dtype=torch.int32, device=dev) for x in (e1i, e2i, e1j, e2j)]
    lv = torch.zeros(M, device=dev)
    res = {}
    for want_grad in (False, True):
        for _ in range(3):
            out = ops.loss_forward_backward(embs, idx, lv, lv, 0.1, want_grad)
        a = torch.cuda.Event(enable_timing=True); b_ = torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            out = ops.loss_forward_backward(embs, idx, lv, lv, 0.1, want_grad)
        b_.record(); torch.cuda.synchronize()
        res[want_grad] = (a.elapsed_time(b_) / 5, out)
    return res
for batch, M in ((32, 2), (32, 4), (128, 4), (5, 4)):
    r = run(batch, M, n=64 if batch != 5 else 37, na=32 if batch != 5 else 11)
    print('%s batch %d M %d: fwd %.3f ms  fwd+bwd %.3f ms  losses %s  |grad|max %.4e' % (os.environ.get('SGA_LOSS_GRAM', 'ts'), batch, M, r[False][0], r[True][0],
          ['%.6f' % v for v in r[True][1][0].tolist()], max(float(g.abs().max()) for g in r[True][1][1])))
    torch.save([r[True][1][0].cpu()] + [g.cpu() for g in r[True][1][1]], '/tmp/gram_%s_%d_%d.pt' % (os.environ.get('SGA_LOSS_GRAM', 'ts'), batch, M))
if os.environ.get('SGA_LOSS_GRAM') is None:
    env = dict(os.environ, SGA_LOSS_GRAM='legacy')
    subprocess.run([sys.executable, __file__], env=env, check=True)
    for batch, M in ((32, 2), (32, 4), (128, 4), (5, 4)):
        a = torch.load('/tmp/gram_ts_%d_%d.pt' % (batch, M)); b = torch.load('/tmp/gram_legacy_%d_%d.pt' % (batch, M))
        print('batch %d M %d  ts vs legacy: loss rel %.2e  grad rel %.2e' % (batch, M, float(((a[0] - b[0]).abs() / b[0].abs().clamp_min(1e-12)).max()),
              max(float((x - y).abs().max() / y.abs().max().clamp_min(1e-30)) for x, y in zip(a[1:], b[1:]))))
