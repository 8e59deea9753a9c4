"""Is the C2 training step GPU bound?  Sum of kernel device time vs the step's wall time (torch.profiler CUDA records)."""
import os, sys
sys.path.insert(0, os.getcwd())
import torch
from torch.profiler import ProfilerActivity, profile
from sgaligner_b200 import synthetic, to_cuda
from sgaligner_b200.losses import CustomMultiLossLayer, OverallLoss
from sgaligner_b200.sg_aligner import MultiModalEncoder
from sgaligner_b200.trainer import FlatAdam, train_step

mods = (sys.argv[1] if len(sys.argv) > 1 else 'point,gat').split(',')
dev = torch.device('cuda:0')
torch.manual_seed(0)
model = MultiModalEncoder(modules=mods, rel_dim=41, attr_dim=164).to(dev).train()
M = len(mods)
li, lc = CustomMultiLossLayer(M).to(dev), CustomMultiLossLayer(M).to(dev)
fn = OverallLoss(li, lc, dev, {'zoom': 0.1, 'wt_align_loss': 1.0, 'wt_contrastive_loss': 1.0, 'modules': mods})
data = to_cuda(synthetic.config_c2(batch=32, seed=100), dev)
opt = FlatAdam(list(model.parameters()) + list(li.parameters()) + list(lc.parameters()), lr=1e-3, weight_decay=1e-6)
for _ in range(10):
    train_step(model, fn, opt, data)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20):
    train_step(model, fn, opt, data)
b.record()
torch.cuda.synchronize()
print('step %.3f ms (20 back to back)' % (a.elapsed_time(b) / 20))
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5):
        train_step(model, fn, opt, data)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type.name == 'CUDA']
tot = sum(e.device_time_total if hasattr(e, 'device_time_total') else e.cuda_time_total for e in ev) / 5 / 1e3
print('sum of kernel device time per step: %.3f ms over %d launches per step (streams overlap, so the sum may exceed the wall time)' % (tot, len(ev) // 5))
rows = {}
for e in ev:
    t = (e.device_time_total if hasattr(e, 'device_time_total') else e.cuda_time_total) / 5 / 1e3
    rows[e.name[:70]] = rows.get(e.name[:70], 0) + t
for k, v in sorted(rows.items(), key=lambda x: -x[1])[:14]:
    print('  %-72s %.3f ms' % (k, v))
