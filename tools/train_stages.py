"""Stage timing of one training step (CUDA events), C2 workload or all four modalities (tools only).
    python tools/train_stages.py [modules=point,gat] [batch=32]"""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
import torch
from sgaligner_b200 import synthetic, to_cuda, ops
from sgaligner_b200.losses import CustomMultiLossLayer, OverallLoss
from sgaligner_b200.sg_aligner import MultiModalEncoder
from sgaligner_b200.trainer import FlatAdam

mods = (sys.argv[1] if len(sys.argv) > 1 else 'point,gat').split(',')
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 32
dev = torch.device('cuda:0')
torch.manual_seed(0)
model = MultiModalEncoder(modules=mods, rel_dim=41, attr_dim=164).to(dev)
M = len(mods)
li, lc = CustomMultiLossLayer(M).to(dev), CustomMultiLossLayer(M).to(dev)
fn = OverallLoss(li, lc, dev, {'zoom': 0.1, 'wt_align_loss': 1.0, 'wt_contrastive_loss': 1.0, 'modules': mods})
data = to_cuda(synthetic.config_c2(batch=batch, seed=100), dev)
opt = FlatAdam(list(model.parameters()) + list(li.parameters()) + list(lc.parameters()), lr=1e-3, weight_decay=1e-6)
model.train()


def ev():
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


rows = []
for it in range(12):
    opt.zero_grad()
    t0 = ev()
    out = model(data)
    t1 = ev()
    ld = fn(out, data)
    t2 = ev()
    ld['loss'].backward()
    t3 = ev()
    opt.allreduce_grads() if hasattr(opt, 'allreduce_grads') and False else None
    opt.step()
    t4 = ev()
    torch.cuda.synchronize()
    rows.append([t0.elapsed_time(t1), t1.elapsed_time(t2), t2.elapsed_time(t3), t3.elapsed_time(t4), t0.elapsed_time(t4)])
r = np.median(np.array(rows[4:]), axis=0)
print('modules=%s batch=%d  forward %.3f  loss(fwd+grad) %.3f  backward %.3f  adam %.3f  total %.3f ms' % (mods, batch, *r))
# the loss alone, back to back (GPU time without Python gaps)
embs = [out[m].detach() for m in mods] + ([out['joint'].detach()] if M > 1 else [])
from sgaligner_b200.losses import _index_tensors
idx = _index_tensors(data, dev)
lv = torch.zeros(M, device=dev)
for want_grad in (False, True):
    for _ in range(3):
        ops.loss_forward_backward(embs, idx, lv, lv, 0.1, want_grad)
    a = ev()
    for _ in range(10):
        ops.loss_forward_backward(embs, idx, lv, lv, 0.1, want_grad)
    b = ev()
    torch.cuda.synchronize()
    print('  loss kernel sequence alone, want_grad=%s: %.3f ms' % (want_grad, a.elapsed_time(b) / 10))
