"""One serving step + one training step of a configuration, eager launches (for ncu).  python tools/step_once.py [c2|c2all|c3|c2pct]"""
import os, sys
sys.path.insert(0, os.getcwd())
import torch
from sgaligner_b200 import matching, ops, synthetic, to_cuda
from sgaligner_b200.losses import CustomMultiLossLayer, OverallLoss
from sgaligner_b200.sg_aligner import MultiModalEncoder
from sgaligner_b200.trainer import FlatAdam, train_step
cfg = sys.argv[1] if len(sys.argv) > 1 else 'c2'
dev = torch.device('cuda:0')
mods = ['point', 'gat'] if cfg == 'c2' else (['pct', 'gat', 'rel', 'attr'] if cfg == 'c2pct' else ['point', 'gat', 'rel', 'attr'])
host = synthetic.config_c3(batch=128, seed=1) if cfg == 'c3' else synthetic.config_c2(batch=32, seed=100)
data = to_cuda(dict(host), dev)
torch.manual_seed(0)
model = MultiModalEncoder(modules=mods, rel_dim=41, attr_dim=164).to(dev)
M = len(mods)
li, lc = CustomMultiLossLayer(M).to(dev), CustomMultiLossLayer(M).to(dev)
fn = OverallLoss(li, lc, dev, {'zoom': 0.1, 'wt_align_loss': 1.0, 'wt_contrastive_loss': 1.0, 'modules': mods})
opt = FlatAdam(list(model.parameters()) + list(li.parameters()) + list(lc.parameters()), lr=1e-3, weight_decay=1e-6)
e1 = torch.as_tensor(host['e1i']).to(dev); e2 = torch.as_tensor(host['e2i']).to(dev)
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
for _ in range(reps):
    model.eval()
    with torch.no_grad():
        out = model(data)
        res = matching.match_batch(out['joint'], data, k=6, full_rank=False)
        ops.match_anchor_pos(res['sim'], res['layout'], e1, e2)
        matching.evaluate_pairs(out['joint'], host)
    model.train()
    train_step(model, fn, opt, data)
torch.cuda.synchronize()
