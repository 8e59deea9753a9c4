"""One training step of the C2 workload, for ncu (tools only)."""
import os, sys
sys.path.insert(0, os.getcwd())
import torch
from sgaligner_b200 import synthetic, to_cuda
from sgaligner_b200.losses import CustomMultiLossLayer, OverallLoss
from sgaligner_b200.sg_aligner import MultiModalEncoder
from sgaligner_b200.trainer import FlatAdam, train_step
dev = torch.device('cuda:0')
mods = ['point', 'gat']
torch.manual_seed(0)
model = MultiModalEncoder(modules=mods, rel_dim=41, attr_dim=164).to(dev)
li, lc = CustomMultiLossLayer(2).to(dev), CustomMultiLossLayer(2).to(dev)
fn = OverallLoss(li, lc, dev, {'zoom': 0.1, 'wt_align_loss': 1.0, 'wt_contrastive_loss': 1.0, 'modules': mods})
data = to_cuda(synthetic.config_c2(batch=32, seed=100), dev)
opt = FlatAdam(list(model.parameters()) + list(li.parameters()) + list(lc.parameters()), lr=1e-3, weight_decay=1e-6)
model.train()
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    train_step(model, fn, opt, data)
torch.cuda.synchronize()
