"""Golden vectors for the EVA baseline from the UNMODIFIED reference ``EVA`` module and
``OverallNCALoss`` (GCNConv through the restated stub of ``oracle/ref_import.py``), and the pinning of
``oracle/eva_oracle.py`` against them.  Writes ``tests/golden/eva_ref.npz``.  TEST INFRASTRUCTURE ONLY.

    python -m oracle.make_golden_eva
"""
from __future__ import annotations

import importlib
import os

import numpy as np
import torch

from oracle import eva_oracle, ref_import
from sgaligner_b200 import synthetic

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')
MODULES = ['gcn', 'point', 'rel', 'attr']


def make_data():
    return synthetic.make_batch([6, 9, 4], [7, 5, 8], [3, 4, 2], n_points=64, edge_mode='kout', k_out=3, seed=21)


def main():
    _sg, ls, _al = ref_import.load_reference()
    eva = importlib.import_module('aligner.eva')
    data = make_data()
    torch.manual_seed(5)
    m = eva.EVA(modules=MODULES, rel_dim=41, attr_dim=164)
    with torch.no_grad():
        m.fusion.weight.copy_(1 + 0.5 * torch.randn(len(MODULES), 1))
        for n_, p_ in m.named_parameters():
            if n_.endswith('bias'):
                p_.copy_(0.1 * torch.randn_like(p_))
    m.eval()
    params = {k: v.detach().clone() for k, v in m.state_dict().items()}
    out = m(data)
    loss_fn = ls.OverallNCALoss(MODULES, 'cpu')
    ld = loss_fn(out, data)
    ld['loss'].backward()
    grads = {n_: p_.grad.detach().clone() for n_, p_ in m.named_parameters() if p_.grad is not None}
    # ---- pin the restatement
    p = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and 'running' not in k else v) for k, v in params.items()}
    o_out = eva_oracle.eva_forward(p, data, MODULES)
    o_ld = eva_oracle.overall_nca_loss(o_out, data)
    o_ld['loss'].backward()
    for k in out:
        e = float((o_out[k] - out[k]).abs().max() / out[k].abs().max())
        assert e < 2e-6, (k, e)
    for k in ld:
        assert abs(float(o_ld[k]) - float(ld[k])) <= 2e-6 * abs(float(ld[k])), k
    for k, g in grads.items():
        og = p[k].grad
        assert og is not None and float((og - g).abs().max()) <= 2e-5 * float(g.abs().max() + 1e-12), k
    blob = {'cfg/modules': np.array(MODULES)}
    for k, v in params.items():
        blob['p/' + k] = v.numpy()
    for k, v in out.items():
        blob['out/' + k] = v.detach().numpy()
    for k, v in ld.items():
        blob['loss/' + k] = np.array(float(v))
    for k, v in grads.items():                                   # a subset keeps the fixture small
        if k in ('structure_encoder.layer_stack.0.lin.weight', 'structure_encoder.layer_stack.1.bias', 'object_encoder.conv1.weight',
                 'meta_embedding_rel.weight', 'fusion.weight'):
            blob['grad/' + k] = v.numpy()
    np.savez_compressed(os.path.join(GOLD, 'eva_ref.npz'), **blob)
    print({k: tuple(v.shape) for k, v in out.items()}, {k: round(float(v), 5) for k, v in ld.items()})
    print('wrote eva_ref.npz', os.path.getsize(os.path.join(GOLD, 'eva_ref.npz')) // 1024, 'KiB')


if __name__ == '__main__':
    main()
