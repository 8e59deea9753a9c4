"""Golden vectors for ``MultiModalEncoder(modules=['pct', 'gat', 'rel', 'attr'])`` -- the module list of the shipped
ground-truth config (``configs/scan3r/scan3r_ground_truth.yaml:5``) -- from the UNMODIFIED reference
(``src/aligner/sg_aligner.py`` with ``NaivePCT`` from ``networks/pct.py``).  TEST INFRASTRUCTURE ONLY.

    python -m oracle.make_golden_pct_encoder

The 1.5 M parameters are not stored: the PRODUCT module tree is constructed under ``torch.manual_seed(seed)`` (it is a
parameter container with the reference's names and shapes), BatchNorm affine parameters / running statistics are
randomised, and its ``state_dict`` is loaded STRICTLY into the reference model (which pins the key set); the test
rebuilds the same parameters from the seed.  Stored: the reference's key list, inputs, eval-mode embeddings, the
matching metrics, and the train-mode running statistics of three BatchNorm layers.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_import, sgaligner_oracle as O          # noqa: E402
from sgaligner_b200 import synthetic                          # noqa: E402

GOLD = os.path.join(ROOT, 'tests', 'golden')
MODULES = ['pct', 'gat', 'rel', 'attr']
SEED = 17


def seeded_product_state_dict(seed: int = SEED):
    """Parameters of the product ``MultiModalEncoder(['pct', ...])`` under a seed, with non-trivial BatchNorm state."""
    from sgaligner_b200.sg_aligner import MultiModalEncoder
    torch.manual_seed(seed)
    m = MultiModalEncoder(modules=MODULES, rel_dim=41, attr_dim=164)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for name, mod in m.named_modules():
            if isinstance(mod, torch.nn.BatchNorm1d):
                mod.weight.copy_(1.0 + 0.2 * torch.randn(mod.num_features, generator=g))
                mod.bias.copy_(0.1 * torch.randn(mod.num_features, generator=g))
                mod.running_mean.copy_(0.1 * torch.randn(mod.num_features, generator=g))
                mod.running_var.copy_(0.5 + torch.rand(mod.num_features, generator=g))
    return m.state_dict()


def batch():
    return synthetic.make_batch([5, 7], [6, 4], [3, 3], n_points=160, edge_mode='complete', seed=31)


def main():
    sg, ls, al = ref_import.load_reference()
    ref = sg.MultiModalEncoder(modules=MODULES, rel_dim=41, attr_dim=164)
    keys = {k: list(v.shape) for k, v in ref.state_dict().items()}
    sd = seeded_product_state_dict()
    res = ref.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    data = batch()
    ref.eval()
    with torch.no_grad():
        out = ref(data)
    ev = O.evaluate_batch(out['joint'], data)
    blob = {'keys': np.array(json.dumps(keys)), 'seed': np.array(SEED)}
    for k, v in out.items():
        blob['out/' + k] = v.numpy()
    blob['metric/hits'] = np.array([ev['hits'][k] for k in range(1, 6)])
    # train mode: BatchNorm side effects (dropout masks are drawn from the CPU generator in the reference's order)
    ref.train()
    torch.manual_seed(5)
    with torch.no_grad():
        out_t = ref(data)
    blob['train/pct'] = out_t['pct'].numpy()
    for k, v in ref.state_dict().items():
        if k.startswith('object_encoder.') and ('running' in k or 'num_batches' in k):
            blob['after/' + k] = v.numpy()
    path = os.path.join(GOLD, 'pct_encoder.npz')
    np.savez_compressed(path, **blob)
    print('wrote', path, os.path.getsize(path) // 1024, 'KiB;', len(keys), 'keys; hits', ev['hits'])


if __name__ == '__main__':
    main()
