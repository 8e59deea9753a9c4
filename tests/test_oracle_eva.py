"""The EVA-baseline restatement (oracle/eva_oracle.py, groundwork for SURVEY.md 8(f) row 4) against outputs of the
UNMODIFIED reference ``EVA`` module and ``OverallNCALoss`` frozen in tests/golden/eva_ref.npz
(oracle/make_golden_eva.py; GCNConv is a restated PyG 2.2.0 layer there too -- parity unpinned at that boundary)."""
import os

import numpy as np
import torch

from oracle import eva_oracle
from oracle.make_golden_eva import MODULES, make_data
from tests.util import GOLD


def test_eva_oracle_forward_loss_and_gradients():
    z = np.load(os.path.join(GOLD, 'eva_ref.npz'))
    data = make_data()
    p = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith('p/')}
    p = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and 'running' not in k else v) for k, v in p.items()}
    out = eva_oracle.eva_forward(p, data, MODULES)
    assert {k: tuple(v.shape) for k, v in out.items()} == {'gcn': (39, 400), 'point': (39, 200), 'rel': (39, 100), 'attr': (39, 100),
                                                           'joint': (39, 800)}
    for k in out:
        ref = torch.from_numpy(z['out/' + k])
        assert float((out[k].detach() - ref).abs().max() / ref.abs().max()) < 1e-5, k
    ld = eva_oracle.overall_nca_loss(out, data)
    for k in ld:
        assert abs(float(ld[k].detach()) - float(z['loss/' + k])) <= 1e-5 * abs(float(z['loss/' + k])), k
    ld['loss'].backward()
    for k in z.files:
        if k.startswith('grad/'):
            g, r = p[k[5:]].grad, torch.from_numpy(z[k])
            assert float((g - r).abs().max()) <= 1e-4 * float(r.abs().max() + 1e-12), k


def test_gcn_conv_semantics():
    """Self loops in the input are replaced by exactly one per node; duplicate edges count twice; an isolated node
    keeps its own feature (degree 1)."""
    x = torch.eye(3)
    w = torch.eye(3)
    b = torch.zeros(3)
    e = torch.tensor([[0, 0, 1, 1], [1, 1, 1, 0]])          # 0->1 twice, a self loop on 1, 1->0; node 2 isolated
    y = eva_oracle.gcn_conv(x, e, w, b)
    deg = torch.tensor([2.0, 3.0, 1.0])                      # in-degree + the one self loop
    exp = torch.zeros(3, 3)
    exp[0, 0] = 1 / deg[0]; exp[0, 1] = 1 / (deg[0] * deg[1]).sqrt()
    exp[1, 1] = 1 / deg[1]; exp[1, 0] = 2 / (deg[0] * deg[1]).sqrt()
    exp[2, 2] = 1.0
    assert torch.allclose(y, exp, atol=1e-6)
