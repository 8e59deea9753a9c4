"""Golden vectors for the ``NaivePCT`` encoder from the UNMODIFIED reference module (``src/aligner/networks/pct.py``;
``pointnet2_ops`` -- used only by the ``PCT``/``SG`` classes -- is stubbed by ``oracle/ref_import.py``), and the
pinning of ``oracle/pct_oracle.py`` against them.  Writes ``tests/golden/pct_ref.npz``.  TEST INFRASTRUCTURE ONLY.

    python -m oracle.make_golden_pct
"""
from __future__ import annotations

import importlib
import os

import numpy as np
import torch

from oracle import pct_oracle, ref_import

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')


def main():
    ref_import.load_reference()
    pct = importlib.import_module('aligner.networks.pct')
    m = pct.NaivePCT()
    params0 = pct_oracle.random_params(7)
    res = m.load_state_dict(params0, strict=True)              # exact key set / shapes of the reference module
    assert not res.missing_keys and not res.unexpected_keys
    g = torch.Generator().manual_seed(11)
    x = torch.randn(6, 3, 96, generator=g) + torch.rand(6, 3, 1, generator=g) * 2 - 1
    blob = {'x': x.numpy(), 'param_seed': np.array(7)}
    # ---- eval mode
    m.eval()
    with torch.no_grad():
        y_eval = m(x)
    p = {k: v.clone() for k, v in params0.items()}
    with torch.no_grad():
        o_eval = pct_oracle.naive_pct(x, p, training=False)
    err = float((o_eval - y_eval).abs().max() / y_eval.abs().max())
    assert err < 1e-5, err
    blob['y_eval'] = y_eval.numpy()
    # ---- train mode: batch statistics, running-stat side effect, dropout masks from the same seed
    m.train()
    torch.manual_seed(123)
    with torch.no_grad():
        y_train = m(x)
    after = {k: v.detach().clone() for k, v in m.state_dict().items() if 'running' in k or 'num_batches' in k}
    p = {k: v.clone() for k, v in params0.items()}
    torch.manual_seed(123)
    with torch.no_grad():
        o_train = pct_oracle.naive_pct(x, p, training=True)
    err_t = float((o_train - y_train).abs().max() / y_train.abs().max())
    assert err_t < 1e-5, err_t          # fp32 summation order (einsum vs conv1d) only
    for k, v in after.items():
        if 'running' in k:
            assert float((p[k] - v).abs().max()) <= 1e-5 * float(v.abs().max() + 1e-12), k
        blob['after/' + k] = v.numpy()
    blob['y_train'] = y_train.numpy()
    blob['train_seed'] = np.array(123)
    # ---- parameter gradients of the UNMODIFIED reference module in train mode (same seed -> same dropout masks):
    # loss = sum(y * R) for a fixed random R.  Tensors above 20k elements are stored as every 17th element + their norm.
    m.load_state_dict(params0, strict=True)
    m.train()
    R = torch.randn(y_train.shape, generator=torch.Generator().manual_seed(29))
    torch.manual_seed(123)
    y_g = m(x)
    assert torch.equal(y_g.detach(), y_train)
    (y_g * R).sum().backward()
    po = {k: v.clone().requires_grad_(v.is_floating_point() and 'running' not in k) for k, v in params0.items()}
    for sa in ('sa1', 'sa2', 'sa3', 'sa4'):
        po[sa + '.q_conv.weight'] = po[sa + '.k_conv.weight']
    torch.manual_seed(123)
    (pct_oracle.naive_pct(x, po, training=True) * R).sum().backward()
    blob['grad_R'] = R.numpy()
    worst = 0.0
    gmax = max(float(prm.grad.abs().max()) for prm in m.parameters())
    for name, prm in m.named_parameters():
        g = prm.grad
        key = name.replace('q_conv', 'k_conv')
        go = po[key].grad
        # trans_conv.bias / linear2.bias (a bias in front of a train-mode BatchNorm) and, on this input, linear.1.bias have
        # mathematically zero gradients: both sides hold rounding noise there (3e-7 of the largest gradient in fp32) -> floor the denominator
        worst = max(worst, float((go - g).abs().max() / g.abs().max().clamp_min(1e-3 * gmax)))
        flat = g.reshape(-1)
        blob['gnorm/' + key] = np.array(float(flat.double().norm()))
        blob['grad/' + key] = (flat[::17] if flat.numel() > 20000 else flat).numpy()
    blob['grad_max'] = np.array(gmax)
    assert worst < 2e-3, worst           # fp32 autograd of two differently ordered evaluations
    print('oracle-vs-reference gradient mismatch (fp32 both): %.2e' % worst)
    np.savez_compressed(os.path.join(GOLD, 'pct_ref.npz'), **blob)
    print('eval err %.2e  train err %.2e  out %s  params %d  GFLOP/object@512 %.3f' % (
        err, err_t, tuple(y_eval.shape), sum(p_.numel() for p_ in m.parameters()),
        pct_oracle.flops_per_object(512) / 1e9))
    print('wrote pct_ref.npz', os.path.getsize(os.path.join(GOLD, 'pct_ref.npz')) // 1024, 'KiB')


if __name__ == '__main__':
    main()
