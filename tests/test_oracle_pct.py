"""The NaivePCT restatement (oracle/pct_oracle.py, groundwork for SURVEY.md 8(f) row 1) against outputs of the
UNMODIFIED reference module frozen in tests/golden/pct_ref.npz (oracle/make_golden_pct.py): eval mode, and train mode
including the BatchNorm running-statistics side effect and the dropout masks of the recorded seed."""
import os

import numpy as np
import torch

from oracle import pct_oracle
from tests.util import GOLD


def _gold():
    return np.load(os.path.join(GOLD, 'pct_ref.npz'))


def test_pct_oracle_eval():
    z = _gold()
    p = pct_oracle.random_params(int(z['param_seed']))
    with torch.no_grad():
        y = pct_oracle.naive_pct(torch.from_numpy(z['x']), p, training=False)
    ref = torch.from_numpy(z['y_eval'])
    assert y.shape == ref.shape == (6, 256)
    assert float((y - ref).abs().max() / ref.abs().max()) < 1e-5
    assert int(p['bn1.num_batches_tracked']) == 0                       # eval leaves the buffers alone


def test_pct_oracle_train_mode_side_effects():
    z = _gold()
    p = pct_oracle.random_params(int(z['param_seed']))
    torch.manual_seed(int(z['train_seed']))
    with torch.no_grad():
        y = pct_oracle.naive_pct(torch.from_numpy(z['x']), p, training=True)
    ref = torch.from_numpy(z['y_train'])
    assert float((y - ref).abs().max() / ref.abs().max()) < 1e-5
    assert float((y == 0).float().mean()) > 0.4                         # dropout(0.5) after a ReLU
    for k in z.files:
        if k.startswith('after/') and 'running' in k:
            r = torch.from_numpy(z[k])
            assert float((p[k[6:]] - r).abs().max()) <= 1e-5 * float(r.abs().max()), k


def test_pct_flop_count_matches_survey():
    assert abs(pct_oracle.flops_per_object(512) / 1e9 - 1.06) < 0.05     # ~1.06 GFLOP / object
