"""CPU: the oracle restatement is pinned against the golden vectors generated from the UNMODIFIED
reference (oracle/make_golden.py, run in the build container where /root/reference exists)."""
import numpy as np
import pytest
import torch

from oracle import sgaligner_oracle as O
from tests.util import CASES, grad_close, load_case, rel_inf


@pytest.mark.parametrize('name', CASES)
def test_oracle_embeddings(name):
    c = load_case(name)
    out = O.encoder_forward(c['params'], c['data'], c['modules'])
    assert set(out) == set(c['out'])
    for k in out:
        assert rel_inf(out[k], c['out'][k]) < 5e-6, k


@pytest.mark.parametrize('name', CASES)
def test_oracle_losses(name):
    c = load_case(name)
    ld = O.overall_loss(c['out'], c['data'], c['modules'], c['lv'][0], c['lv'][1])
    for k, ref in c['loss'].items():
        assert abs(float(ld[k]) - ref) <= 5e-5 * max(abs(ref), 1e-9), k


@pytest.mark.parametrize('name', ['small4', 'point_only'])
def test_oracle_gradients(name):
    """autograd through the restatement reproduces the reference's parameter gradients."""
    c = load_case(name)
    p = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and 'running' not in k else v) for k, v in c['params'].items()}
    w = p['structure_encoder.layer_stack.0.lin_src.weight']
    p['structure_encoder.layer_stack.0.lin_dst.weight'] = w
    p['structure_encoder.layer_stack.1.lin_dst.weight'] = p['structure_encoder.layer_stack.1.lin_src.weight']
    lvi, lvc = c['lv'][0].clone().requires_grad_(True), c['lv'][1].clone().requires_grad_(True)
    out = O.encoder_forward(p, c['data'], c['modules'])
    O.overall_loss(out, c['data'], c['modules'], lvi, lvc)['loss'].backward()
    for k, g in c['grad'].items():
        if k.startswith('__'):
            got = (lvi if k == '__lv_ial' else lvc).grad
            got = torch.zeros_like(g) if got is None else got
        else:
            got = p[k].grad
        assert got is not None, k
        assert grad_close(got, g), (k, rel_inf(got, g))


@pytest.mark.parametrize('name', CASES)
def test_oracle_matching(name):
    c = load_case(name)
    key = 'joint' if len(c['modules']) > 1 else c['modules'][0]
    ev = O.evaluate_batch(c['out'][key], c['data'])
    assert [ev['hits'][k] for k in range(1, 6)] == c['hits'].tolist()
    assert abs(ev['mrr'] - float(np.mean(c['rr']))) < 1e-12
    for b, r in enumerate(ev['rank']):
        srt = np.take_along_axis(c['sim'][b], c['rank'][b], 1)
        gaps = np.diff(srt, axis=1) > 1e-6
        ok = np.ones_like(r, dtype=bool)
        ok[:, 1:] &= gaps
        ok[:, :-1] &= gaps
        assert (r[ok] == c['rank'][b][ok]).all()


def test_bn_batch_stats_match_reference_side_effect():
    """running_mean/var after one train-mode forward of the reference = momentum update with the
    batch statistics of the pre-ReLU conv outputs (the BN outputs themselves are discarded)."""
    c = load_case('small4')
    stats = O.pointnet_bn_batch_stats(c['data']['tot_obj_pts'], c['params'])
    for i, (mean, var) in enumerate(stats, 1):
        rm = 0.9 * c['params'][f'object_encoder.bn{i}.running_mean'] + 0.1 * mean
        rv = 0.9 * c['params'][f'object_encoder.bn{i}.running_var'] + 0.1 * var
        assert rel_inf(rm, c['bn'][f'object_encoder.bn{i}.running_mean']) < 1e-5
        assert rel_inf(rv, c['bn'][f'object_encoder.bn{i}.running_var']) < 1e-5


@pytest.mark.parametrize('name', CASES)
def test_oracle_pair_metrics_vs_reference(name):
    """SGAR / alignment score / top-1 node correspondences of the restatement against the values the
    reference's own ``utils/alignment.py`` functions produced (tests/golden/metrics_ref.npz, written by
    oracle/make_golden_metrics.py)."""
    import os
    from tests.util import GOLD
    ref = np.load(os.path.join(GOLD, 'metrics_ref.npz'))
    c = load_case(name)
    key = 'joint' if len(c['modules']) > 1 else c['modules'][0]
    ev = O.evaluate_batch(c['out'][key], c['data'])
    si = 0
    for b in range(c['data']['batch_size']):
        assert abs(ev['alignment_score'][b] - float(ref[f'{name}/{b}/alignment_score'])) < 1e-12
        assert ev['node_corrs'][b] == [tuple(int(v) for v in r) for r in ref[f'{name}/{b}/node_corrs']]
        if int(c['data']['e1i_count'][b]):
            got = [ev['sgar'][m][si] for m in ('2', '50', '100')]
            assert got == ref[f'{name}/{b}/sgar'].tolist()
            si += 1
