"""CPU, world_size 2 over gloo: the data-parallel plumbing of the training path -- contiguous pair
sharding + ONE flat-gradient all-reduce -- reproduces the mean of the per-shard gradients
(rank-0 gradients equal the mean of G independent oracle runs)."""
import os
import tempfile

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.util import ROOT

MODULES = ['point', 'gat', 'rel', 'attr']


def _oracle_grads(params, data):
    from oracle import sgaligner_oracle as O
    p = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and 'running' not in k else v) for k, v in params.items()}
    for i in (0, 1):
        p[f'structure_encoder.layer_stack.{i}.lin_dst.weight'] = p[f'structure_encoder.layer_stack.{i}.lin_src.weight']
    lvi = torch.zeros(4, requires_grad=True)
    lvc = torch.zeros(4, requires_grad=True)
    out = O.encoder_forward(p, data, MODULES)
    O.overall_loss(out, data, MODULES, lvi, lvc)['loss'].backward()
    return p, lvi, lvc


def _worker(rank, world, init_file, out_file):
    import sys
    sys.path.insert(0, ROOT)
    from sgaligner_b200 import synthetic
    from sgaligner_b200.losses import CustomMultiLossLayer
    from sgaligner_b200.sg_aligner import MultiModalEncoder
    from sgaligner_b200.trainer import FlatAdam
    dist.init_process_group('gloo', init_method=f'file://{init_file}', rank=rank, world_size=world)
    torch.manual_seed(0)
    model = MultiModalEncoder(modules=MODULES, rel_dim=41, attr_dim=164)          # CPU: parameter container only
    li, lc = CustomMultiLossLayer(4), CustomMultiLossLayer(4)
    params = {k: v.detach().clone() for k, v in model.state_dict().items()}
    data = synthetic.make_batch([6, 5, 7, 4], [5, 6, 4, 6], [3, 3, 3, 2], n_points=32, edge_mode='complete', seed=7)
    shard = synthetic.shard_batch(data, rank, world)
    assert shard['batch_size'] == 2
    opt = FlatAdam(list(model.parameters()) + list(li.parameters()) + list(lc.parameters()))
    opt.zero_grad()
    p, lvi, lvc = _oracle_grads(params, shard)                                   # this rank's local-batch gradient
    named = dict(model.named_parameters())
    with torch.no_grad():
        for k, prm in named.items():
            if p[k].grad is not None:
                prm.grad.copy_(p[k].grad)
        li.log_vars.grad.copy_(lvi.grad)
        lc.log_vars.grad.copy_(lvc.grad)
    w = opt.allreduce_grads()
    assert w == world
    if rank == 0:
        torch.save({'flat': opt.flat_grad.clone() / world, 'offsets': opt.offsets,
                    'names': [k for k, _ in model.named_parameters()] + ['__lv_ial', '__lv_icl']}, out_file)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_flat_gradient_allreduce():
    from sgaligner_b200 import synthetic
    from sgaligner_b200.sg_aligner import MultiModalEncoder
    with tempfile.TemporaryDirectory() as td:
        init_file, out_file = os.path.join(td, 'init'), os.path.join(td, 'out.pt')
        mp.spawn(_worker, args=(2, init_file, out_file), nprocs=2, join=True)
        res = torch.load(out_file)
    torch.manual_seed(0)
    model = MultiModalEncoder(modules=MODULES, rel_dim=41, attr_dim=164)
    params = {k: v.detach().clone() for k, v in model.state_dict().items()}
    data = synthetic.make_batch([6, 5, 7, 4], [5, 6, 4, 6], [3, 3, 3, 2], n_points=32, edge_mode='complete', seed=7)
    runs = [_oracle_grads(params, synthetic.shard_batch(data, r, 2)) for r in range(2)]
    seen = set()
    names = []
    for k, prm in model.named_parameters():
        names.append(k)
    for k, off in zip(res['names'], res['offsets']):
        if k == '__lv_ial':
            want = (runs[0][1].grad + runs[1][1].grad) / 2
        elif k == '__lv_icl':
            want = (runs[0][2].grad + runs[1][2].grad) / 2
        else:
            g0, g1 = runs[0][0][k].grad, runs[1][0][k].grad
            if g0 is None:
                continue
            want = (g0 + g1) / 2
        got = res['flat'][off:off + want.numel()].view_as(want)
        assert torch.allclose(got, want, rtol=1e-5, atol=1e-7), k
