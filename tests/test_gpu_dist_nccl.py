"""GPU, world_size 2 over NCCL (needs two devices: ``gpurun --gpus 2``; skipped on a single-GPU box): the
data-parallel training path of SURVEY.md 8(e) / C4 end to end on CUDA -- contiguous pair sharding, this repo's
forward / loss / backward kernels on each rank, ONE ``ncclAllReduce`` over the flat gradient buffer -- against the
CPU oracle: rank-0's reduced gradient equals the MEAN of the per-shard oracle gradients (each rank's loss is the
reference loss of its local batch; the reference loss is not separable over pairs), and after the Adam step both
ranks hold bit-identical parameters."""
import os
import tempfile

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.util import ROOT, grad_close, rel_inf

pytestmark = pytest.mark.gpu

MODULES = ['point', 'gat', 'rel', 'attr']


def _batch():
    from sgaligner_b200 import synthetic
    return synthetic.make_batch([9, 12, 7, 10], [11, 8, 10, 9], [5, 6, 4, 6], [4, 3, 4, 5], n_points=256, edge_mode='complete', seed=21)


def _worker(rank, world, init_file, out_file):
    import sys
    sys.path.insert(0, ROOT)
    from sgaligner_b200 import synthetic, to_cuda
    from sgaligner_b200.losses import CustomMultiLossLayer, OverallLoss
    from sgaligner_b200.sg_aligner import MultiModalEncoder
    from sgaligner_b200.trainer import FlatAdam
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', init_method=f'file://{init_file}', rank=rank, world_size=world, device_id=dev)
    torch.manual_seed(0)
    model = MultiModalEncoder(modules=MODULES, rel_dim=41, attr_dim=164).to(dev).train()
    li, lc = CustomMultiLossLayer(4).to(dev), CustomMultiLossLayer(4).to(dev)
    fn = OverallLoss(li, lc, dev, {'zoom': 0.1, 'wt_align_loss': 1.0, 'wt_contrastive_loss': 1.0, 'modules': MODULES})
    opt = FlatAdam(list(model.parameters()) + list(li.parameters()) + list(lc.parameters()), lr=1e-3, weight_decay=1e-6)
    shard = to_cuda(synthetic.shard_batch(_batch(), rank, world), dev)
    opt.zero_grad()
    ld = fn(model(shard), shard)
    ld['loss'].backward()
    w = opt.allreduce_grads()
    assert w == world
    reduced = (opt.flat_grad / world).cpu()
    opt.step(grad_scale=1.0 / world)
    torch.cuda.synchronize()
    flat_after = opt.flat_param.detach().clone()
    gathered = [torch.empty_like(flat_after) for _ in range(world)]
    dist.all_gather(gathered, flat_after)
    same = all(torch.equal(gathered[0], g) for g in gathered)
    if rank == 0:
        names = [k for k, p in model.named_parameters() if p.requires_grad]
        torch.save({'flat': reduced, 'offsets': opt.offsets, 'names': names + ['__lv_ial', '__lv_icl'],
                    'loss': float(ld['loss'].detach()), 'replicas_identical': same}, out_file)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_nccl_gradients_equal_mean_of_shard_oracles():
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 CUDA devices (gpurun --gpus 2); the same plumbing runs over gloo in tests/test_dist_gloo.py')
    from tests.test_dist_gloo import _oracle_grads
    from sgaligner_b200 import synthetic
    from sgaligner_b200.sg_aligner import MultiModalEncoder
    with tempfile.TemporaryDirectory() as td:
        init_file, out_file = os.path.join(td, 'init'), os.path.join(td, 'out.pt')
        mp.spawn(_worker, args=(2, init_file, out_file), nprocs=2, join=True)
        res = torch.load(out_file)
    assert res['replicas_identical']
    torch.manual_seed(0)
    model = MultiModalEncoder(modules=MODULES, rel_dim=41, attr_dim=164)
    params = {k: v.detach().clone() for k, v in model.state_dict().items()}
    runs = [_oracle_grads(params, synthetic.shard_batch(_batch(), r, 2)) for r in range(2)]
    scale = 0.0
    wants = {}
    for k in res['names']:
        if k == '__lv_ial':
            wants[k] = (runs[0][1].grad + runs[1][1].grad) / 2
        elif k == '__lv_icl':
            wants[k] = (runs[0][2].grad + runs[1][2].grad) / 2
        else:
            g0, g1 = runs[0][0][k].grad, runs[1][0][k].grad
            wants[k] = None if g0 is None else (g0 + g1) / 2
        if wants[k] is not None:
            scale = max(scale, float(wants[k].abs().max()))
    worst = ('', 0.0)
    for k, off in zip(res['names'], res['offsets']):
        want = wants[k]
        if want is None:            # BatchNorm affine parameters: no gradient in the reference, zeros here
            n = dict(model.named_parameters())[k].numel()
            assert float(res['flat'][off:off + n].abs().max()) == 0.0, k
            continue
        got = res['flat'][off:off + want.numel()].view_as(want)
        worst = max(worst, (k, rel_inf(got, want)), key=lambda t: t[1])
        assert grad_close(got, want, rtol=1e-3, atol=1e-6 * scale), (k, rel_inf(got, want))
    print(f'2-rank NCCL: worst gradient error vs mean of shard oracles {worst[1]:.2e} ({worst[0]})')
