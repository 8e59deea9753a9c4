"""Training-step plumbing around the hot path: flat parameter / gradient buffers, the fused Adam
kernel and the single gradient all-reduce of the data-parallel path.

Mirrors what the reference does around ``model(data_dict)`` / ``loss_func(...)``:
``src/trainers/trainval_sgaligner.py:47-54,71-74`` (Adam(lr 1e-3, wd 1e-6) over the encoder and both
``CustomMultiLossLayer``s; ``train_step`` = forward + loss) and
``src/engine/epoch_based_trainer.py:88-100`` (zero_grad, backward, optimizer step).  The reference's
DDP scaffold is dead code (``base_trainer.py:70``); the B200 path shards the pairs of a batch
across ranks (``synthetic.shard_batch``) and sums ONE flat fp32 gradient buffer (~730 KB) with a
single NCCL all-reduce per step -- no bucketing, the collective is latency-bound.
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch
import torch.distributed as dist

from . import ops

_ALIGN = 64   # floats: keeps every parameter 256-byte aligned inside the flat buffer


class FlatAdam:
    """torch.optim.Adam semantics (L2 weight decay folded into the gradient, bias correction) over
    one flat buffer: parameters and their ``.grad`` become views into ``flat_param`` /
    ``flat_grad``, so the optimiser is one kernel launch and the data-parallel reduction one
    collective."""

    def __init__(self, params: Iterable[torch.nn.Parameter], lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        self.params: List[torch.nn.Parameter] = []
        seen = set()
        for p in params:
            if p.requires_grad and id(p) not in seen:
                seen.add(id(p))
                self.params.append(p)
        assert self.params, 'no trainable parameters'
        dev = self.params[0].device
        self.offsets = []
        total = 0
        for p in self.params:
            assert p.dtype == torch.float32 and p.device == dev
            self.offsets.append(total)
            total += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
        self.numel = total
        self.flat_param = torch.zeros(total, device=dev, dtype=torch.float32)
        self.flat_grad = torch.zeros(total, device=dev, dtype=torch.float32)
        self.exp_avg = torch.zeros(total, device=dev, dtype=torch.float32)
        self.exp_avg_sq = torch.zeros(total, device=dev, dtype=torch.float32)
        with torch.no_grad():
            for p, o in zip(self.params, self.offsets):
                view = self.flat_param[o:o + p.numel()].view_as(p)
                view.copy_(p.data)
                p.data = view
                p.grad = self.flat_grad[o:o + p.numel()].view_as(p)
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.step_count = 0

    def zero_grad(self):
        self.flat_grad.zero_()
        for p, o in zip(self.params, self.offsets):   # re-attach in case something replaced .grad
            if p.grad is None or p.grad.data_ptr() != self.flat_grad.data_ptr() + 4 * o:
                p.grad = self.flat_grad[o:o + p.numel()].view_as(p)

    def allreduce_grads(self, group=None):
        """The one collective of the data-parallel path: sum the flat gradient over all ranks."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM, group=group)
            return dist.get_world_size(group)
        return 1

    def step(self, grad_scale: float = 1.0):
        self.step_count += 1
        ops.adam_step(self.flat_param, self.flat_grad, self.exp_avg, self.exp_avg_sq, self.lr, self.betas[0], self.betas[1],
                      self.eps, self.weight_decay, self.step_count, grad_scale)

    def state_dict(self):
        return {'step': self.step_count, 'exp_avg': self.exp_avg.clone(), 'exp_avg_sq': self.exp_avg_sq.clone()}

    def load_state_dict(self, sd):
        self.step_count = int(sd['step'])
        self.exp_avg.copy_(sd['exp_avg'])
        self.exp_avg_sq.copy_(sd['exp_avg_sq'])


def train_step(model, loss_fn, optimizer: FlatAdam, data_dict: dict, group=None) -> dict:
    """One optimisation step on this rank's shard of pairs: forward, loss, backward, one gradient
    all-reduce (if distributed), Adam.  Returns the loss dict of ``OverallLoss.forward``."""
    optimizer.zero_grad()
    out = model(data_dict)
    ld = loss_fn(out, data_dict)
    ld['loss'].backward()
    world = optimizer.allreduce_grads(group)
    optimizer.step(grad_scale=1.0 / world)
    return ld


def infer_step(model, data_dict: dict, k: int = 6) -> dict:
    """One serving step: encoder forward + matching head (top-k per node) + anchor positions
    (Hits@k / MRR inputs), all on the device."""
    from . import matching
    with torch.no_grad():
        out = model(data_dict)
        mods = model.modules
        emb = out['joint'] if len(mods) > 1 else out[mods[0]]
        res = matching.match_batch(emb, data_dict, k=k, full_rank=False)
    res['embeddings'] = out
    return res
