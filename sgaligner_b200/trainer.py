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
        # torch.optim.Optimizer surface the reference's trainer touches: ``param_groups[0]['lr']`` is read by
        # ``base_trainer.get_lr``, scaled by the world size in ``register_optimizer`` and written by lr schedulers
        self.param_groups = [{'params': self.params, 'lr': lr, 'betas': tuple(betas), 'eps': eps, 'weight_decay': weight_decay,
                              'amsgrad': False, 'maximize': False}]
        self.defaults = {k: v for k, v in self.param_groups[0].items() if k != 'params'}
        self.step_count = 0
        # parameter segments of the flat buffer (padding belongs to the preceding parameter; its gradient stays 0)
        self._seg_off = torch.tensor(self.offsets + [total], dtype=torch.int64, device=dev)
        self._seg_active = torch.zeros(len(self.params), dtype=torch.int32, device=dev)
        # the backward kernels may add parameter gradients straight into the flat buffer (explicit opt-in)
        ops.enable_direct_grad(self.params)

    # single-group conveniences (kept for callers that set them directly)
    lr = property(lambda self: self.param_groups[0]['lr'], lambda self, v: self.param_groups[0].__setitem__('lr', v))
    betas = property(lambda self: self.param_groups[0]['betas'])
    eps = property(lambda self: self.param_groups[0]['eps'])
    weight_decay = property(lambda self: self.param_groups[0]['weight_decay'])

    def zero_grad(self, set_to_none: bool = False):
        """One memset of the flat gradient buffer (``set_to_none`` is accepted for signature compatibility and
        ignored: the gradients are views of one allocation)."""
        self.flat_grad.zero_()
        for p, o in zip(self.params, self.offsets):   # re-attach in case something replaced .grad
            if p.grad is None or p.grad.data_ptr() != self.flat_grad.data_ptr() + 4 * o:
                p.grad = self.flat_grad[o:o + p.numel()].view_as(p)

    def allreduce_grads(self, group=None):
        """The one collective of the data-parallel path: sum the flat gradient over all ranks."""
        ops.join_side_streams()
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM, group=group)
            return dist.get_world_size(group)
        return 1

    def step(self, grad_scale: float = 1.0, closure=None):
        """One Adam update of the whole flat buffer.  A parameter whose gradient is exactly zero in this step (torch:
        ``.grad is None`` -- modules that are not selected, the BatchNorm affine parameters whose outputs the
        reference discards) is skipped like torch.optim.Adam skips it: no weight-decay drift, moments untouched.
        Deviation: one global step counter (torch keeps one per parameter; they only differ for a parameter that
        receives gradients in some steps and none in others)."""
        g = self.param_groups[0]
        self.step_count += 1
        ops.join_side_streams()
        ops.adam_step_segments(self.flat_param, self.flat_grad, self.exp_avg, self.exp_avg_sq, self._seg_off, self._seg_active,
                               float(g['lr']), g['betas'][0], g['betas'][1], g['eps'], g['weight_decay'], self.step_count, grad_scale)

    def state_dict(self):
        """``torch.optim.Adam.state_dict()`` layout (``state[i] = {step, exp_avg, exp_avg_sq}`` per parameter index,
        ``param_groups`` with index lists), so the reference's ``save_snapshot`` / ``load_snapshot`` round-trip and a
        checkpoint written by ``torch.optim.Adam`` over the same parameter list loads here."""
        state = {}
        for i, (p, o) in enumerate(zip(self.params, self.offsets)):
            n = p.numel()
            state[i] = {'step': torch.tensor(float(self.step_count)),
                        'exp_avg': self.exp_avg[o:o + n].view_as(p).clone(),
                        'exp_avg_sq': self.exp_avg_sq[o:o + n].view_as(p).clone()}
        groups = [{**{k: v for k, v in self.param_groups[0].items() if k != 'params'}, 'params': list(range(len(self.params)))}]
        return {'state': state, 'param_groups': groups}

    def load_state_dict(self, sd):
        if 'state' not in sd:                      # round-1 flat format
            self.step_count = int(sd['step'])
            self.exp_avg.copy_(sd['exp_avg'])
            self.exp_avg_sq.copy_(sd['exp_avg_sq'])
            return
        steps = []
        for i, (p, o) in enumerate(zip(self.params, self.offsets)):
            st = sd['state'].get(i, sd['state'].get(str(i)))
            if st is None:                          # torch keeps no state for a parameter that never had a gradient
                continue
            n = p.numel()
            self.exp_avg[o:o + n].copy_(st['exp_avg'].reshape(-1))
            self.exp_avg_sq[o:o + n].copy_(st['exp_avg_sq'].reshape(-1))
            steps.append(int(float(st['step'])))
        self.step_count = max(steps) if steps else 0
        for k, v in sd['param_groups'][0].items():
            if k != 'params' and k in self.param_groups[0]:
                self.param_groups[0][k] = tuple(v) if k == 'betas' else v


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
