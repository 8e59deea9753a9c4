"""B200-native mirror of ``src/aligner/sg_aligner.py`` of the reference.

Same class names, constructor signatures, ``forward(data_dict)`` dict contract and ``state_dict``
keys, so ``from aligner.sg_aligner import *`` in the reference's trainers / testers can resolve to
this module.  The arithmetic runs in the CUDA kernels of ``libsga_b200.so``;
the ``nn.Linear`` / ``nn.Conv1d`` / ``nn.BatchNorm1d`` objects below are parameter containers that
reproduce the reference's key names, shapes and initialisation -- their own ``forward`` is never
called.
"""
from __future__ import annotations

import math
from contextlib import nullcontext as _nullcontext

import os

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F   # noqa: F401  (re-exported: the reference scripts rely on `import *`)

from . import autograd as ag
from . import ops

__all__ = ['torch', 'nn', 'F', 'ProjectionHead', 'MultiModalFusion', 'MultiModalEncoder', 'PointNetfeat', 'MultiGAT', 'GATConv']


class ProjectionHead(nn.Module):
    """``sg_aligner.py:9-21`` (not used by the trainer; kept for API completeness).  Small MLP on
    top of an embedding -- plain PyTorch, it is not on the hot path."""

    def __init__(self, in_dim, hidden_dim, out_dim, dropout):
        super().__init__()
        self.l1 = nn.Linear(in_dim, hidden_dim, bias=False)
        self.l2 = nn.Linear(hidden_dim, out_dim, bias=False)
        self.dropout = dropout

    def forward(self, x):
        return self.l2(F.dropout(F.relu(self.l1(x)), self.dropout, training=self.training))


class MultiModalFusion(nn.Module):
    """``sg_aligner.py:23-35``: holds the modality weights; the fused kernel applies them."""

    def __init__(self, modal_num, with_weight=1):
        super().__init__()
        self.modal_num = modal_num
        self.requires_grad = True if with_weight > 0 else False
        self.weight = nn.Parameter(torch.ones((self.modal_num, 1)), requires_grad=self.requires_grad)

    def forward(self, embs):
        # stand-alone use on already-projected embeddings: identity projections are not worth a
        # kernel of their own -- route through the fused node with explicit identity weights
        assert len(embs) == self.modal_num
        args = []
        for e in embs:
            d = e.shape[1]
            args += [e, torch.eye(d, device=e.device, dtype=torch.float32), torch.zeros(d, device=e.device)]
        return ag.ProjectFuse.apply(self.weight, self.modal_num, *args)[-1]


class PointNetfeat(nn.Module):
    """Parameter container + launcher for ``networks/pointnet.py:87-175`` in the configuration the
    aligner uses (no STN, global feature, batch_norm=True whose outputs are discarded)."""

    def __init__(self, global_feat=True, input_transform=False, feature_transform=False, point_size=3, out_size=1024,
                 batch_norm=True, init_weights=True, pointnet_str=None):
        super().__init__()
        if input_transform or feature_transform or not global_feat or point_size != 3:
            raise NotImplementedError('only the configuration used by MultiModalEncoder is implemented')
        self.name = 'pnetenc'
        self.use_batch_norm = batch_norm
        self.point_size = point_size
        self.out_size = out_size
        self.conv1 = nn.Conv1d(point_size, 64, 1)
        self.conv2 = nn.Conv1d(64, 128, 1)
        self.conv3 = nn.Conv1d(128, out_size, 1)
        if batch_norm:
            self.bn1 = nn.BatchNorm1d(64)
            self.bn2 = nn.BatchNorm1d(128)
            self.bn3 = nn.BatchNorm1d(out_size)
        self.track_bn_stats = True
        self.kernel_mode = ops.POINTNET_TC if out_size % 128 == 0 else ops.POINTNET_SIMT
        # how train() obtains the BatchNorm batch statistics on the tensor-core path: 'fused' = summed in the forward
        # kernel's epilogues (one launch), 'gram' = from Gram matrices by a separate kernel (ops.pointnet_bn_moments_gram)
        self.bn_stats_mode = os.environ.get('SGA_BN_STATS', 'gram')
        if init_weights:
            # networks/base.py:5-56 as called at pointnet.py:116-118: xavier_normal(gain 1), zero bias,
            # BatchNorm weight 1 / bias 0
            for conv in (self.conv1, self.conv2, self.conv3):
                nn.init.xavier_normal_(conv.weight.data, gain=1)
                nn.init.constant_(conv.bias.data, 0.0)

    def forward(self, pts_npc: torch.Tensor, chunks=None, cta_cap: int = 0) -> torch.Tensor:
        """``pts_npc``: [N, P, 3] (the collated layout; the reference permutes to [N,3,P] first).
        ``chunks``: readiness events of a streamed host-to-device copy (``data.to_cuda_streamed``).
        ``cta_cap``: SMs this encoder's backward may occupy (0 = all); the caller caps the forward itself."""
        want_stats = self.use_batch_norm and self.training and self.track_bn_stats
        gram = want_stats and self.kernel_mode == ops.POINTNET_TC and self.bn_stats_mode == 'gram' and self.conv3.weight.shape[1] == 128
        fused = want_stats and self.kernel_mode == ops.POINTNET_TC and not gram
        if gram:
            # statistics from tensor-core Gram matrices of the ReLU outputs (a third of a forward) + the plain forward,
            # instead of summing every conv output in the forward's epilogues
            if chunks:
                for (_, _, ev) in chunks:
                    torch.cuda.current_stream().wait_event(ev)
            with torch.no_grad():
                mom = ops.pointnet_bn_moments_gram(pts_npc, self.conv1.weight, self.conv1.bias, self.conv2.weight, self.conv2.bias,
                                                   self.conv3.weight, self.conv3.bias)
            self._update_bn_running_stats(mom, float(pts_npc.shape[0] * pts_npc.shape[1]))
        if want_stats and not fused and not gram:   # fp32 FMA kernel (pt_out_dim not a multiple of 128): separate statistics pass
            if chunks:
                for (_, _, ev) in chunks:
                    torch.cuda.current_stream().wait_event(ev)
            mom = ops.pointnet_bn_moments(pts_npc, self.conv1.weight, self.conv1.bias, self.conv2.weight, self.conv2.bias,
                                          self.conv3.weight, self.conv3.bias)
            self._update_bn_running_stats(mom, float(pts_npc.shape[0] * pts_npc.shape[1]))
        out, mom = ag.PointNetFeat.apply(pts_npc, self.conv1.weight, self.conv1.bias, self.conv2.weight, self.conv2.bias,
                                         self.conv3.weight, self.conv3.bias, self.kernel_mode, chunks, fused, torch.is_grad_enabled(), cta_cap)
        if fused:                           # statistics came out of the forward launch itself
            self._update_bn_running_stats(mom, float(pts_npc.shape[0] * pts_npc.shape[1]))
        return out

    @torch.no_grad()
    def _update_bn_running_stats(self, mom, cnt):
        """Train-mode side effect of the discarded BatchNorm calls (pointnet.py:141-142,154-155,
        158-159): running_mean/var <- momentum update with the batch statistics of the pre-ReLU
        conv outputs; outputs are unaffected.  ``mom``: f64 {sum1,sq1,sum2,sq2,sum3,sq3}."""
        bns = (self.bn1, self.bn2, self.bn3)
        if mom.is_cuda and all(b.running_mean.is_cuda and b.running_mean.dtype == torch.float32 and b.momentum is not None for b in bns) \
                and len({b.momentum for b in bns}) == 1:
            ops.bn_running_update(mom, cnt, bns)        # one launch instead of ~30 element-wise kernels
            return
        o = 0
        for bn, c in ((self.bn1, 64), (self.bn2, 128), (self.bn3, self.out_size)):
            s, sq = mom[o:o + c], mom[o + c:o + 2 * c]
            o += 2 * c
            mean = s / cnt
            var_unbiased = (sq - s * mean) / max(cnt - 1.0, 1.0)
            m = bn.momentum
            bn.running_mean.mul_(1 - m).add_(mean.float(), alpha=m)
            bn.running_var.mul_(1 - m).add_(var_unbiased.float(), alpha=m)
            bn.num_batches_tracked += 1


class _SharedLinear(nn.Module):
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels))


class GATConv(nn.Module):
    """Parameter container with torch_geometric 2.2.0 ``GATConv`` names: ``lin_src`` / ``lin_dst``
    (one shared module), ``att_src``, ``att_dst`` [1,H,C], ``bias`` [H*C]; glorot init."""

    def __init__(self, in_channels, out_channels, heads=1):
        super().__init__()
        self.in_channels, self.out_channels, self.heads = in_channels, out_channels, heads
        self.lin_src = _SharedLinear(in_channels, heads * out_channels)
        self.lin_dst = self.lin_src
        self.att_src = nn.Parameter(torch.empty(1, heads, out_channels))
        self.att_dst = nn.Parameter(torch.empty(1, heads, out_channels))
        self.bias = nn.Parameter(torch.zeros(heads * out_channels))
        for t in (self.lin_src.weight, self.att_src, self.att_dst):
            a = math.sqrt(6.0 / (t.size(-2) + t.size(-1)))
            with torch.no_grad():
                t.uniform_(-a, a)


class MultiGAT(nn.Module):
    """``networks/gat.py:27-48`` over a whole batch of graphs at once."""

    def __init__(self, n_units=(17, 128, 100), n_heads=(2, 2), dropout=0.0):
        super().__init__()
        self.num_layers = len(n_units) - 1
        self.dropout = dropout
        if dropout != 0.0:
            raise NotImplementedError('dropout > 0 is not used by the reference configs')
        layers = []
        for i in range(self.num_layers):
            in_channels = n_units[i] * n_heads[i - 1] if i else n_units[i]
            layers.append(GATConv(in_channels=in_channels, out_channels=n_units[i + 1], heads=n_heads[i]))
        self.layer_stack = nn.ModuleList(layers)

    def forward(self, x: torch.Tensor, graph: 'ops.BatchGraph') -> torch.Tensor:
        for idx, layer in enumerate(self.layer_stack):
            x = ag.GATLayer.apply(x, layer.lin_src.weight, layer.att_src, layer.att_dst, layer.bias, graph, layer.heads,
                                  layer.out_channels, idx + 1 < self.num_layers)
        return x


_TRAIN_SIDE = {}      # device -> side stream of the training step's graph branch (not a module attribute: modules get deep-copied / pickled)


class MultiModalEncoder(nn.Module):
    """``sg_aligner.py:37-137``."""

    def __init__(self, modules, rel_dim, attr_dim, hidden_units=[3, 128, 128], heads=[2, 2], emb_dim=100, pt_out_dim=256,
                 dropout=0.0, attn_dropout=0.0, instance_norm=False):
        super().__init__()
        self.modules = modules            # a plain list attribute, exactly as in the reference (:42)
        self.pt_out_dim = pt_out_dim
        self.rel_dim = rel_dim
        self.emb_dim = emb_dim
        self.attr_dim = attr_dim
        self.hidden_units = hidden_units
        self.heads = heads
        self.dropout = dropout
        self.attn_dropout = attn_dropout
        self.instance_norm = instance_norm
        self.inner_view_num = len(self.modules)

        self.meta_embedding_rel = nn.Linear(self.rel_dim, self.emb_dim)
        self.meta_embedding_attr = nn.Linear(self.attr_dim, self.emb_dim)
        if 'point' in self.modules:
            self.object_encoder = PointNetfeat(global_feat=True, batch_norm=True, point_size=3, input_transform=False,
                                               feature_transform=False, out_size=self.pt_out_dim)
        elif 'pct' in self.modules:
            from .pct import NaivePCT
            self.object_encoder = NaivePCT()
        else:
            raise NotImplementedError
        self.object_embedding = nn.Linear(self.pt_out_dim, self.emb_dim)
        self.structure_encoder = MultiGAT(n_units=self.hidden_units, n_heads=self.heads, dropout=self.dropout)
        self.structure_embedding = nn.Linear(256, self.emb_dim)
        self.fusion = MultiModalFusion(modal_num=self.inner_view_num, with_weight=1)
        self.train_overlap = os.environ.get('SGA_TRAIN_OVERLAP', '1') != '0'

    def forward(self, data_dict):
        pts = data_dict['tot_obj_pts']
        if not (torch.is_tensor(pts) and pts.is_cuda):
            raise RuntimeError('sgaligner_b200.MultiModalEncoder needs the batch on a CUDA device (no CPU fallback)')
        ready = data_dict.get('_sga_ready')          # streamed H2D copy in flight (data.to_cuda_streamed)
        if ready is not None:
            torch.cuda.current_stream().wait_event(ready['small'])
        # Launch order.  Streamed batch (H2D copy still in flight): the graph branch only needs the small
        # tensors, so it goes first and overlaps the point copy.  Resident batch: the point encoder goes
        # first -- it is one long launch, and the host-side work of the graph branch (CSR offsets, four
        # short launches) is then enqueued while it runs instead of leaving the GPU idle.
        point_x = None
        # serving.CapturedInference: the graph branch on a forked stream, concurrent with the point encoder; the
        # fork has to precede the point-encoder launch, or the side stream would wait for it
        side = data_dict.get('_sga_side_stream') if ('point' in self.modules and 'gat' in self.modules and ready is None) else None
        cur = torch.cuda.current_stream()
        # Training: the same fork for the whole step.  The graph branch runs on a side stream, and because autograd replays
        # every node on the stream of its forward, so does its backward -- next to the point encoder's backward, whose
        # persistent kernel leaves it 16 SMs.
        cta_cap = 0
        if (side is None and self.train_overlap and self.training and torch.is_grad_enabled() and ready is None
                and 'point' in self.modules and 'gat' in self.modules and not torch.cuda.is_current_stream_capturing()):
            side = _TRAIN_SIDE.get(pts.device)
            if side is None:
                side = _TRAIN_SIDE[pts.device] = torch.cuda.Stream(device=pts.device)
            cta_cap = max(1, ops.sm_count() - 16)
            ops.note_side_stream(side)
        if side is not None:
            side.wait_stream(cur)
        if 'point' in self.modules and ready is None:
            # the cap is handed to the BACKWARD only: in the forward the point encoder is 10 % faster on all SMs than on
            # 132, which is more than hiding the 0.1 ms graph branch gains; the 0.3 ms backward of the graph branch is
            # worth the 16 SMs (measured: tools/train_stages.py)
            point_x = self.object_encoder(pts, data_dict.get('_sga_point_ranges'), cta_cap)
        elif 'pct' in self.modules:
            if ready is not None:
                for (_, _, ev) in ready['pts']:
                    cur.wait_event(ev)
            point_x = self.object_encoder(pts)
        gat_out = None
        if 'gat' in self.modules:
            with torch.cuda.stream(side) if side is not None else _nullcontext():
                graph = data_dict.get('_sga_graph')
                if graph is None:
                    graph = ops.BatchGraph(data_dict['edges'], np.asarray(data_dict['graph_per_obj_count']),
                                           np.asarray(data_dict['graph_per_edge_count']), layout=data_dict.get('_sga_graph_layout'))
                gat_out = self.structure_encoder(data_dict['tot_rel_pose'], graph)
            if side is not None:
                cur.wait_stream(side)
        args = []
        for module in self.modules:
            if module == 'gat':
                args += [gat_out, self.structure_embedding.weight, self.structure_embedding.bias]
            elif module == 'point':
                x = point_x if point_x is not None else self.object_encoder(pts, None if ready is None else ready['pts'])
                args += [x, self.object_embedding.weight, self.object_embedding.bias]
            elif module == 'pct':
                args += [point_x, self.object_embedding.weight, self.object_embedding.bias]
            elif module == 'rel':
                args += [data_dict['tot_bow_vec_object_edge_feats'], self.meta_embedding_rel.weight, self.meta_embedding_rel.bias]
            elif module == 'attr':
                args += [data_dict['tot_bow_vec_object_attr_feats'], self.meta_embedding_attr.weight, self.meta_embedding_attr.bias]
            else:
                raise NotImplementedError
        outs = ag.ProjectFuse.apply(self.fusion.weight, len(self.modules), *args)
        embs = {module: outs[i] for i, module in enumerate(self.modules)}
        if len(self.modules) > 1:
            embs['joint'] = outs[len(self.modules)]
        return embs
