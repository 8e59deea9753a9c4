"""B200-native mirror of the EVA baseline (``src/aligner/eva.py:9-96``): same class name,
constructor signature, ``state_dict`` key set and output dict as the reference module.

    'gcn'   MultiGCN([3, 200, 400]) over ``tot_rel_pose`` -- the raw 400-d GCN output IS the embedding (eva.py:72)
    'point' PointNetfeat(out_size=200) directly, no projection (eva.py:75)
    'rel' / 'attr'  Linear(-> emb_dim) of the bag-of-words vectors (eva.py:78-81)
    'joint' MultiModalFusion of the above (eva.py:88-94)

All graphs of the batch go through the GCN in one block-diagonal launch per stage (the reference loops over the 2B
graphs in Python, eva.py:47-70).  ``GCNConv`` is restated from torch_geometric 2.2.0 (un-vendored dependency; parity
unpinned at that boundary, see oracle/eva_oracle.py).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn as nn

from . import autograd as ag
from . import ops
from .sg_aligner import MultiModalFusion, PointNetfeat

__all__ = ['EVA', 'MultiGCN', 'GCNConv']


class _Lin(nn.Module):
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels))


class GCNConv(nn.Module):
    """Parameter container with torch_geometric 2.2.0 ``GCNConv`` names: ``lin.weight`` [out, in] (glorot), ``bias`` [out] (zeros)."""

    def __init__(self, in_channels, out_channels, cached=False):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.bias = nn.Parameter(torch.zeros(out_channels))
        self.lin = _Lin(in_channels, out_channels)
        a = math.sqrt(6.0 / (in_channels + out_channels))
        with torch.no_grad():
            self.lin.weight.uniform_(-a, a)


class MultiGCN(nn.Module):
    """``networks/gat.py:6-25`` over a whole batch of graphs at once."""

    def __init__(self, n_units=(17, 128, 100), dropout=0.0):
        super().__init__()
        self.num_layers = len(n_units) - 1
        self.dropout = dropout
        if dropout != 0.0:
            raise NotImplementedError('dropout > 0 is not used by the reference configs')
        self.layer_stack = nn.ModuleList([GCNConv(n_units[i], n_units[i + 1]) for i in range(self.num_layers)])

    def forward(self, x: torch.Tensor, graph: 'ops.BatchGraph', graph_t: 'ops.BatchGraph') -> torch.Tensor:
        x = ops.as_f32(x)
        for idx, layer in enumerate(self.layer_stack):
            x = ag.GCNLayer.apply(x, layer.lin.weight, layer.bias, graph, graph_t, idx + 1 < self.num_layers)
        return x


class _Affine(torch.autograd.Function):
    """y = x W^T + b for the two meta embeddings (eva.py:78-81): the projection kernel of the aligner without its fusion part."""

    @staticmethod
    def forward(ctx, x, W, b):
        y = ops.project_fuse(x, W, b, None, 0, None, 1, 0)
        ctx.save_for_backward(x, W, y)
        return y

    @staticmethod
    def backward(ctx, g):
        x, W, y = ctx.saved_tensors
        gW, gb, _, _ = ops.project_fuse_backward(x, W, y, g.contiguous(), None, 0, None, 1, 0, False)
        return None, gW, gb


class EVA(nn.Module):
    """``eva.py:9-96``."""

    def __init__(self, modules, rel_dim, attr_dim, n_units=[3, 200, 400], emb_dim=100, pt_out_dim=256, dropout=0.0,
                 attn_dropout=0.0, instance_norm=False):
        super().__init__()
        self.modules = modules
        self.pt_out_dim = pt_out_dim
        self.rel_dim = rel_dim
        self.emb_dim = emb_dim
        self.attr_dim = attr_dim
        self.n_units = n_units
        self.dropout = dropout
        self.attn_dropout = attn_dropout
        self.instance_norm = instance_norm
        self.inner_view_num = len(self.modules)
        self.meta_embedding_rel = nn.Linear(self.rel_dim, self.emb_dim)
        self.meta_embedding_attr = nn.Linear(self.attr_dim, self.emb_dim)
        self.object_encoder = PointNetfeat(global_feat=True, batch_norm=True, point_size=3, input_transform=False,
                                           feature_transform=False, out_size=200)
        self.structure_encoder = MultiGCN(n_units=self.n_units, dropout=self.dropout)
        self.fusion = MultiModalFusion(modal_num=self.inner_view_num, with_weight=1)

    def forward(self, data_dict):
        pts = data_dict['tot_obj_pts']
        if not (torch.is_tensor(pts) and pts.is_cuda):
            raise RuntimeError('sgaligner_b200.EVA needs the batch on a CUDA device (no CPU fallback)')
        embs = {}
        for module in self.modules:
            if module == 'gcn':
                oc, ec = np.asarray(data_dict['graph_per_obj_count']), np.asarray(data_dict['graph_per_edge_count'])
                graph = ops.BatchGraph(data_dict['edges'], oc, ec)
                # the backward needs the transposed (by-source) adjacency: the CSR of the reversed edges, same layout
                graph_t = ops.BatchGraph(data_dict['edges'].flip(1), layout=graph.layout) if torch.is_grad_enabled() else None
                emb = self.structure_encoder(data_dict['tot_rel_pose'], graph, graph_t)
            elif module == 'point':
                emb = self.object_encoder(pts)
            elif module == 'rel':
                emb = _Affine.apply(data_dict['tot_bow_vec_object_edge_feats'], self.meta_embedding_rel.weight, self.meta_embedding_rel.bias)
            elif module == 'attr':
                emb = _Affine.apply(data_dict['tot_bow_vec_object_attr_feats'], self.meta_embedding_attr.weight, self.meta_embedding_attr.bias)
            else:
                raise NotImplementedError
            embs[module] = emb
        if len(self.modules) > 1:
            embs['joint'] = ag.FuseRows.apply(self.fusion.weight, *[embs[m] for m in self.modules])
        return embs
