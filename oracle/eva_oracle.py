"""CPU restatement of the EVA baseline path (``src/aligner/eva.py``, ``MultiGCN`` of
``networks/gat.py:6-25``, ``NCALoss`` / ``OverallNCALoss`` of ``losses.py:154-205``).  TEST INFRASTRUCTURE ONLY --
groundwork for a later round; nothing in ``sgaligner_b200/`` imports it.

``GCNConv`` lives in the un-vendored ``torch_geometric==2.2.0`` (``req.yml:259``): restated here from its published
algorithm (``nn/conv/gcn_conv.py``: ``gcn_norm`` with ``add_remaining_self_loops``, symmetric normalisation by the
degree over TARGET nodes, linear map without bias, sum aggregation, bias afterwards) -- parity unpinned at that
boundary, exactly as for ``GATConv``.  ``NCALoss`` and the structure of ``EVA.forward`` are pinned against the
unmodified reference by ``oracle/make_golden_eva.py`` (``tests/golden/eva_ref.npz``).
"""
from __future__ import annotations

from typing import Dict, Sequence

import numpy as np
import torch
import torch.nn.functional as F

from oracle import sgaligner_oracle as O

Tensor = torch.Tensor


def gcn_conv(x: Tensor, edge_index: Tensor, weight: Tensor, bias: Tensor) -> Tensor:
    """PyG 2.2.0 ``GCNConv(in, out, cached=False)`` forward (defaults: ``improved=False, add_self_loops=True,
    normalize=True, bias=True``).  ``edge_index`` [2, e], row 0 = source j, row 1 = target i; self loops in the input
    are replaced by exactly one per node (``add_remaining_self_loops``), duplicate edges count separately;
    ``deg_i = #edges into i`` (self loop included); ``out_i = sum_{j->i} (x_j W^T) / sqrt(deg_j deg_i) + b``."""
    n = x.shape[0]
    ei = edge_index.long()
    keep = ei[0] != ei[1]
    loops = torch.arange(n, dtype=torch.long)
    row = torch.cat([ei[0][keep], loops])
    col = torch.cat([ei[1][keep], loops])
    w = torch.ones(row.shape[0], dtype=x.dtype)
    deg = torch.zeros(n, dtype=x.dtype).scatter_add_(0, col, w)
    dis = deg.pow(-0.5)
    dis[torch.isinf(dis)] = 0
    norm = dis[row] * w * dis[col]
    xw = x @ weight.t()
    out = torch.zeros(n, weight.shape[0], dtype=x.dtype).index_add_(0, col, norm[:, None] * xw[row])
    return out + bias


def multi_gcn(x: Tensor, edge_index: Tensor, p: Dict[str, Tensor], prefix: str = 'structure_encoder') -> Tensor:
    """``MultiGCN.forward`` (gat.py:17-25): GCNConv layers with ReLU (+ dropout p = 0) between them, none after the last."""
    n_layers = len([k for k in p if k.startswith(prefix + '.layer_stack.') and k.endswith('.lin.weight')])
    for i in range(n_layers):
        x = gcn_conv(x, edge_index, p[f'{prefix}.layer_stack.{i}.lin.weight'], p[f'{prefix}.layer_stack.{i}.bias'])
        if i + 1 < n_layers:
            x = F.relu(x)
    return x


def eva_forward(p: Dict[str, Tensor], data: dict, modules: Sequence[str]) -> Dict[str, Tensor]:
    """``EVA.forward`` (eva.py:33-109): 'gcn' per graph over ``tot_rel_pose`` (the raw GCN output IS the embedding, 400-d,
    eva.py:72), 'point' = PointNetfeat(out 200) directly (no projection), 'rel' / 'attr' = Linear(->100); fusion as in
    the aligner."""
    embs: Dict[str, Tensor] = {}
    oc = np.asarray(data['graph_per_obj_count']).reshape(-1, 2)
    ec = np.asarray(data['graph_per_edge_count']).reshape(-1, 2)
    pose = data['tot_rel_pose'].float()
    for module in modules:
        if module == 'gcn':
            outs, o, e = [], 0, 0
            for b in range(oc.shape[0]):
                for gi in range(2):
                    n, ne = int(oc[b, gi]), int(ec[b, gi])
                    edges = data['edges'][e:e + ne].t().to(torch.int32)
                    outs.append(multi_gcn(pose[o:o + n], edges, p))
                    o += n
                    e += ne
            embs[module] = torch.cat(outs)
        elif module == 'point':
            embs[module] = O.pointnet_feat(data['tot_obj_pts'], p)
        elif module == 'rel':
            embs[module] = data['tot_bow_vec_object_edge_feats'].float() @ p['meta_embedding_rel.weight'].t() + p['meta_embedding_rel.bias']
        elif module == 'attr':
            embs[module] = data['tot_bow_vec_object_attr_feats'].float() @ p['meta_embedding_attr.weight'].t() + p['meta_embedding_attr.bias']
        else:
            raise NotImplementedError(module)
    if len(modules) > 1:
        embs['joint'] = O.fusion([embs[m] for m in modules], p['fusion.weight'])
    return embs


def nca_loss(src: Tensor, ref: Tensor, alpha: float = 1.0, beta: float = 1.0, ep: float = 0.0) -> Tensor:
    """``NCALoss.forward`` (losses.py:161-176)."""
    n = src.shape[0]
    scores = src @ ref.t()
    eye = torch.eye(n, dtype=src.dtype)
    s_diag = eye * scores
    s_ = torch.exp(alpha * (scores - ep))
    s_ = s_ - s_ * eye
    loss_diag = -torch.log(1 + F.relu(s_diag.sum(0)))
    return (torch.log(1 + s_.sum(0)) / alpha).mean() + (torch.log(1 + s_.sum(1)) / alpha).mean() + (beta * loss_diag).mean()


def overall_nca_loss(out: Dict[str, Tensor], data: dict) -> Dict[str, Tensor]:
    """``OverallNCALoss.forward`` (losses.py:189-205): one NCA term per entry of the output dict (joint included), summed."""
    e1 = torch.as_tensor(np.asarray(data['e1i']).astype(np.int64))
    e2 = torch.as_tensor(np.asarray(data['e2i']).astype(np.int64))
    losses = {}
    for k, emb in out.items():
        en = F.normalize(emb)
        losses[k] = nca_loss(en[e1], en[e2])
    losses['loss'] = sum(losses.values())
    return losses
