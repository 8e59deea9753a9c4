"""CPU oracle for the SGAligner node-embedding / matching / contrastive-loss hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``sgaligner_b200/`` may import this module; it is
used by ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` as the checker / reported CPU baseline, never as the product path.

It is a functional restatement (torch CPU ops, dtype chosen by the caller: fp32 to mirror the
reference, fp64 to bound the oracle's own rounding) of the algorithm the reference runs in

* ``src/aligner/sg_aligner.py:23-137``   (MultiModalFusion, MultiModalEncoder.forward)
* ``src/aligner/networks/pointnet.py:120-175`` (PointNetfeat.forward, no STN, global_feat)
* ``src/aligner/networks/gat.py:27-48``  (MultiGAT.forward)
* ``torch_geometric==2.2.0`` ``nn/conv/gat_conv.py`` + ``utils/softmax.py`` + ``utils/loop.py``
  (GATConv; un-vendored third-party dependency pinned in ``req.yml:259`` -- restated from the
  published algorithm; NO reference test pins results at that boundary => "parity unpinned"
  for the GAT branch)
* ``src/aligner/losses.py:5-152``        (calculate_prob_dist, ICL, IAL, multi-loss, OverallLoss)
* ``src/inference/sgaligner/inference_align_reg.py:122-128`` (matching head)
* ``utils/alignment.py:3-89``            (rank metrics)

Pinning: ``oracle/make_golden.py`` imports the real reference modules (with three stub
modules for the missing import-only dependencies) in the build container, checks this
restatement against them on seeded inputs and writes ``tests/golden/*.npz``; the CPU test-suite
re-checks the restatement against those committed vectors.

Parameters are passed as a flat ``dict`` keyed by the reference ``state_dict`` names
(``object_encoder.conv1.weight`` ...), so a reference checkpoint can be fed straight in.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------
# parameter initialisation (mirrors what the reference constructors do, by key)
# --------------------------------------------------------------------------------------
def init_params(modules: Sequence[str], rel_dim: int, attr_dim: int, hidden_units=(3, 128, 128),
                heads=(2, 2), emb_dim: int = 100, pt_out_dim: int = 256, seed: int = 0,
                dtype=torch.float32) -> Dict[str, Tensor]:
    """Random parameters with the reference's key names / shapes / init distributions.

    Distributions follow ``networks/base.py:5-56`` (xavier_normal gain 1 on the PointNet convs,
    zero biases), ``nn.Linear`` defaults for the four projections (``sg_aligner.py:54-67``) and
    PyG glorot for GATConv.  Values are NOT bit-identical to a reference construction (different
    RNG consumption order); parity tests always copy one parameter set into both sides.
    """
    g = torch.Generator().manual_seed(seed)
    p: Dict[str, Tensor] = {}

    def linear(name, fan_in, fan_out):
        bound = 1.0 / math.sqrt(fan_in)
        p[name + '.weight'] = (torch.rand(fan_out, fan_in, generator=g) * 2 - 1) * bound
        p[name + '.bias'] = (torch.rand(fan_out, generator=g) * 2 - 1) * bound

    linear('meta_embedding_rel', rel_dim, emb_dim)
    linear('meta_embedding_attr', attr_dim, emb_dim)
    chans = [3, 64, 128, pt_out_dim]
    for i in range(3):
        cin, cout = chans[i], chans[i + 1]
        std = math.sqrt(2.0 / (cin + cout))
        p[f'object_encoder.conv{i+1}.weight'] = torch.randn(cout, cin, 1, generator=g) * std
        p[f'object_encoder.conv{i+1}.bias'] = torch.zeros(cout)
        p[f'object_encoder.bn{i+1}.weight'] = torch.ones(cout)
        p[f'object_encoder.bn{i+1}.bias'] = torch.zeros(cout)
        p[f'object_encoder.bn{i+1}.running_mean'] = torch.zeros(cout)
        p[f'object_encoder.bn{i+1}.running_var'] = torch.ones(cout)
        p[f'object_encoder.bn{i+1}.num_batches_tracked'] = torch.zeros((), dtype=torch.long)
    linear('object_embedding', pt_out_dim, emb_dim)
    n_layers = len(hidden_units) - 1
    for i in range(n_layers):
        cin = hidden_units[i] * heads[i - 1] if i else hidden_units[i]
        H, C = heads[i], hidden_units[i + 1]
        a = math.sqrt(6.0 / (cin + H * C))
        w = (torch.rand(H * C, cin, generator=g) * 2 - 1) * a
        pre = f'structure_encoder.layer_stack.{i}'
        p[pre + '.lin_src.weight'] = w
        p[pre + '.lin_dst.weight'] = w
        a = math.sqrt(6.0 / (H + C))
        p[pre + '.att_src'] = (torch.rand(1, H, C, generator=g) * 2 - 1) * a
        p[pre + '.att_dst'] = (torch.rand(1, H, C, generator=g) * 2 - 1) * a
        p[pre + '.bias'] = torch.zeros(H * C)
    linear('structure_embedding', 256, emb_dim)
    p['fusion.weight'] = torch.ones(len(modules), 1)
    return {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in p.items()}


# --------------------------------------------------------------------------------------
# encoder pieces
# --------------------------------------------------------------------------------------
def pointnet_feat(pts: Tensor, p: Dict[str, Tensor], prefix: str = 'object_encoder') -> Tensor:
    """``pointnet.py:140-163``: three 1x1 convs (= per-point affine maps) each followed by ReLU,
    then a max over the points.  The BatchNorm layers are invoked but their outputs discarded
    (``pointnet.py:141-142,154-155,158-159``) so they do not appear here.  ``pts`` is ``[N,P,3]``
    (the layout of ``data_dict['tot_obj_pts']``; the reference permutes to ``[N,3,P]``)."""
    h = pts
    for i in (1, 2, 3):
        w = p[f'{prefix}.conv{i}.weight'].squeeze(-1)      # [out,in]
        b = p[f'{prefix}.conv{i}.bias']
        h = torch.relu(h @ w.t() + b)
    return h.max(dim=1).values                              # [N, out]


def pointnet_feat_reference_ops(pts: Tensor, p: Dict[str, Tensor], training: bool = False, prefix: str = 'object_encoder') -> Tensor:
    """The same function with the reference's OWN op sequence (``pointnet.py:140-163``): permute to [N,3,P], ``Conv1d``
    (kernel size 1), the BatchNorm call whose output is discarded, ReLU, max over the last axis.  Used by the
    eager-CUDA baseline leg of ``bench.py`` (cuDNN / ATen on the same B200) -- numerically equal to
    :func:`pointnet_feat` up to summation order (and to cuDNN's TF32 convolutions when torch's default allows them)."""
    x = pts.permute(0, 2, 1)
    for i in (1, 2, 3):
        x = F.conv1d(x, p[f'{prefix}.conv{i}.weight'], p[f'{prefix}.conv{i}.bias'])
        _ = F.batch_norm(x, p[f'{prefix}.bn{i}.running_mean'], p[f'{prefix}.bn{i}.running_var'], p[f'{prefix}.bn{i}.weight'],
                         p[f'{prefix}.bn{i}.bias'], training, 0.1, 1e-5)          # invoked, result dropped
        x = F.relu(x)
    return torch.max(x, 2, keepdim=True)[0].view(x.shape[0], -1)


def pointnet_bn_batch_stats(pts: Tensor, p: Dict[str, Tensor], prefix: str = 'object_encoder'):
    """Per-channel batch mean / unbiased variance of the three *pre-ReLU* conv outputs: what a
    train-mode forward folds into ``bn{1,2,3}.running_*`` (momentum 0.1) as a side effect."""
    h = pts
    out = []
    for i in (1, 2, 3):
        w = p[f'{prefix}.conv{i}.weight'].squeeze(-1)
        z = h @ w.t() + p[f'{prefix}.conv{i}.bias']
        flat = z.reshape(-1, z.shape[-1])
        out.append((flat.mean(0), flat.var(0, unbiased=True)))
        h = torch.relu(z)
    return out


def gat_conv(x: Tensor, edge_index: Tensor, w: Tensor, att_src: Tensor, att_dst: Tensor,
             bias: Tensor, heads: int, negative_slope: float = 0.2) -> Tensor:
    """PyG-2.2.0 ``GATConv`` (concat heads, self loops removed then one added per node,
    ``flow=source_to_target``: row 0 = source j, row 1 = target i, duplicate edges kept)."""
    n = x.shape[0]
    H = heads
    C = w.shape[0] // H
    xs = (x @ w.t()).view(n, H, C)
    a_s = (xs * att_src.view(1, H, C)).sum(-1)              # [n,H]
    a_d = (xs * att_dst.view(1, H, C)).sum(-1)
    ei = edge_index.long()
    keep = ei[0] != ei[1]
    loops = torch.arange(n, dtype=torch.long, device=x.device)
    j = torch.cat([ei[0][keep], loops])
    i = torch.cat([ei[1][keep], loops])
    e = F.leaky_relu(a_s[j] + a_d[i], negative_slope)       # [E',H]
    m = torch.full((n, H), -float('inf'), dtype=e.dtype, device=x.device)
    m = m.scatter_reduce(0, i.view(-1, 1).expand(-1, H), e, reduce='amax', include_self=True)
    pexp = torch.exp(e - m[i])
    den = torch.zeros((n, H), dtype=e.dtype, device=x.device).index_add_(0, i, pexp) + 1e-16
    alpha = pexp / den[i]
    out = torch.zeros((n, H, C), dtype=x.dtype, device=x.device).index_add_(0, i, alpha.unsqueeze(-1) * xs[j])
    return out.reshape(n, H * C) + bias


def multi_gat(x: Tensor, edge_index: Tensor, p: Dict[str, Tensor], heads: Sequence[int],
              prefix: str = 'structure_encoder') -> Tensor:
    """``gat.py:40-48``: GATConv layers with ELU in between (dropout p=0 is the identity)."""
    n_layers = len(heads)
    for li in range(n_layers):
        pre = f'{prefix}.layer_stack.{li}'
        x = gat_conv(x, edge_index, p[pre + '.lin_src.weight'], p[pre + '.att_src'],
                     p[pre + '.att_dst'], p[pre + '.bias'], heads[li])
        if li + 1 < n_layers:
            x = F.elu(x)
    return x


def fusion(embs: List[Tensor], weight: Tensor) -> Tensor:
    """``sg_aligner.py:30-35``: softmax over the modality weights, each modality L2-normalised
    (eps 1e-12) and scaled, concatenated along the feature axis."""
    wn = torch.softmax(weight, dim=0)
    return torch.cat([wn[k] * F.normalize(e, dim=1) for k, e in enumerate(embs)], dim=1)


def encoder_forward(p: Dict[str, Tensor], data: dict, modules: Sequence[str],
                    heads=(2, 2), reference_ops: bool = False) -> Dict[str, Tensor]:
    """``sg_aligner.py:71-137``.  ``data`` follows the Scan3R collate contract
    (``src/datasets/scan3r.py:179-209``).  Device-agnostic (the eager-CUDA baseline leg of ``bench.py`` passes CUDA
    tensors); ``reference_ops`` selects the Conv1d + discarded-BatchNorm op sequence for the point encoder."""
    dt = p['object_embedding.weight'].dtype
    pts = data['tot_obj_pts'].to(dt)
    attr = data['tot_bow_vec_object_attr_feats'].to(dt)
    rel = data['tot_bow_vec_object_edge_feats'].to(dt)
    pose = data['tot_rel_pose'].to(dt)
    embs: Dict[str, Tensor] = {}
    for mod in modules:
        if mod == 'gat':
            outs = []
            o = 0
            e = 0
            for b in range(int(data['batch_size'])):
                for side in (0, 1):
                    n = int(data['graph_per_obj_count'][b][side])
                    ne = int(data['graph_per_edge_count'][b][side])
                    ei = data['edges'][e:e + ne].t()
                    outs.append(multi_gat(pose[o:o + n], ei, p, heads))
                    o += n
                    e += ne
            emb = torch.cat(outs) @ p['structure_embedding.weight'].t() + p['structure_embedding.bias']
        elif mod == 'point':
            feat = pointnet_feat_reference_ops(pts, p) if reference_ops else pointnet_feat(pts, p)
            emb = feat @ p['object_embedding.weight'].t() + p['object_embedding.bias']
        elif mod == 'rel':
            emb = rel @ p['meta_embedding_rel.weight'].t() + p['meta_embedding_rel.bias']
        elif mod == 'attr':
            emb = attr @ p['meta_embedding_attr.weight'].t() + p['meta_embedding_attr.bias']
        else:
            raise NotImplementedError(mod)
        embs[mod] = emb
    if len(modules) > 1:
        embs['joint'] = fusion([embs[m] for m in modules], p['fusion.weight'])
    return embs


# --------------------------------------------------------------------------------------
# losses  (losses.py:5-152)
# --------------------------------------------------------------------------------------
def prob_dist(ai: Tensor, bi: Tensor, aj: Tensor, bj: Tensor, temp: float) -> Tensor:
    """``losses.py:5-15``.  The two normalisers are sums over the WHOLE matrices (scalars)."""
    mx = torch.exp(ai @ bi.t() / temp)
    s_a = torch.exp(ai @ aj.t() / temp).sum()
    s_b = torch.exp(ai @ bj.t() / temp).sum()
    r1 = mx / (s_a + 1e-9)
    r2 = mx / (s_b + 1e-9)
    inv = 1.0 + 1.0 / (r1 + 1e-9) + 1.0 / (r2 + 1e-9)
    return 1.0 / (inv + 1e-9)


def _gather4(emb: Tensor, data: dict):
    idx = [torch.as_tensor(np.asarray(data[k]), dtype=torch.long) for k in ('e1i', 'e2i', 'e1j', 'e2j')]
    return [emb[i] for i in idx]


def icl_loss(emb: Tensor, data: dict, temp: float = 0.1, alpha: float = 0.5) -> Tensor:
    """``losses.py:43-58``: temperature hard-wired to 0.1; the 2->1 matrix is added
    un-transposed; mean over the full A x A matrix."""
    z = F.normalize(emb, dim=1)
    e1i, e2i, e1j, e2j = _gather4(z, data)
    q12 = prob_dist(e1i, e2i, e1j, e2j, temp)
    q21 = prob_dist(e2i, e1i, e2j, e1j, temp)
    return -torch.log(alpha * q12 + (1 - alpha) * q21).mean()


def ial_loss(modal_emb: Tensor, joint_emb: Tensor, data: dict, temp: float = 1.0,
             alpha: float = 0.5, zoom: float = 0.1) -> Tensor:
    """``losses.py:68-97`` as called at ``losses.py:122`` (first arg = modal, second = joint).
    ``KLDivLoss(reduction='sum', log_target=True)(input=log qm, target=qo)`` evaluates
    ``sum(exp(qo) * (qo - log qm))`` -- the target is not in log space; reproduced as-is."""
    zo = F.normalize(modal_emb, dim=1)
    zm = F.normalize(joint_emb, dim=1)
    o = _gather4(zo, data)
    m = _gather4(zm, data)
    qo12 = prob_dist(o[0], o[1], o[2], o[3], temp)
    qo21 = prob_dist(o[1], o[0], o[3], o[2], temp)
    qm12 = prob_dist(m[0], m[1], m[2], m[3], temp)
    qm21 = prob_dist(m[1], m[0], m[3], m[2], temp)
    la = (torch.exp(qo12) * (qo12 - torch.log(qm12))).sum()
    lb = (torch.exp(qo21) * (qo21 - torch.log(qm21))).sum()
    return zoom * (alpha * la + (1 - alpha) * lb)


def multi_loss(losses: List[Tensor], log_vars: Tensor) -> Tensor:
    """``losses.py:28-34``: sum_i exp(-s_i) L_i + s_i."""
    tot = 0
    for i, l in enumerate(losses):
        tot = tot + torch.exp(-log_vars[i]) * l + log_vars[i]
    return tot


def overall_loss(out: Dict[str, Tensor], data: dict, modules: Sequence[str],
                 log_vars_ial: Tensor, log_vars_icl: Tensor, zoom: float = 0.1) -> Dict[str, Tensor]:
    """``losses.py:114-152``."""
    if len(modules) > 1:
        ial = multi_loss([ial_loss(out[m], out['joint'], data) for m in modules], log_vars_ial) * zoom
        icl_uni = multi_loss([icl_loss(out[m], data) for m in modules], log_vars_icl)
        icl_multi = icl_loss(out['joint'], data)
        loss = ial + icl_uni + icl_multi
    else:
        ial = 0.0
        icl_multi = 0.0
        icl_uni = icl_loss(out[modules[0]], data)
        loss = icl_uni
    return {'loss': loss, 'icl_loss_unimodal': icl_uni, 'icl_loss_multimodal': icl_multi, 'ial_loss': ial}


# --------------------------------------------------------------------------------------
# matching head + rank metrics
# --------------------------------------------------------------------------------------
def pair_offsets(data: dict) -> np.ndarray:
    cnt = np.asarray(data['graph_per_obj_count']).reshape(-1, 2).sum(1)
    return np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)


def match_pair(emb: Tensor):
    """``inference_align_reg.py:125-128``: L2-normalise (no eps), ``sim = 1 - E E^T`` over
    source AND reference nodes of the pair, rows sorted ascending.  ``stable=True`` pins the
    tie order (lowest index first) that the product kernels also use; the reference's
    ``argsort`` is unstable so it does not define one."""
    e = emb / emb.norm(dim=1)[:, None]
    sim = 1 - e @ e.t()
    rank = torch.argsort(sim, dim=1, stable=True)
    return sim, rank


def ranks_without_self(rank_row: np.ndarray, self_idx: int) -> np.ndarray:
    return rank_row[rank_row != self_idx]


def hits_and_rr(rank: np.ndarray, e1i: np.ndarray, e2i: np.ndarray, ks=(1, 2, 3, 4, 5)):
    """``utils/alignment.py:3-25``: per anchor drop the node itself from its ranked row, then
    Hits@k <=> target within the first k, reciprocal rank = 1 / (1 + position of target)."""
    hits = {k: 0 for k in ks}
    rr = []
    for a, t in zip(e1i, e2i):
        row = ranks_without_self(rank[a], a)
        pos = int(np.nonzero(row == t)[0][0])
        rr.append(1.0 / (pos + 1))
        for k in ks:
            hits[k] += int(pos < k)
    return hits, rr


def sgar(sim: np.ndarray, rank: np.ndarray, e1i, e2i, modes=('2', '50', '100')):
    """``utils/alignment.py:27-58``."""
    pred, dist = [], []
    for a in e1i:
        row = ranks_without_self(rank[a], a)
        pred.append(int(row[0]))
        dist.append(float(sim[a, row[0]]))
    order = np.argsort(dist)
    vals = {}
    for mode in modes:
        sel = order[:2] if mode == '2' else (order[:len(order) // 2] if mode == '50' else order)
        vals[mode] = float(all(pred[i] == int(e2i[i]) for i in sel))
    return vals


def node_corrs(rank: np.ndarray, n_src: int, k: int = 1):
    """``utils/alignment.py:60-71``."""
    out = []
    for i in range(n_src):
        row = ranks_without_self(rank[i], i)[:k]
        out.extend((i, int(r)) for r in row if r >= n_src)
    return out


def alignment_score(rank: np.ndarray, n_src: int, n_ref: int) -> float:
    """``utils/alignment.py:79-89``."""
    c = sum(int(ranks_without_self(rank[i], i)[0] >= n_src) for i in range(n_src))
    return c / n_ref


def evaluate_batch(emb: Tensor, data: dict, ks=(1, 2, 3, 4, 5)):
    """The per-pair loop of ``inference_align_reg.py:107-145`` restricted to the alignment
    metrics.  Returns per-pair rank lists plus aggregated Hits@k / MRR."""
    offs = pair_offsets(data)
    e1c = np.asarray(data['e1i_count']).reshape(-1)
    a0 = 0
    ranks, sims = [], []
    hits_tot = {k: 0 for k in ks}
    rr_all: List[float] = []
    total = 0
    goc = np.asarray(data['graph_per_obj_count']).reshape(-1, 2)
    sgar_all = {m: [] for m in ('2', '50', '100')}
    align, corrs = [], []
    for b in range(int(data['batch_size'])):
        o0, o1 = int(offs[b]), int(offs[b + 1])
        na = int(e1c[b])
        e1 = np.asarray(data['e1i'][a0:a0 + na]).astype(np.int64) - o0
        e2 = np.asarray(data['e2i'][a0:a0 + na]).astype(np.int64) - o0
        a0 += na
        sim, rank = match_pair(emb[o0:o1])
        ranks.append(rank.cpu().numpy())             # the reference moves rank_list to the host per metric call
        sims.append(sim.cpu().numpy())
        ns, nr = int(goc[b, 0]), int(goc[b, 1])
        align.append(alignment_score(ranks[-1], ns, nr) if nr else 0.0)
        corrs.append(node_corrs(ranks[-1], ns, 1))
        if na:
            h, rr = hits_and_rr(ranks[-1], e1, e2, ks)
            for k in ks:
                hits_tot[k] += h[k]
            rr_all.extend(rr)
            total += na
            sv = sgar(sims[-1], ranks[-1], e1, e2)      # inference_align_reg.py:139-141
            for m in sgar_all:
                sgar_all[m].append(sv[m])
    return {'rank': ranks, 'sim': sims, 'hits': hits_tot, 'total': total,
            'mrr': float(np.mean(rr_all)) if rr_all else 0.0, 'sgar': sgar_all, 'alignment_score': align,
            'node_corrs': corrs}
