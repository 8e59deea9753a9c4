"""``torch.autograd.Function`` glue: forward and backward of every stage are CUDA kernels of
``libsga_b200.so``; autograd only routes the tensors between them."""
from __future__ import annotations

import torch

from . import ops


class PointNetFeat(torch.autograd.Function):
    """PointNetfeat.forward (pointnet.py:140-163) as one fused kernel; the backward goes through
    the saved per-channel argmax of the max-pool."""

    @staticmethod
    def forward(ctx, pts, W1, b1, W2, b2, W3, b3, mode, chunks=None, want_stats=False, grad_mode=True, cta_cap=0):
        """Returns (pooled feature, moments); moments (f64, non-differentiable) is None unless want_stats.
        ``grad_mode``: torch.is_grad_enabled() at the call site -- inside forward() autograd is always off and
        needs_input_grad mirrors requires_grad of the parameters even under torch.no_grad(), so without it the
        serving path would run the (slower) argmax-tracking variant of the kernel."""
        need = grad_mode and any(ctx.needs_input_grad[1:7])
        mom = None
        if want_stats:
            if chunks:      # statistics span the whole batch: one launch once every chunk has landed
                for (_, _, ev) in chunks:
                    torch.cuda.current_stream().wait_event(ev)
            out, arg, mom = ops.pointnet_forward_stats(pts, W1, b1, W2, b2, W3, b3, want_argmax=need)
        else:
            out, arg = ops.pointnet_forward(pts, W1, b1, W2, b2, W3, b3, want_argmax=need, mode=mode, chunks=chunks)
        if need:
            ctx.save_for_backward(pts, W1, b1, W2, b2, W3, b3, out, arg)
            ctx.mode = mode if W3.shape[0] % 128 == 0 else ops.POINTNET_SIMT
            ctx.params = (W1, b1, W2, b2, W3, b3)
            ctx.cta_cap = cta_cap       # SMs left to a concurrent branch in the forward are left to its backward too
        if mom is not None:
            ctx.mark_non_differentiable(mom)
        return out, mom

    @staticmethod
    def backward(ctx, gout, _gmom=None):
        pts, W1, b1, W2, b2, W3, b3, out, arg = ctx.saved_tensors
        into = [ops.grad_target(p) if need else None for p, need in zip(ctx.params, ctx.needs_input_grad[1:7])]
        if ctx.cta_cap:
            ops.pointnet_set_max_ctas(ctx.cta_cap)
        try:
            g = ops.pointnet_backward(pts, W1, b1, W2, b2, W3, b3, out, arg, gout.contiguous(), mode=ctx.mode, into=into)
        finally:
            if ctx.cta_cap:
                ops.pointnet_set_max_ctas(0)
        g = [None if t is None else t.view_as(p) for t, p in zip(g, ctx.params)]     # None: accumulated into p.grad
        return (None, *g, None, None, None, None, None)


class GATLayer(torch.autograd.Function):
    """One GATConv layer (+ optional ELU) over the block-diagonal batch graph."""

    @staticmethod
    def forward(ctx, x, W, att_src, att_dst, bias, graph, H, C, apply_elu):
        xs, a_s, a_d = ops.gat_linear(x, W, att_src, att_dst, H, C)
        out = ops.gat_aggregate(xs, a_s, a_d, graph, bias, apply_elu)
        ctx.graph, ctx.H, ctx.C, ctx.apply_elu = graph, H, C, apply_elu
        ctx.need_gx = x.requires_grad
        ctx.params = (W, att_src, att_dst, bias)
        ctx.save_for_backward(x, W, att_src, att_dst, xs, a_s, a_d, out)
        return out

    @staticmethod
    def backward(ctx, gout):
        x, W, att_src, att_dst, xs, a_s, a_d, out = ctx.saved_tensors
        need = ctx.needs_input_grad[1:5]
        tW, ts, td, tb = [ops.grad_target(p, n) for p, n in zip(ctx.params, need)]
        g_xs, g_as, g_ad, g_bias = ops.gat_aggregate_backward(xs, a_s, a_d, ctx.graph, ctx.apply_elu, out, gout.contiguous(), into_bias=tb)
        gW, g_att_s, g_att_d, gx = ops.gat_linear_backward(x, W, att_src, att_dst, ctx.H, ctx.C, xs, g_xs, g_as, g_ad, ctx.need_gx,
                                                           into=(tW, ts, td))
        return (gx, gW if need[0] else None,
                None if (g_att_s is None or not need[1]) else g_att_s.view_as(att_src),
                None if (g_att_d is None or not need[2]) else g_att_d.view_as(att_dst),
                g_bias if need[3] else None, None, None, None, None)


class ProjectFuse(torch.autograd.Function):
    """All modality projections + the fusion in one autograd node.

    forward(fusion_w, M, *[x_m, W_m, b_m]*M) -> (emb_0, ..., emb_{M-1}, joint)  (joint omitted if M == 1)
    """

    @staticmethod
    def forward(ctx, fusion_w, M, *args):
        xs, Ws, bs = args[0::3], args[1::3], args[2::3]
        N = xs[0].shape[0]
        out_dims = [W.shape[0] for W in Ws]
        dev = xs[0].device
        if len(set(out_dims)) == 1 and out_dims[0] <= 128 and M <= 8:
            embs, joint = ops.project_fuse_multi(xs, Ws, bs, fusion_w, want_joint=M > 1)      # one launch
        else:
            joint = torch.empty((N, sum(out_dims)), device=dev, dtype=torch.float32) if M > 1 else None
            embs, col = [], 0
            for m in range(M):
                embs.append(ops.project_fuse(xs[m], Ws[m], bs[m], joint, col, fusion_w, M, m))
                col += out_dims[m]
        ctx.M, ctx.out_dims = M, out_dims
        ctx.need_gx = [x.requires_grad for x in xs]
        ctx.params = (fusion_w, Ws, bs)
        ctx.save_for_backward(fusion_w, *xs, *Ws, *embs)
        return (*embs, joint) if M > 1 else (embs[0],)

    @staticmethod
    def backward(ctx, *gouts):
        M = ctx.M
        saved = ctx.saved_tensors
        fusion_w, xs, Ws, embs = saved[0], saved[1:1 + M], saved[1 + M:1 + 2 * M], saved[1 + 2 * M:1 + 3 * M]
        g_joint = gouts[M] if M > 1 else None
        if g_joint is not None:
            g_joint = g_joint.contiguous()
        p_fw, p_Ws, p_bs = ctx.params
        nig = ctx.needs_input_grad
        t_fw = ops.grad_target(p_fw, nig[0])
        # the kernels ADD into g_fusion_w: one buffer (the parameter's own .grad when it is kept allocated) for all M
        g_fw = t_fw if t_fw is not None else torch.zeros(M, device=fusion_w.device, dtype=torch.float32)
        grads, col = [], 0
        for m in range(M):
            g_emb = gouts[m]
            gW, gb, _, gx = ops.project_fuse_backward(xs[m], Ws[m], embs[m], None if g_emb is None else g_emb.contiguous(),
                                                      g_joint, col, fusion_w, M, m, ctx.need_gx[m],
                                                      into=(ops.grad_target(p_Ws[m], nig[3 + 3 * m]),
                                                            ops.grad_target(p_bs[m], nig[4 + 3 * m]), g_fw.view(-1)))
            grads += [gx, gW if nig[3 + 3 * m] else None, gb if nig[4 + 3 * m] else None]
            col += ctx.out_dims[m]
        return (None if (t_fw is not None or not nig[0]) else g_fw.view_as(fusion_w), None, *grads)


class OverallLossFn(torch.autograd.Function):
    """OverallLoss.forward (losses.py:114-152): value and gradient come out of one fused pass."""

    @staticmethod
    def forward(ctx, idx, zoom, lv_ial, lv_icl, *embs):
        want_grad = any(ctx.needs_input_grad[2:])
        losses, grads, g_ial, g_icl = ops.loss_forward_backward(embs, idx, lv_ial, lv_icl, zoom, want_grad,
                                                                partition=getattr(idx, 'partition', True))
        ctx.n = len(embs)
        ctx.has_lv = lv_ial is not None
        if want_grad:
            saved = list(grads) + ([g_ial, g_icl] if ctx.has_lv else [])
            ctx.save_for_backward(*saved)
        ctx.mark_non_differentiable()
        return losses

    @staticmethod
    def backward(ctx, g):
        # only losses[0] (= 'loss') carries gradient; the other three entries are reporting values
        s = g[0]
        saved = list(ctx.saved_tensors)
        scaled = torch._foreach_mul(saved, s)          # one multi-tensor launch for all embeddings (+ log_vars)
        grads = scaled[:ctx.n]
        if ctx.has_lv:
            g_ial, g_icl = scaled[ctx.n], scaled[ctx.n + 1]
        else:
            g_ial = g_icl = None
        return (None, None, g_ial, g_icl, *grads)


class StandaloneIAL(torch.autograd.Function):
    """``IALLoss.forward(src_emb, ref_emb)`` (losses.py:68-97) on its own.  The fused kernel evaluates
    ``zoom * IAL + ICL terms`` (log_vars = 0); the value is its ``ial`` output at zoom = 1 and, the loss being affine
    in zoom, the gradient is (fused gradient at zoom 1) - (fused gradient at zoom 0).  Off the hot path (the trainer
    goes through ``OverallLoss``): two fused passes are fine."""

    @staticmethod
    def forward(ctx, idx, src_emb, ref_emb):
        dev = src_emb.device
        zero = torch.zeros(1, device=dev)
        want_grad = any(ctx.needs_input_grad[1:])
        part = getattr(idx, 'partition', True)
        l1, g1, _, _ = ops.loss_forward_backward([src_emb, ref_emb], idx, zero, zero, 1.0, want_grad, partition=part)
        if want_grad:
            _, g0, _, _ = ops.loss_forward_backward([src_emb, ref_emb], idx, zero, zero, 0.0, True, partition=part)
            ctx.save_for_backward(g1[0] - g0[0], g1[1] - g0[1])
        return l1[3].clone()

    @staticmethod
    def backward(ctx, g):
        gs, gr = ctx.saved_tensors
        return None, gs * g, gr * g


# ------------------------------------------------------------------------------------------- EVA baseline
class GCNLayer(torch.autograd.Function):
    """One PyG-2.2.0 ``GCNConv`` (gat.py:15,21) over the block-diagonal batch graph: linear map without bias, symmetric
    normalisation, sum aggregation, bias afterwards (+ the ReLU between layers, gat.py:22-23)."""

    @staticmethod
    def forward(ctx, x, W, bias, graph, graph_t, relu):
        xw = ops.linear_nobias(x, W)
        out = ops.gcn_aggregate(xw, graph, graph.row_cnt, bias, relu)
        ctx.graph, ctx.graph_t, ctx.relu = graph, graph_t, relu
        ctx.need_gx = x.requires_grad
        ctx.save_for_backward(x, W, out)
        return out

    @staticmethod
    def backward(ctx, gout):
        x, W, out = ctx.saved_tensors
        g = gout.contiguous()
        if ctx.relu:
            g = ops.relu_mask(g, out)
        g_bias = ops.colsum_rows(g) if ctx.needs_input_grad[2] else None
        g_xw = ops.gcn_aggregate(g, ctx.graph_t, ctx.graph.row_cnt)        # transpose of the normalised adjacency
        gW, gx = ops.linear_nobias_backward(x, W, g_xw, ctx.need_gx)
        return gx, (gW if ctx.needs_input_grad[1] else None), g_bias, None, None, None


class FuseRows(torch.autograd.Function):
    """MultiModalFusion (sg_aligner.py:23-35) over embeddings of different widths: forward(fusion_w, *embs) -> joint."""

    @staticmethod
    def forward(ctx, fusion_w, *embs):
        joint = ops.fuse_rows(embs, fusion_w)
        ctx.save_for_backward(fusion_w, *embs)
        return joint

    @staticmethod
    def backward(ctx, g_joint):
        fusion_w, *embs = ctx.saved_tensors
        gxs, g_fw = ops.fuse_rows_backward(embs, fusion_w, g_joint.contiguous())
        return (g_fw.view_as(fusion_w) if ctx.needs_input_grad[0] else None,
                *[g if need else None for g, need in zip(gxs, ctx.needs_input_grad[1:])])


def _nca_forward(ctx, normalize, emb, e1, e2, alpha, beta, ep):
    want = ctx.needs_input_grad[0]
    loss, g = ops.nca_loss_forward_backward(emb, e1, e2, alpha, beta, ep, want, normalize=normalize)
    if want:
        ctx.save_for_backward(g)
    return loss.clone()


class NCAFn(torch.autograd.Function):
    """NCALoss of one embedding (losses.py:161-176 as called at :189-200: F.normalize + the e1i / e2i gathers included);
    the gradient is produced with the value."""

    @staticmethod
    def forward(ctx, emb, e1, e2, alpha, beta, ep):
        return _nca_forward(ctx, True, emb, e1, e2, alpha, beta, ep)

    @staticmethod
    def backward(ctx, gl):
        (g,) = ctx.saved_tensors
        return g * gl, None, None, None, None, None


class NCAPairFn(torch.autograd.Function):
    """NCALoss.forward(src_emb, ref_emb) on rows that are already normalised and gathered (losses.py:161-176)."""

    @staticmethod
    def forward(ctx, emb, e1, e2, alpha, beta, ep):
        return _nca_forward(ctx, False, emb, e1, e2, alpha, beta, ep)

    @staticmethod
    def backward(ctx, gl):
        (g,) = ctx.saved_tensors
        return g * gl, None, None, None, None, None
