"""B200-native ``NaivePCT`` object encoder (``src/aligner/networks/pct.py:275-317``; the ``'pct'`` module of
``MultiModalEncoder``, ``sg_aligner.py:59-60`` -- the point encoder the shipped ground-truth config selects,
``configs/scan3r/scan3r_ground_truth.yaml:5``).

Same sub-module / parameter names as the reference (``embedding.conv1`` ... ``sa4.after_norm``, ``linear.0/1``,
``linear1``, ``linear2``, ``bn1``, ``bn2``; ``q_conv.weight`` IS ``k_conv.weight``, ``pct.py:199``) so a reference
``state_dict`` loads strictly.  The torch modules are parameter containers; the arithmetic runs in the tcgen05 kernels of
``csrc/pct_*.cu``:

    points -> [pct_embed] -> z2 -> 4 x { [pct_pointwise k|v] -> [pct_attn_stats] -> [pct_attn] -> [pct_pointwise trans] }
           -> [pct_cat_linear: concat + 512->1024 conv + max/min pooling] -> head (two small GEMMs)

Every BatchNorm is applied as a folded affine pair by the NEXT kernel's prologue (``ops.bn_fold``), so ``train()``
(batch statistics over the whole batch, running-statistics side effect) and ``eval()`` (running statistics) run the same
kernels; the residual ``x = x + x_s`` (``pct.py:230``) and the ReLUs live in those prologues too.

The backward (``_PCTFunction.backward``) differentiates every BatchNorm in closed form (``csrc/pct_bwd.cu``), the
attention through three more tcgen05 products per layer (``pct_attn_bwd_dv`` / ``pct_attn_bwd_dk``), and the pooled
512 -> 1024 convolution without ever forming its [N, P, 1024] output or gradient (section "concat stage" below).
``P <= 512`` points per object.  The two ``nn.Dropout(0.5)`` masks are drawn with torch's generator on the device (``dropout_rng = 'cpu'``
draws them with the CPU generator in the reference's order instead -- what the parity tests against CPU goldens use).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops


class Embedding(nn.Module):
    """``pct.py:101-125`` parameter container."""

    def __init__(self, in_channels=3, out_channels=128):
        super().__init__()
        self.conv1 = nn.Conv1d(in_channels, out_channels, kernel_size=1, bias=False)
        self.conv2 = nn.Conv1d(out_channels, out_channels, kernel_size=1, bias=False)
        self.bn1 = nn.BatchNorm1d(out_channels)
        self.bn2 = nn.BatchNorm1d(out_channels)


class SA(nn.Module):
    """``pct.py:187-209`` parameter container; q and k share one weight tensor."""

    def __init__(self, channels):
        super().__init__()
        self.da = channels // 4
        self.q_conv = nn.Conv1d(channels, channels // 4, 1, bias=False)
        self.k_conv = nn.Conv1d(channels, channels // 4, 1, bias=False)
        self.q_conv.weight = self.k_conv.weight
        self.v_conv = nn.Conv1d(channels, channels, 1)
        self.trans_conv = nn.Conv1d(channels, channels, 1)
        self.after_norm = nn.BatchNorm1d(channels)


class _PCTFunction(torch.autograd.Function):
    """The whole encoder as one autograd node: ``apply(module, points, *parameters)``.  The forward keeps the per-layer
    activations the kernels wrote anyway (z2, k, v, x_s, t, x_l: ~4.5 GB per SA layer at 4096 objects x 512 points,
    sized for 180 GB of HBM); the backward returns one gradient per parameter."""

    @staticmethod
    def forward(ctx, mod, pts, *params):
        N, P, _ = pts.shape
        out, saved = mod._forward(pts, N, P, mod.training, save=True)
        ctx.mod, ctx.saved, ctx.training, ctx.params = mod, saved, mod.training, params
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_out):
        # ctx.saved is left intact: the reference calls loss.backward(retain_graph=True) (a second pass over the same graph
        # must find its activations); autograd drops the node -- and with it the tensors -- when the graph is released
        grads = ctx.mod._backward(ctx.saved, g_out.contiguous(), ctx.training)
        return (None, None) + tuple(grads.get(id(p)) if ctx.needs_input_grad[2 + i] else None for i, p in enumerate(ctx.params))


class NaivePCT(nn.Module):
    def __init__(self):
        super().__init__()
        self.embedding = Embedding(3, 128)
        self.sa1 = SA(128)
        self.sa2 = SA(128)
        self.sa3 = SA(128)
        self.sa4 = SA(128)
        self.linear = nn.Sequential(nn.Conv1d(512, 1024, kernel_size=1, bias=False), nn.BatchNorm1d(1024), nn.LeakyReLU(negative_slope=0.2))
        self.linear1 = nn.Linear(1024, 512, bias=False)
        self.linear2 = nn.Linear(512, 256)
        self.bn1 = nn.BatchNorm1d(512)
        self.bn2 = nn.BatchNorm1d(256)
        self.dp1 = nn.Dropout(p=0.5)
        self.dp2 = nn.Dropout(p=0.5)
        self.dropout_rng = 'device'

    def _mask(self, n, c, dev):
        if self.dropout_rng == 'cpu':      # the reference's F.dropout on CPU tensors: empty_like(x).bernoulli_(1 - p)
            return torch.empty(n, c).bernoulli_(0.5).to(dev)
        return torch.empty(n, c, device=dev).bernoulli_(0.5)

    def forward(self, pts_npc: torch.Tensor) -> torch.Tensor:
        """``pts_npc``: [N, P, 3] (the collated layout; the reference permutes to [N, 3, P] first).  -> [N, 256]"""
        if not pts_npc.is_cuda:
            raise RuntimeError('sgaligner_b200.NaivePCT needs CUDA tensors (no CPU fallback)')
        N, P, _ = pts_npc.shape
        if P > 512:
            raise NotImplementedError('NaivePCT kernels hold a [128 x P] score block in tensor memory: P <= 512')
        tr = self.training
        if tr and N == 1:
            raise ValueError('Expected more than 1 value per channel when training')      # torch's BatchNorm1d on [1, C]
        params = [p for p in self.parameters()]
        if torch.is_grad_enabled() and any(p.requires_grad for p in params):
            return _PCTFunction.apply(self, pts_npc.detach(), *params)
        with torch.no_grad():
            return self._forward(pts_npc, N, P, tr)[0]

    def _forward(self, pts, N, P, tr, save: bool = False):
        emb = self.embedding
        cnt = float(N) * float(P)
        pts = ops._f32c(pts)
        S = {'pts': pts, 'N': N, 'P': P} if save else None
        # ---- Embedding (pct.py:120-125)
        mom = ops.pct_point_moments(pts) if (tr or save) else None
        st1 = ops.pct_affine_stats(mom, emb.conv1.weight) if tr else None
        ab1 = ops.bn_fold(emb.bn1, st1, cnt, tr)
        z2, st2 = ops.pct_embed(pts, emb.conv1.weight, ab1[0], ab1[1], emb.conv2.weight, tr)
        ab2 = ops.bn_fold(emb.bn2, st2, cnt, tr)
        if save:
            S.update(mom=mom, st1=st1, ab1=ab1, z2=z2, st2=st2, ab2=ab2, layers=[])
        # ---- four self-attention layers (pct.py:211-232); x_l = x_{l-1} + relu(after_norm(t_l)) is formed by the next prologue
        src1, g1 = z2, ab2                 # x0 = relu(bn2(z2))
        src2, g2 = None, None
        xs_saved, t, abt = [], None, None
        for li, sa in enumerate((self.sa1, self.sa2, self.sa3, self.sa4)):
            Wkv = torch.cat([sa.k_conv.weight.reshape(32, 128), sa.v_conv.weight.reshape(128, 128)])
            bkv = torch.cat([torch.zeros(32, device=pts.device), sa.v_conv.bias])
            v_absmax = None
            if save and ops.pw2_records_absmax():      # training: the backward's operand scale needs max |v| per object
                k, v, x_in, v_absmax = ops.pct_pointwise_kv(src1, g1, src2, g2, Wkv, bkv, want_x=True)
            else:
                k, v, x_in, _ = ops.pct_pointwise(src1, g1, src2, g2, Wkv, bkv, 32, want_x=(li > 0 or save), want_stats=False)
            if li > 0:
                xs_saved.append(x_in)      # x1, x2, x3
            x_s, c2 = ops.pct_attention(k, v, want_c2=True)
            t, _, _, stt = ops.pct_pointwise(x_s, None, None, None, sa.trans_conv.weight.reshape(128, 128), sa.trans_conv.bias, 128,
                                             want_x=False, want_stats=tr)
            abt = ops.bn_fold(sa.after_norm, stt, cnt, tr)
            if save:
                S['layers'].append(dict(x_in=x_in, k=k, v=v, c2=c2, x_s=x_s, t=t, stt=stt, abt=abt, v_absmax=v_absmax))
            if li == 0:
                src2, g2 = t, abt          # x1 = relu(bn2(z2)) + relu(after_norm(t1))
            else:
                src1, g1, src2, g2 = x_in, None, t, abt
        x1, x2, x3 = xs_saved
        # ---- concat + linear (512 -> 1024) + BN + LeakyReLU + max over points (pct.py:306-310)
        zmax, zmin, stl, imax, imin = ops.pct_cat_linear(x1, x2, x3, t, abt, self.linear[0].weight.reshape(1024, 512), track=save)
        abl = ops.bn_fold(self.linear[1], stl if tr else None, cnt, tr)
        if save:
            pooled, pstar, zsel = ops.pct_pool_act(zmax, zmin, abl[0], abl[1], P, imax, imin)
            S.update(stl=stl, abl=abl, pstar=pstar, zsel=zsel, pooled=pooled)
        else:
            pooled = ops.pct_pool_act(zmax, zmin, abl[0], abl[1], P)
        # ---- head (pct.py:311-316)
        y1 = ops.gemm_tf32x3(pooled, self.linear1.weight, N, 512, 1024)
        sy1 = ops.col_stats(y1) if tr else None
        abh1 = ops.bn_fold(self.bn1, sy1, float(N), tr)
        m1 = self._mask(N, 512, pts.device) if tr else None
        h1 = ops.bn_act_rows(y1, abh1[0], abh1[1], m1, 2.0)
        y2 = ops.gemm_tf32x3(h1, self.linear2.weight, N, 256, 512)
        sy2 = ops.col_stats(y2) if tr else None
        abh2 = ops.bn_fold(self.bn2, sy2, float(N), tr, lin_bias=self.linear2.bias)
        m2 = self._mask(N, 256, pts.device) if tr else None
        out = ops.bn_act_rows(y2, abh2[0], abh2[1], m2, 2.0)
        if save:
            S.update(y1=y1, sy1=sy1, abh1=abh1, m1=m1, h1=h1, y2=y2, sy2=sy2, abh2=abh2, m2=m2)
        return out, S

    # ------------------------------------------------------------------------------------------------ backward
    def _backward(self, S, g_out, tr):
        """d(loss)/d(parameter) for every parameter, keyed by ``id(parameter)``; ``g_out`` [N,256] = d(loss)/d(output).

        Head and SA layers: plain chain rule, BatchNorm in closed form (``ops.bn_backward``), weight gradients as grouped
        contractions over all N*P points (``ops.pct_wgrad``), input gradients on the pointwise tensor-core kernel (fp16 pairs, per-object power-of-two scaling of the gradient operand).

        Concat stage (pct.py:306-310): z = WL xcat is never stored.  With gy = d/d(BN output) -- non-zero only at the
        arg-max point p*(n, c) -- the BatchNorm formula gives  dz[n,p,c] = a_c gy [p = p*] - e_c - f_c (z[n,p,c] - mean_c),
        and mean = WL xbar (no bias), hence with xc = xcat - xbar
            d xcat[n,p,:] = (sparse) sum_{c: p* = p} a_c gy WL[c,:]  -  WL^T e  -  (WL^T diag(f) WL) xc[n,p,:]
            d WL[c,:]     = (sparse) sum_n a_c gy xcat[n,p*,:]      -  e_c sum xcat  -  f_c (WL Gc)[c,:],  Gc = sum xc xc^T
        i.e. one 512 -> 512 streaming product of the (centred) inputs, one 512 x 512 Gram matrix and two gather / scatter
        passes.  The centring is not cosmetic: M xcat and M xbar agree in their leading digits."""
        N, P, pts = S['N'], S['P'], S['pts']
        dev = pts.device
        cnt = float(N) * float(P)
        grads = {}

        def put(p, g):
            g = g.reshape(p.shape)
            grads[id(p)] = g if id(p) not in grads else grads[id(p)] + g

        def colsum(x, C):
            return ops.col_stats(x.reshape(-1, C))[:C].float()

        # ---- head (pct.py:311-316)
        dy2, dga, dbe, _ = ops.bn_backward(g_out, S['y2'], S['abh2'], self.bn2, S['sy2'], float(N), tr, mask=S['m2'], scale=2.0,
                                           lin_bias=self.linear2.bias)
        put(self.bn2.weight, dga); put(self.bn2.bias, dbe)
        put(self.linear2.bias, colsum(dy2, 256))
        put(self.linear2.weight, ops.gemm_tf32x3(dy2, S['h1'], 256, 512, N, a_mn=True, b_mn=True))
        dh1 = ops.gemm_tf32x3(dy2, self.linear2.weight, N, 512, 256, b_mn=True)
        dy1, dga, dbe, _ = ops.bn_backward(dh1, S['y1'], S['abh1'], self.bn1, S['sy1'], float(N), tr, mask=S['m1'], scale=2.0)
        put(self.bn1.weight, dga); put(self.bn1.bias, dbe)
        put(self.linear1.weight, ops.gemm_tf32x3(dy1, S['pooled'], 512, 1024, N, a_mn=True, b_mn=True))
        dpooled = ops.gemm_tf32x3(dy1, self.linear1.weight, N, 1024, 512, b_mn=True)
        # ---- concat stage
        WL = ops._f32c(self.linear[0].weight.reshape(1024, 512))
        L = S['layers']
        x1, x2, x3 = L[1]['x_in'], L[2]['x_in'], L[3]['x_in']
        x4 = ops.pct_residual(x3, L[3]['t'], L[3]['abt'])
        xs4 = [x1, x2, x3, x4]
        _, dga, dbe, (e, f, _) = ops.bn_backward(dpooled, S['zsel'], S['abl'], self.linear[1], S['stl'], cnt, tr, slope=0.2, want_dy=False)
        put(self.linear[1].weight, dga); put(self.linear[1].bias, dbe)
        coef = ops.bn_backward_apply(dpooled, S['zsel'], S['abl'], slope=0.2)            # a_c gy  [N,1024]
        s_cat = torch.cat([ops.col_stats(x.reshape(-1, 128))[:128] for x in xs4])            # sum xcat (f64)
        xbar = (s_cat / cnt).float()
        Wf = ops.axpby_rows(torch.empty_like(WL), 0.0, WL, 1.0, rowscale=f)               # diag(f) WL
        M = ops.gemm_tf32x3(WL, Wf, 512, 512, 1024, a_mn=True, b_mn=True)                  # WL^T diag(f) WL
        e4 = torch.zeros((1024, 4), device=dev, dtype=torch.float32)
        e4[:, 0] = e
        u = ops.gemm_tf32x3(WL, e4, 512, 4, 1024, a_mn=True, b_mn=True)[:, 0].contiguous()  # WL^T e
        gcat = ops.pct_cat_dense_backward(x1, x2, x3, x4, M, u, xbar)
        dWL = torch.zeros_like(WL)
        ops.pct_cat_sparse_backward(coef, S['pstar'], WL, xs4, gcat, dWL)
        G = torch.zeros((512, 512), device=dev, dtype=torch.float32)
        for a in range(4):      # row block a of the Gram matrix: x_a against x_a .. x_4, accumulators resident in tensor memory
            ops.pct_wgrad(xs4[a].reshape(-1, 128), [xs4[b].reshape(-1, 128) for b in range(a, 4)],
                          [G[128 * a:128 * a + 128, 128 * b:128 * b + 128] for b in range(a, 4)])
        for a in range(4):
            for b in range(a + 1, 4):
                G[128 * b:128 * b + 128, 128 * a:128 * a + 128] = G[128 * a:128 * a + 128, 128 * b:128 * b + 128].t()
        Gc = (G.double() - torch.outer(s_cat, s_cat) / cnt).float()                          # centred Gram: sum (x - xbar)(x - xbar)^T
        WG = ops.gemm_tf32x3(WL, Gc, 1024, 512, 512)                                       # WL Gc (symmetric)
        ops.axpby_rows(dWL, 1.0, WG, -1.0, rowscale=f, gamma=-1.0, rowscale2=e, colvec=s_cat)
        put(self.linear[0].weight, dWL)
        del x4, xs4, WG, G, M
        # ---- SA layers 4 .. 1 (pct.py:211-232)
        gx = gcat[3]
        for li in (3, 2, 1, 0):
            sa = (self.sa1, self.sa2, self.sa3, self.sa4)[li]
            Lr = L[li]
            dt, dga, dbe, ex = ops.bn_backward(gx, Lr['t'], Lr['abt'], sa.after_norm, Lr['stt'], cnt, tr, want_absmax=True)
            put(sa.after_norm.weight, dga); put(sa.after_norm.bias, dbe)
            # a bias in front of a batch-statistics BatchNorm has gradient sum_r dt = 0 identically (the reference's autograd
            # returns rounding noise of 1e-7 of the largest gradient there): no pass over dt in train()
            put(sa.trans_conv.bias, torch.zeros(128, device=dev) if tr else colsum(dt, 128))
            Wt = sa.trans_conv.weight.reshape(128, 128)
            dxs, dxs_absmax = ops.pct_pointwise_grad(dt, Wt.t().contiguous(), absmax=ex[3], want_absmax=True)
            dk1, dk2, dv, dv_colsum, dv_absmax = ops.pct_attention_backward(Lr['k'], Lr['v'], Lr['c2'], dxs, dxs_absmax, Lr.get('v_absmax'))
            del dxs
            Wv = sa.v_conv.weight.reshape(128, 128)
            Wk = ops._f32c(sa.k_conv.weight.reshape(32, 128))
            dxv = ops.pct_pointwise_grad(dv, Wv.t().contiguous(), absmax=dv_absmax)
            gx = ops.pct_sa_input_grad(gx, gcat[li - 1] if li > 0 else None, dxv, dk1, dk2, Wk)      # dk1 <- dk
            del dxv, dk2
            put(sa.v_conv.bias, dv_colsum.float() if dv_colsum is not None else colsum(dv, 128))
            dWt = torch.zeros((128, 128), device=dev, dtype=torch.float32)
            dWv = torch.zeros((128, 128), device=dev, dtype=torch.float32)
            dWk = torch.zeros((32, 128), device=dev, dtype=torch.float32)
            xin = Lr['x_in'].reshape(-1, 128)
            ops.pct_wgrad(dt.reshape(-1, 128), [Lr['x_s'].reshape(-1, 128)], [dWt])
            ops.pct_wgrad(xin, [dv.reshape(-1, 128), dk1.reshape(-1, 32)], [dWv, dWk], transpose=True)      # one pass over x_in
            put(sa.trans_conv.weight, dWt); put(sa.v_conv.weight, dWv); put(sa.k_conv.weight, dWk)
            del dt, dv, dk1
        # ---- Embedding (pct.py:120-125): gx = d/d x0, x0 = relu(bn2(conv2(a1))), a1 = relu(bn1(conv1(points)))
        emb = self.embedding
        dz2, dga, dbe, ex = ops.bn_backward(gx, S['z2'], S['ab2'], emb.bn2, S['st2'], cnt, tr, want_absmax=True)
        put(emb.bn2.weight, dga); put(emb.bn2.bias, dbe)
        W1 = ops._f32c(emb.conv1.weight.reshape(128, 3))
        a1 = ops.pct_embed_a1(pts, W1, S['ab1'])
        dW2 = torch.zeros((128, 128), device=dev, dtype=torch.float32)
        ops.pct_wgrad(dz2.reshape(-1, 128), [a1.reshape(-1, 128)], [dW2])
        put(emb.conv2.weight, dW2)
        del a1
        da1 = ops.pct_pointwise_grad(dz2, emb.conv2.weight.reshape(128, 128).t().contiguous(), absmax=ex[3])
        dW1, dga, dbe = ops.pct_embed1_backward(da1, pts, W1, S['ab1'], emb.bn1, S['st1'], S['mom'], cnt, tr)
        put(emb.bn1.weight, dga); put(emb.bn1.bias, dbe)
        put(emb.conv1.weight, dW1)
        return grads
