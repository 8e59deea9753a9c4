"""Serving step captured into one CUDA graph.

The serving hot path (``MultiModalEncoder.forward`` + matching head + anchor positions) is a dozen
short launches around one long one; issued eagerly from Python the GPU idles between them.  For a fixed
batch *layout* (object / edge / anchor counts per pair -- what a bucketed serving loop sees) the whole
step is captured once into a CUDA graph and replayed per batch: inputs are copied into static device
buffers (H2D from pinned host memory or D2D), one graph launch runs every kernel, outputs live in
static buffers.  Nothing here changes what is computed: the captured region is exactly
``trainer.infer_step`` + ``ops.match_anchor_pos``.

Reference call sites this replaces: ``src/inference/sgaligner/inference_align_reg.py:98-145`` (per-batch
model call, per-pair matching loop, metric accumulation).
"""
from __future__ import annotations

from typing import Dict, Optional

import contextlib
import gc

import numpy as np
import torch

from . import matching, ops

_COUNT_KEYS = ('graph_per_obj_count', 'graph_per_edge_count', 'e1i', 'e2i')


@contextlib.contextmanager
def _capture(graph):
    """``torch.cuda.graph`` with the cyclic garbage collector parked: an unreachable CUDA graph / tensor of an earlier instance
    that is collected in the middle of a capture frees device memory, which is not permitted while a stream is capturing and
    invalidates the capture (seen as a sporadic ``cudaErrorStreamCaptureInvalidated``)."""
    gc.collect()
    was = gc.isenabled()
    gc.disable()
    try:
        with torch.cuda.graph(graph):
            yield
    finally:
        if was:
            gc.enable()


class CapturedInference:
    def __init__(self, model, example: Dict, k: int = 6, want_sim: bool = True, graph_branch_sms: int = 16, point_chunks: int = 1):
        """``graph_branch_sms`` > 0: the point encoder (one persistent CTA per SM) leaves that many SMs free and the
        graph branch (CSR build + two GAT layers) is captured on a forked stream, so the two branches of the
        encoder run concurrently inside the graph; 0: one stream, everything back to back.
        ``point_chunks`` > 1: the point encoder is cut into that many launches over object ranges (each a static
        partition over its CTAs): when several steps are in flight on different streams (``PipelinedServing``) the
        shorter launches let the hardware scheduler pack the steps' CTAs tightly instead of in whole-step blocks."""
        pts = example['tot_obj_pts']
        if not (torch.is_tensor(pts) and pts.is_cuda):
            raise RuntimeError('CapturedInference needs a device-resident example batch (no CPU fallback)')
        self.model, self.k, self.want_sim = model, k, want_sim
        self.dev = pts.device
        self.modules = list(model.modules)
        self.side = None
        self.pointnet_ctas = 0
        if graph_branch_sms > 0 and 'gat' in self.modules and 'point' in self.modules:
            # balanced split: every point-encoder CTA gets the same number of objects (+-1)
            sms, n_obj = ops.sm_count(), int(pts.shape[0])
            per = -(-n_obj // max(1, sms - graph_branch_sms))
            self.pointnet_ctas = min(sms - 1, -(-n_obj // per))
            self.side = torch.cuda.Stream(device=self.dev)
        self.point_ranges = None
        if point_chunks > 1 and 'point' in self.modules:
            n_obj = int(pts.shape[0])
            per = -(-n_obj // point_chunks)
            self.point_ranges = [(lo, min(n_obj, lo + per), None) for lo in range(0, n_obj, per)]
            if self.pointnet_ctas:            # the balanced CTA count is per launch now
                sms = ops.sm_count()
                per_cta = -(-per // max(1, sms - graph_branch_sms))
                self.pointnet_ctas = min(sms - 1, -(-per // per_cta))
        # Static input buffers.  Everything a host-fed step copies besides the points (edges, poses, bag-of-words rows, the
        # anchor indices) lives in ONE device arena, so that a serving loop moves it with a single H2D copy from a pinned
        # mirror (``arena_bytes`` / ``arena_slices``) instead of one small copy -- one DMA latency -- per tensor.
        from .data import needed_keys
        need = needed_keys(self.modules)
        self.n_anchor = int(np.asarray(example['e1i']).size)
        e1_host = torch.as_tensor(np.asarray(example['e1i']).astype(np.int32))
        e2_host = torch.as_tensor(np.asarray(example['e2i']).astype(np.int32))
        plan, off = [], 0
        for key, v in example.items():
            if torch.is_tensor(v) and not key.startswith('_sga') and key in need and key != 'tot_obj_pts':
                plan.append((key, off, v))
                off += (v.numel() * v.element_size() + 255) & ~255
        for key, v in (('e1i', e1_host), ('e2i', e2_host)):
            plan.append((key, off, v))
            off += (v.numel() * v.element_size() + 255) & ~255
        self.arena_bytes = off
        self.arena = torch.empty(max(off, 256), dtype=torch.uint8, device=self.dev)
        self.arena_slices = {key: (o, tuple(v.shape), v.dtype) for key, o, v in plan}

        arena_view = self.arena_view
        self.static = {}
        for key, v in example.items():
            if key.startswith('_sga'):
                continue
            if torch.is_tensor(v) and key in self.arena_slices:
                self.static[key] = arena_view(self.arena, key)
                self.static[key].copy_(v)
            else:
                self.static[key] = v.clone() if torch.is_tensor(v) else v
        self.oc = np.asarray(example['graph_per_obj_count']).copy()
        self.ec = np.asarray(example['graph_per_edge_count']).copy()
        self.graph_layout = ops.GraphLayout(self.oc, self.ec, self.dev) if 'gat' in self.modules else None
        self.pair_layout = ops.PairLayout(self.oc, self.dev)
        self.e1 = arena_view(self.arena, 'e1i')
        self.e2 = arena_view(self.arena, 'e2i')
        self.e1.copy_(e1_host)
        self.e2.copy_(e2_host)
        was_training = model.training
        model.eval()
        cur = torch.cuda.current_stream(self.dev)
        side = torch.cuda.Stream(device=self.dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(2):            # warm-up outside the capture (lazy attribute / workspace initialisation)
                self._step()
        cur.wait_stream(side)
        torch.cuda.synchronize(self.dev)
        l0 = ops.LAUNCHES
        self.graph = torch.cuda.CUDAGraph()
        with _capture(self.graph):
            self.out = self._step()
        self.launches_per_replay = ops.LAUNCHES - l0
        if was_training:
            model.train()

    def arena_view(self, buf, key):
        """The tensor ``key`` of the arena layout as a typed view into ``buf`` (the device arena or a pinned mirror).  A method,
        not a stored closure: a reference cycle through ``self`` would leave dropped instances (and their CUDA graphs) to the
        cyclic collector, which may then free them in the middle of a later capture and invalidate it."""
        o, shape, dtype = self.arena_slices[key]
        n = int(np.prod(shape)) if len(shape) else 1
        return buf[o:o + n * torch.empty((), dtype=dtype).element_size()].view(dtype).view(shape)

    def _step(self):
        with torch.no_grad():
            d = dict(self.static)
            if self.graph_layout is not None:
                d['_sga_graph_layout'] = self.graph_layout
            if self.point_ranges is not None:
                d['_sga_point_ranges'] = self.point_ranges
            if self.side is not None:
                d['_sga_side_stream'] = self.side
                ops.pointnet_set_max_ctas(self.pointnet_ctas)
            try:
                out = self.model(d)
            finally:
                if self.side is not None:
                    ops.pointnet_set_max_ctas(0)
            emb = out['joint'] if len(self.modules) > 1 else out[self.modules[0]]
            res = matching.match_batch(emb, d, k=self.k, full_rank=False, want_sim=True, layout=self.pair_layout)
            pos = ops.match_anchor_pos(res['sim'], self.pair_layout, self.e1, self.e2) if self.n_anchor else None
        return {'embeddings': out, 'topk_idx': res['topk_idx'], 'topk_dist': res['topk_dist'], 'sim': res['sim'],
                'anchor_pos': pos, 'layout': self.pair_layout}

    def check_layout(self, batch: Dict):
        if not (np.array_equal(np.asarray(batch['graph_per_obj_count']), self.oc)
                and np.array_equal(np.asarray(batch['graph_per_edge_count']), self.ec)
                and int(np.asarray(batch['e1i']).size) == self.n_anchor):
            raise ValueError('batch layout differs from the captured one: capture a new CapturedInference for it')

    def load(self, batch: Dict, e1i: Optional[torch.Tensor] = None, e2i: Optional[torch.Tensor] = None):
        """Copy a batch with the captured layout into the static buffers (async on the current stream;
        pinned host tensors give true asynchronous H2D copies).  ``e1i`` / ``e2i``: int32 tensors (host,
        pinned, or device); default: taken from the batch dict."""
        self.check_layout(batch)
        for key, dst in self.static.items():
            if torch.is_tensor(dst):
                dst.copy_(batch[key], non_blocking=True)
        if self.n_anchor:
            self.e1.copy_(e1i if e1i is not None else torch.as_tensor(np.asarray(batch['e1i']).astype(np.int32)), non_blocking=True)
            self.e2.copy_(e2i if e2i is not None else torch.as_tensor(np.asarray(batch['e2i']).astype(np.int32)), non_blocking=True)

    def replay(self) -> Dict:
        self.graph.replay()
        return self.out

    # ------------------------------------------------------------------ host-to-host step in one graph
    def capture_host_step(self, host_batch: Dict, n_chunks: int = 8, taper: bool = False):
        """Second graph for batches that start in HOST memory: pinned staging buffers (returned; a loader collates
        into them in place) -> H2D copies on a forked copy stream (small tensors first, then the points in
        ``n_chunks`` object ranges) -> graph branch as soon as the small tensors have landed, point encoder chunk by
        chunk as the copies land -> matching head -> D2H of top-k / anchor positions into pinned result buffers.
        Only the tensors the configured modalities read are staged (``data.needed_keys``)."""
        from .data import needed_keys
        self.check_layout(host_batch)
        keys = [k for k in needed_keys(self.modules) if k in self.static and torch.is_tensor(self.static[k])]
        self.host_in = {k: torch.empty(self.static[k].shape, dtype=self.static[k].dtype).pin_memory() for k in keys}
        self.host_e1 = torch.empty(self.n_anchor, dtype=torch.int32).pin_memory()
        self.host_e2 = torch.empty(self.n_anchor, dtype=torch.int32).pin_memory()
        self.fill_host(host_batch)
        N = int(self.static['tot_obj_pts'].shape[0])
        if taper and n_chunks > 1:
            # the step is PCIe bound: the encoder always waits for the copy, so what the step pays after the last
            # copy is the encoder time of the LAST range -- make the ranges shrink (n : n-1 : ... : 1)
            w = np.arange(n_chunks, 0, -1, dtype=np.float64)
            cuts = np.concatenate([[0], np.round(np.cumsum(w) / w.sum() * N)]).astype(np.int64)
            ranges = [(int(a), int(b)) for a, b in zip(cuts[:-1], cuts[1:]) if b > a]
        else:
            per = -(-N // max(1, n_chunks))
            ranges = [(s, min(N, s + per)) for s in range(0, N, per)]
        cs = torch.cuda.Stream(device=self.dev)
        self.host_out = {}

        def body():
            cur = torch.cuda.current_stream(self.dev)
            cs.wait_stream(cur)
            with torch.cuda.stream(cs):
                for k in keys:
                    if k != 'tot_obj_pts':
                        self.static[k].copy_(self.host_in[k], non_blocking=True)
                if self.n_anchor:
                    self.e1.copy_(self.host_e1, non_blocking=True)
                    self.e2.copy_(self.host_e2, non_blocking=True)
                ev_small = torch.cuda.Event()
                ev_small.record(cs)
                chunks = []
                if 'tot_obj_pts' in self.host_in:
                    for (a, b) in ranges:
                        self.static['tot_obj_pts'][a:b].copy_(self.host_in['tot_obj_pts'][a:b], non_blocking=True)
                        ev = torch.cuda.Event()
                        ev.record(cs)
                        chunks.append((a, b, ev))
            with torch.no_grad():
                d = dict(self.static)
                if self.graph_layout is not None:
                    d['_sga_graph_layout'] = self.graph_layout
                d['_sga_ready'] = {'small': ev_small, 'pts': chunks, 'stream': cs}
                out = self.model(d)
                emb = out['joint'] if len(self.modules) > 1 else out[self.modules[0]]
                res = matching.match_batch(emb, d, k=self.k, full_rank=False, want_sim=True, layout=self.pair_layout)
                pos = ops.match_anchor_pos(res['sim'], self.pair_layout, self.e1, self.e2) if self.n_anchor else None
                if 'topk_idx' not in self.host_out:
                    self.host_out['topk_idx'] = torch.empty(res['topk_idx'].shape, dtype=torch.int32).pin_memory()
                    if pos is not None:
                        self.host_out['anchor_pos'] = torch.empty(pos.shape, dtype=torch.int32).pin_memory()
                self.host_out['topk_idx'].copy_(res['topk_idx'], non_blocking=True)
                if pos is not None:
                    self.host_out['anchor_pos'].copy_(pos, non_blocking=True)
            cur.wait_stream(cs)
            return {'embeddings': out, 'topk_idx': res['topk_idx'], 'anchor_pos': pos, 'sim': res['sim']}

        was_training = self.model.training
        self.model.eval()
        cur = torch.cuda.current_stream(self.dev)
        side = torch.cuda.Stream(device=self.dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            body()
        cur.wait_stream(side)
        torch.cuda.synchronize(self.dev)
        self.graph_host = torch.cuda.CUDAGraph()
        with _capture(self.graph_host):
            self.out_host = body()
        if was_training:
            self.model.train()
        return self.host_in

    # ------------------------------------------------------------------ host-to-host step, copies outside the graph
    def capture_host_hybrid(self, host_batch: Dict, n_chunks: int = 4):
        """Third form of the host-to-host step.  H2D memcpy NODES inside a CUDA graph showed box-dependent gaps (the
        same captured step took 0.71-1.17 ms on different hosts while stream-ordered copies were stable), so here
        the copies are issued eagerly on the copy stream (they pipeline back to back on the DMA engine), the point
        encoder is launched range by range as they land, and only the compute around it is replayed from two small
        graphs: the graph branch (CSR + GAT, needs the small tensors only) and the tail (projection / fusion,
        matching head, anchor positions, D2H of the results).  Host cost: ~10 copy calls, ``n_chunks`` launches, two
        graph launches -- hidden under the copy itself."""
        from . import autograd as ag
        from .data import needed_keys
        if not ('point' in self.modules and len(self.modules) > 1):
            raise NotImplementedError('hybrid host step: needs the point modality plus at least one more')
        self.check_layout(host_batch)
        m = self.model
        keys = [k for k in needed_keys(self.modules) if k in self.static and torch.is_tensor(self.static[k])]
        if not hasattr(self, 'host_in'):
            self.host_in = {k: torch.empty(self.static[k].shape, dtype=self.static[k].dtype).pin_memory() for k in keys}
            self.host_e1 = torch.empty(self.n_anchor, dtype=torch.int32).pin_memory()
            self.host_e2 = torch.empty(self.n_anchor, dtype=torch.int32).pin_memory()
        self.fill_host(host_batch)
        N = int(self.static['tot_obj_pts'].shape[0])
        per = -(-N // max(1, n_chunks))
        self.hy_ranges = [(s_, min(N, s_ + per)) for s_ in range(0, N, per)]
        self.hy_keys = keys
        self.hy_cs = torch.cuda.Stream(device=self.dev)
        self.hy_side = torch.cuda.Stream(device=self.dev)
        self.hy_px = torch.empty((N, m.object_encoder.out_size), device=self.dev, dtype=torch.float32)
        self.hy_events = [torch.cuda.Event() for _ in self.hy_ranges]
        self.hy_ev_small = torch.cuda.Event()
        self.hy_ev_gat = torch.cuda.Event()
        was_training = m.training
        m.eval()
        d = self.static
        mods = self.modules

        def graph_branch():
            with torch.no_grad():
                graph = ops.BatchGraph(d['edges'], self.oc, self.ec, layout=self.graph_layout)
                return m.structure_encoder(d['tot_rel_pose'], graph)

        def tail(gat_out):
            with torch.no_grad():
                args = []
                for module in mods:
                    if module == 'gat':
                        args += [gat_out, m.structure_embedding.weight, m.structure_embedding.bias]
                    elif module == 'point':
                        args += [self.hy_px, m.object_embedding.weight, m.object_embedding.bias]
                    elif module == 'rel':
                        args += [d['tot_bow_vec_object_edge_feats'], m.meta_embedding_rel.weight, m.meta_embedding_rel.bias]
                    elif module == 'attr':
                        args += [d['tot_bow_vec_object_attr_feats'], m.meta_embedding_attr.weight, m.meta_embedding_attr.bias]
                    else:
                        raise NotImplementedError
                outs = ag.ProjectFuse.apply(m.fusion.weight, len(mods), *args)
                t_idx, t_dist, sim = ops.match_topk_tc(outs[len(mods)], self.pair_layout, self.k, True)
                pos = ops.match_anchor_pos(sim, self.pair_layout, self.e1, self.e2) if self.n_anchor else None
                if not hasattr(self, 'hy_out'):
                    self.hy_out = {'topk_idx': torch.empty(t_idx.shape, dtype=torch.int32).pin_memory()}
                    if pos is not None:
                        self.hy_out['anchor_pos'] = torch.empty(pos.shape, dtype=torch.int32).pin_memory()
                self.hy_out['topk_idx'].copy_(t_idx, non_blocking=True)
                if pos is not None:
                    self.hy_out['anchor_pos'].copy_(pos, non_blocking=True)
                return outs, t_idx, pos, sim

        has_gat = 'gat' in mods
        cur = torch.cuda.current_stream(self.dev)
        warm = torch.cuda.Stream(device=self.dev)
        warm.wait_stream(cur)
        with torch.cuda.stream(warm):
            g0 = graph_branch() if has_gat else None
            tail(g0)
        cur.wait_stream(warm)
        torch.cuda.synchronize(self.dev)
        self.hy_gat = None
        if has_gat:
            self.hy_graph_a = torch.cuda.CUDAGraph()
            with _capture(self.hy_graph_a):
                self.hy_gat = graph_branch()
        self.hy_graph_b = torch.cuda.CUDAGraph()
        with _capture(self.hy_graph_b):
            self.hy_tail_out = tail(self.hy_gat)
        if was_training:
            m.train()
        return self.host_in

    def run_host_hybrid(self) -> Dict:
        """Eager stream-ordered H2D copies + range-by-range point encoder + two graph replays; blocks until the pinned
        results are valid."""
        m = self.model
        cur = torch.cuda.current_stream(self.dev)
        cs = self.hy_cs
        cs.wait_stream(cur)
        with torch.cuda.stream(cs):
            for k in self.hy_keys:
                if k != 'tot_obj_pts':
                    self.static[k].copy_(self.host_in[k], non_blocking=True)
            if self.n_anchor:
                self.e1.copy_(self.host_e1, non_blocking=True)
                self.e2.copy_(self.host_e2, non_blocking=True)
            self.hy_ev_small.record(cs)
            chunks = []
            for (a, b), ev in zip(self.hy_ranges, self.hy_events):
                self.static['tot_obj_pts'][a:b].copy_(self.host_in['tot_obj_pts'][a:b], non_blocking=True)
                ev.record(cs)
                chunks.append((a, b, ev))
        if self.hy_gat is not None:
            self.hy_side.wait_stream(cur)
            self.hy_side.wait_event(self.hy_ev_small)
            with torch.cuda.stream(self.hy_side):
                self.hy_graph_a.replay()
                self.hy_ev_gat.record(self.hy_side)
        enc = m.object_encoder
        ops.pointnet_set_max_ctas(self.pointnet_ctas)
        try:
            with torch.no_grad():
                ops.pointnet_forward(self.static['tot_obj_pts'], enc.conv1.weight, enc.conv1.bias, enc.conv2.weight, enc.conv2.bias,
                                     enc.conv3.weight, enc.conv3.bias, want_argmax=False, mode=enc.kernel_mode, chunks=chunks,
                                     out=self.hy_px)
        finally:
            ops.pointnet_set_max_ctas(0)
        cur.wait_event(self.hy_ev_small)
        if self.hy_gat is not None:
            cur.wait_event(self.hy_ev_gat)
        self.hy_graph_b.replay()
        cur.synchronize()
        return self.hy_out

    def fill_host(self, host_batch: Dict):
        """Slow path: copy a host batch into the pinned staging buffers (a production loader collates into
        ``host_in`` directly)."""
        self.check_layout(host_batch)
        for k, dst in self.host_in.items():
            dst.copy_(host_batch[k])
        if self.n_anchor:
            self.host_e1.copy_(torch.as_tensor(np.asarray(host_batch['e1i']).astype(np.int32)))
            self.host_e2.copy_(torch.as_tensor(np.asarray(host_batch['e2i']).astype(np.int32)))

    def run_host(self) -> Dict:
        """One replay of the host-to-host graph on the staging buffers' current content; blocks until the pinned
        results (``topk_idx``, ``anchor_pos``) are valid."""
        self.graph_host.replay()
        torch.cuda.current_stream(self.dev).synchronize()
        return self.host_out

    def __call__(self, batch: Dict, **kw) -> Dict:
        self.load(batch, **kw)
        return self.replay()


class PipelinedServing:
    """Steady-state host-to-host serving: step k+1's H2D copy runs under step k's compute.

    ``CapturedInference.run_host`` is ONE synchronous step (copy -> compute -> copy back -> host sync): its time is
    the SUM of the PCIe time and the compute time.  A serving loop does not need that: with ``n_slots`` >= 2
    independent slots (static device buffers + captured graph + pinned staging / result buffers each) the copy
    engine fills slot k+1 while the SMs work on slot k, and throughput becomes 1 / max(copy, compute).

        pipe = PipelinedServing(model, example_device_batch, n_slots=3)
        pipe.staging(s)          # pinned host tensors of slot s: the loader collates INTO them
        pipe.submit(s)           # enqueue H2D -> graph -> D2H for slot s; returns at once
        pipe.wait(s)             # block until slot s's pinned results are valid, return them

    Every slot's batch must have the captured layout (see ``LayoutCache`` for ragged streams).  Reference loop this
    replaces: ``inference_align_reg.py:98-145`` (synchronous per-batch ``to_cuda`` + model call + per-pair matching).
    """

    def __init__(self, model, example: Dict, k: int = 6, n_slots: int = 2, compute_streams: int = 2, **capture_kw):
        """``compute_streams`` > 1: consecutive steps are replayed on alternating streams, so the short tail of step k
        (projection, matching head, anchor positions: latency-bound launches that leave most SMs idle) runs on the SMs
        the point encoder of step k+1 does not occupy, and that encoder's CTAs start as soon as the previous one's
        retire -- steady-state throughput approaches the SM time of the point encoder instead of the latency of a step."""
        from .data import needed_keys
        assert n_slots >= 2 and compute_streams >= 1
        self.dev = example['tot_obj_pts'].device
        self.slots = [CapturedInference(model, example, k=k, **capture_kw) for _ in range(n_slots)]
        self.copy_stream = torch.cuda.Stream(device=self.dev)
        self.cstreams = [torch.cuda.Stream(device=self.dev) for _ in range(compute_streams)]
        self.n_submitted = 0
        self.d2h_stream = torch.cuda.Stream(device=self.dev)      # results leave on their own stream: the next graph replay does not queue behind them
        c0 = self.slots[0]
        keys = [k_ for k_ in needed_keys(c0.modules) if k_ in c0.static and torch.is_tensor(c0.static[k_])]
        self.keys = keys
        self.small_stream = torch.cuda.Stream(device=self.dev)     # the arena copy runs beside the points copy, not behind it
        # On some hosts one DMA stream tops out near 35 GB/s where two concurrent copies reach the 54 GB/s of the link
        # (tools/h2d_split_probe.py); where one stream already does, splitting only adds events and jitter (measured: 0.486 ->
        # 0.52-0.62 ms per step).  So the split is chosen from a 2 ms measurement on slot 0's own buffers (below).
        self.copy_split = 1
        self.extra_copy_streams = [torch.cuda.Stream(device=self.dev)]
        for c in self.slots:
            # pinned mirror of the slot's device arena: the small tensors and the anchor indices are views into it
            c.p_arena = torch.empty(max(c.arena_bytes, 256), dtype=torch.uint8).pin_memory()
            c.p_in = {k_: (c.arena_view(c.p_arena, k_) if k_ in c.arena_slices
                           else torch.empty(c.static[k_].shape, dtype=c.static[k_].dtype).pin_memory()) for k_ in keys}
            c.p_e1 = c.arena_view(c.p_arena, 'e1i')
            c.p_e2 = c.arena_view(c.p_arena, 'e2i')
            c.ev_small = torch.cuda.Event()
            c.ev_part = [torch.cuda.Event() for _ in self.extra_copy_streams]
            c.p_out = {'topk_idx': torch.empty(c.out['topk_idx'].shape, dtype=torch.int32).pin_memory()}
            if c.out['anchor_pos'] is not None:
                c.p_out['anchor_pos'] = torch.empty(c.out['anchor_pos'].shape, dtype=torch.int32).pin_memory()
            c.ev_in = torch.cuda.Event()
            c.ev_graph = torch.cuda.Event()
            c.ev_done = torch.cuda.Event()
            c.in_flight = False
        self.big_keys = [k_ for k_ in keys if k_ not in c0.arena_slices]
        self.copy_split = self._pick_copy_split(c0)
        self.h2d_bytes = sum(c0.p_in[k_].numel() * c0.p_in[k_].element_size() for k_ in self.big_keys) + c0.arena_bytes
        self.d2h_bytes = sum(t.numel() * t.element_size() for t in c0.p_out.values())

    def staging(self, slot: int) -> Dict:
        c = self.slots[slot]
        return {**c.p_in, 'e1i': c.p_e1, 'e2i': c.p_e2}

    def fill(self, slot: int, host_batch: Dict):
        """Slow path for callers whose batch is not already in the staging buffers (a loader collates into them)."""
        c = self.slots[slot]
        c.check_layout(host_batch)
        for k_, dst in c.p_in.items():
            dst.copy_(host_batch[k_])
        if c.n_anchor:
            c.p_e1.copy_(torch.as_tensor(np.asarray(host_batch['e1i']).astype(np.int32)))
            c.p_e2.copy_(torch.as_tensor(np.asarray(host_batch['e2i']).astype(np.int32)))

    def submit(self, slot: int):
        c = self.slots[slot]
        cur = torch.cuda.current_stream(self.dev)
        cs = self.copy_stream
        if c.in_flight:
            cs.wait_event(c.ev_done)        # the slot's previous compute has read its static buffers
            self.small_stream.wait_event(c.ev_done)
            for xs in self.extra_copy_streams:
                xs.wait_event(c.ev_done)
        self.h2d(c)
        st = self.cstreams[self.n_submitted % len(self.cstreams)]
        self.n_submitted += 1
        st.wait_stream(cur)                 # whatever the caller enqueued before this submit
        st.wait_event(c.ev_in)
        if c.in_flight:
            st.wait_event(c.ev_done)        # the slot's previous results have left its output buffers
        with torch.cuda.stream(st):
            c.graph.replay()
            c.ev_graph.record(st)
        ds = self.d2h_stream
        ds.wait_event(c.ev_graph)
        with torch.cuda.stream(ds):
            c.p_out['topk_idx'].copy_(c.out['topk_idx'], non_blocking=True)
            if 'anchor_pos' in c.p_out:
                c.p_out['anchor_pos'].copy_(c.out['anchor_pos'], non_blocking=True)
            c.ev_done.record(ds)
        c.in_flight = True

    def h2d(self, c):
        """One step's host-to-device traffic: the points (and any other large tensor) on the copy stream, the slot's arena
        (every small tensor + the anchor indices) as ONE copy on a second stream; ``c.ev_in`` fires when both have landed."""
        with torch.cuda.stream(self.small_stream):
            if c.arena_bytes:
                c.arena.copy_(c.p_arena, non_blocking=True)
            c.ev_small.record(self.small_stream)
        cs = self.copy_stream
        streams = ([cs] + self.extra_copy_streams)[:self.copy_split]
        for i, st in enumerate(streams):
            with torch.cuda.stream(st):
                for k_ in self.big_keys:
                    dst, src = c.static[k_], c.p_in[k_]
                    rows = dst.shape[0]
                    per = -(-rows // len(streams))
                    lo, hi = min(rows, i * per), min(rows, (i + 1) * per)
                    if hi > lo:
                        dst[lo:hi].copy_(src[lo:hi], non_blocking=True)
                if i > 0:
                    c.ev_part[i - 1].record(st)
        with torch.cuda.stream(cs):
            for ev in c.ev_part[:self.copy_split - 1]:
                cs.wait_event(ev)
            cs.wait_event(c.ev_small)
            c.ev_in.record(cs)

    def _pick_copy_split(self, c) -> int:
        """1 or 2 concurrent H2D streams for the large tensors: whichever moves slot 0's staging buffers faster (best of 3
        passes each; 2 only if it wins by more than 10 %)."""
        if not self.big_keys:
            return 1
        # the measurement copies staging -> static buffers: make staging a mirror of what the slot holds first, so the slot
        # still contains the example batch afterwards
        if c.arena_bytes:
            c.p_arena.copy_(c.arena)
        for k_ in self.big_keys:
            c.p_in[k_].copy_(c.static[k_])
        torch.cuda.synchronize(self.dev)
        best = {}
        for split in (1, 2):
            self.copy_split = split
            ts = []
            for _ in range(4):
                t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                cur = torch.cuda.current_stream(self.dev)
                t0.record(cur)
                for st in [self.copy_stream, self.small_stream] + self.extra_copy_streams:
                    st.wait_event(t0)
                self.h2d(c)
                cur.wait_event(c.ev_in)
                t1.record(cur)
                t1.synchronize()
                ts.append(t0.elapsed_time(t1))
            best[split] = min(ts[1:])
        return 2 if best[2] < 0.9 * best[1] else 1

    def wait(self, slot: int) -> Dict:
        c = self.slots[slot]
        if c.in_flight:
            c.ev_done.synchronize()
        return c.p_out

    # ---- device-resident variant (inputs already in the slots' static buffers, results stay on the device)
    def load_resident(self, slot: int, device_batch: Dict):
        self.slots[slot].load(device_batch)

    def submit_resident(self, slot: int):
        """Replay slot ``slot`` on the next compute stream; no host traffic.  ``sync_resident`` joins the streams."""
        c = self.slots[slot]
        st = self.cstreams[self.n_submitted % len(self.cstreams)]
        self.n_submitted += 1
        if getattr(c, 'resident_busy', False):
            st.wait_event(c.ev_graph)       # a slot never overlaps with itself (static output buffers)
        with torch.cuda.stream(st):
            c.graph.replay()
            c.ev_graph.record(st)
        c.resident_busy = True
        return c.out

    def fork_resident(self, stream=None):
        """Make every compute stream wait for the work already enqueued on ``stream`` (default: the current one)."""
        cur = stream or torch.cuda.current_stream(self.dev)
        for st in self.cstreams:
            st.wait_stream(cur)

    def sync_resident(self, stream=None):
        """Make ``stream`` (default: the current one) wait for every compute stream."""
        cur = stream or torch.cuda.current_stream(self.dev)
        for st in self.cstreams:
            cur.wait_stream(st)


class LayoutCache:
    """Ragged serving: real 3RScan batches differ in object / edge / anchor counts, and a captured graph is frozen to
    one layout.  This cache keeps one :class:`CapturedInference` per layout it has seen (LRU, ``capacity`` entries)
    and falls back to eager launches for a layout the first time it appears (the capture costs a few steps; a layout
    is captured once it has been seen ``capture_after`` times).  Results are identical either way (same kernels)."""

    def __init__(self, model, k: int = 6, capacity: int = 16, capture_after: int = 2):
        from collections import OrderedDict
        self.model, self.k = model, k
        self.capacity, self.capture_after = capacity, capture_after
        self.graphs = OrderedDict()
        self.seen = {}
        self.hits = self.misses = 0

    @staticmethod
    def layout_key(batch: Dict):
        return (np.asarray(batch['graph_per_obj_count']).astype(np.int64).tobytes(),
                np.asarray(batch['graph_per_edge_count']).astype(np.int64).tobytes(), int(np.asarray(batch['e1i']).size),
                tuple(batch['tot_obj_pts'].shape))

    def __call__(self, batch: Dict) -> Dict:
        """``batch``: device-resident collated batch.  Returns the dict ``CapturedInference.replay`` returns."""
        key = self.layout_key(batch)
        cap = self.graphs.get(key)
        if cap is not None:
            self.graphs.move_to_end(key)
            self.hits += 1
            return cap(batch)
        self.misses += 1
        n = self.seen.get(key, 0) + 1
        self.seen[key] = n
        if n >= self.capture_after:
            cap = CapturedInference(self.model, batch, k=self.k)
            self.graphs[key] = cap
            if len(self.graphs) > self.capacity:
                self.graphs.popitem(last=False)
            return cap(batch)
        return self._eager(batch)

    def _eager(self, batch: Dict) -> Dict:
        mods = list(self.model.modules)
        lay = ops.PairLayout(np.asarray(batch['graph_per_obj_count']), batch['tot_obj_pts'].device)
        with torch.no_grad():
            out = self.model(batch)
            emb = out['joint'] if len(mods) > 1 else out[mods[0]]
            res = matching.match_batch(emb, batch, k=self.k, full_rank=False, want_sim=True, layout=lay)
            na = int(np.asarray(batch['e1i']).size)
            dev = emb.device
            pos = None
            if na:
                e1 = torch.as_tensor(np.asarray(batch['e1i']).astype(np.int32)).to(dev, non_blocking=True)
                e2 = torch.as_tensor(np.asarray(batch['e2i']).astype(np.int32)).to(dev, non_blocking=True)
                pos = ops.match_anchor_pos(res['sim'], lay, e1, e2)
        return {'embeddings': out, 'topk_idx': res['topk_idx'], 'topk_dist': res['topk_dist'], 'sim': res['sim'],
                'anchor_pos': pos, 'layout': lay}
