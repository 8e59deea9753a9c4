"""CUDA-graph captured serving step: identical results to the eager path, replayable on new batches of
the same layout, rejects a different layout."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available()
    return torch.device('cuda:0')


def _eager(model, data, k):
    from sgaligner_b200 import matching, ops
    with torch.no_grad():
        out = model(data)
        res = matching.match_batch(out['joint'], data, k=k, full_rank=False)
        e1 = torch.as_tensor(np.asarray(data['e1i']).astype(np.int32)).to(out['joint'].device)
        e2 = torch.as_tensor(np.asarray(data['e2i']).astype(np.int32)).to(out['joint'].device)
        pos = ops.match_anchor_pos(res['sim'], res['layout'], e1, e2)
    return out, res, pos


def test_captured_inference_matches_eager_and_replays(dev):
    from sgaligner_b200 import synthetic, to_cuda
    from sgaligner_b200.data import pin
    from sgaligner_b200.serving import CapturedInference
    from sgaligner_b200.sg_aligner import MultiModalEncoder
    torch.manual_seed(0)
    modules = ['point', 'gat', 'rel', 'attr']
    model = MultiModalEncoder(modules=modules, rel_dim=41, attr_dim=164).to(dev).eval()
    ns, nr, na = [9, 12, 7, 30], [11, 8, 10, 25], [5, 6, 4, 12]
    host_a = synthetic.make_batch(ns, nr, na, n_points=256, edge_mode='complete', seed=5)
    host_b = synthetic.make_batch(ns, nr, na, n_points=256, edge_mode='complete', seed=6)
    a, b = to_cuda(dict(host_a), dev), to_cuda(dict(host_b), dev)
    cap = CapturedInference(model, a, k=6)
    assert cap.launches_per_replay > 0
    for host, d in ((host_a, a), (host_b, b), (host_a, a)):
        got = cap(pin(host))                 # H2D from pinned host memory into the static buffers + one replay
        torch.cuda.synchronize()
        out, res, pos = _eager(model, d, 6)
        torch.cuda.synchronize()
        for key in out:
            assert torch.equal(got['embeddings'][key], out[key]), key
        assert torch.equal(got['topk_idx'], res['topk_idx'])
        assert torch.equal(got['sim'], res['sim'])
        assert torch.equal(got['anchor_pos'], pos)
    other = to_cuda(dict(synthetic.make_batch([9, 12, 7, 31], nr, na, n_points=256, edge_mode='complete', seed=5)), dev)
    with pytest.raises(ValueError):
        cap(other)


def test_captured_host_to_host_step(dev):
    """H2D staging + encoder + matching + D2H in one graph: same top-k / anchor positions as the eager path, for
    the captured batch and for another batch of the same layout written into the staging buffers."""
    from sgaligner_b200 import synthetic, to_cuda
    from sgaligner_b200.serving import CapturedInference
    from sgaligner_b200.sg_aligner import MultiModalEncoder
    torch.manual_seed(0)
    for modules in (['point', 'gat'], ['point', 'gat', 'rel', 'attr']):
        model = MultiModalEncoder(modules=modules, rel_dim=41, attr_dim=164).to(dev).eval()
        ns, nr, na = [9, 12, 7, 30], [11, 8, 10, 25], [5, 6, 4, 12]
        host_a = synthetic.make_batch(ns, nr, na, n_points=256, edge_mode='complete', seed=5)
        host_b = synthetic.make_batch(ns, nr, na, n_points=256, edge_mode='complete', seed=6)
        cap = CapturedInference(model, to_cuda(dict(host_a), dev), k=6)
        cap.capture_host_step(host_a, n_chunks=3)
        for host in (host_a, host_b, host_a):
            cap.fill_host(host)
            got = cap.run_host()
            out, res, pos = _eager(model, to_cuda(dict(host), dev), 6)
            torch.cuda.synchronize()
            assert torch.equal(got['topk_idx'], res['topk_idx'].cpu())
            assert torch.equal(got['anchor_pos'], pos.cpu())
        # the hybrid form (stream-ordered copies + range-by-range point encoder + two small graphs)
        cap.capture_host_hybrid(host_a, n_chunks=3)
        for host in (host_b, host_a, host_b):
            cap.fill_host(host)
            got = cap.run_host_hybrid()
            out, res, pos = _eager(model, to_cuda(dict(host), dev), 6)
            torch.cuda.synchronize()
            assert torch.equal(got['topk_idx'], res['topk_idx'].cpu())
            assert torch.equal(got['anchor_pos'], pos.cpu())


def test_pipelined_serving_overlaps_slots_and_matches_eager(dev):
    """Steady-state loop: three slots in flight with DIFFERENT batches, nothing synchronised between submissions;
    every slot's pinned results equal the eager step on its batch, also after the slots have been reused."""
    from sgaligner_b200 import synthetic, to_cuda
    from sgaligner_b200.serving import PipelinedServing
    from sgaligner_b200.sg_aligner import MultiModalEncoder
    torch.manual_seed(0)
    modules = ['point', 'gat']
    model = MultiModalEncoder(modules=modules, rel_dim=41, attr_dim=164).to(dev).eval()
    ns, nr, na = [9, 12, 7, 30], [11, 8, 10, 25], [5, 6, 4, 12]
    hosts = [synthetic.make_batch(ns, nr, na, n_points=256, edge_mode='complete', seed=30 + i) for i in range(5)]
    want = []
    for h in hosts:
        _, res, pos = _eager(model, to_cuda(dict(h), dev), 6)
        want.append((res['topk_idx'].cpu(), pos.cpu()))
    pipe = PipelinedServing(model, to_cuda(dict(hosts[0]), dev), k=6, n_slots=3)
    assert pipe.h2d_bytes > 0 and pipe.d2h_bytes > 0 and pipe.copy_split in (1, 2)
    # every slot (also slot 0, whose buffers the copy-stream measurement of the constructor went through) still holds the
    # example batch, and the staging of slot 0 mirrors it: submitting it untouched serves the example
    pipe.submit(0)
    got = pipe.wait(0)
    assert torch.equal(got['topk_idx'], want[0][0]) and torch.equal(got['anchor_pos'], want[0][1])
    # the small tensors and the anchor indices are views of ONE pinned arena (a single H2D copy per step)
    st = pipe.staging(0)
    assert st['e1i'].is_pinned() and st['e1i'].untyped_storage().data_ptr() == pipe.slots[0].p_arena.untyped_storage().data_ptr()
    order = [0, 1, 2, 3, 4, 2, 0, 4, 1, 3, 3, 0]
    pending = {}
    for step, b in enumerate(order):
        s = step % 3
        if s in pending:
            got = pipe.wait(s)
            tk, pos = want[pending.pop(s)]
            assert torch.equal(got['topk_idx'], tk) and torch.equal(got['anchor_pos'], pos)
        pipe.fill(s, hosts[b])
        pipe.submit(s)
        pending[s] = b
    for s, b in pending.items():
        got = pipe.wait(s)
        assert torch.equal(got['topk_idx'], want[b][0]) and torch.equal(got['anchor_pos'], want[b][1])


def test_resident_two_stream_loop_matches_eager(dev):
    """Device-resident serving loop on two alternating streams: 4 slots with different batches, 11 back-to-back
    replays without any synchronisation in between; every slot ends with the eager results of its batch."""
    from sgaligner_b200 import synthetic, to_cuda
    from sgaligner_b200.serving import PipelinedServing
    from sgaligner_b200.sg_aligner import MultiModalEncoder
    torch.manual_seed(0)
    model = MultiModalEncoder(modules=['point', 'gat'], rel_dim=41, attr_dim=164).to(dev).eval()
    ns, nr, na = [40, 12, 64, 30], [33, 8, 64, 25], [5, 6, 30, 12]
    hosts = [synthetic.make_batch(ns, nr, na, n_points=512, edge_mode='kout', k_out=5, seed=60 + i) for i in range(4)]
    devs = [to_cuda(dict(h), dev) for h in hosts]
    want = []
    for d in devs:
        _, res, pos = _eager(model, d, 6)
        want.append((res['topk_idx'].clone(), pos.clone()))
    pipe = PipelinedServing(model, devs[0], k=6, n_slots=4, compute_streams=2)
    for s in range(4):
        pipe.load_resident(s, devs[s])
    pipe.fork_resident()
    for step in range(11):
        pipe.submit_resident(step % 4)
    pipe.sync_resident()
    torch.cuda.synchronize()
    for s in range(4):
        out = pipe.slots[s].out
        assert torch.equal(out['topk_idx'], want[s][0]) and torch.equal(out['anchor_pos'], want[s][1]), s


def test_layout_cache_serves_ragged_batches(dev):
    """Ragged stream: batches of three different layouts interleaved; a layout is served eagerly at first, from its
    own captured graph from the second time on -- identical results throughout."""
    from sgaligner_b200 import synthetic, to_cuda
    from sgaligner_b200.serving import LayoutCache
    from sgaligner_b200.sg_aligner import MultiModalEncoder
    torch.manual_seed(0)
    model = MultiModalEncoder(modules=['point', 'gat', 'rel', 'attr'], rel_dim=41, attr_dim=164).to(dev).eval()
    layouts = [([9, 12], [11, 8], [5, 6]), ([20, 7, 14], [18, 9, 11], [9, 3, 7]), ([33], [29], [15]), ([8, 8], [9, 9], [4, 4])]
    cache = LayoutCache(model, k=6, capacity=3, capture_after=2)
    order = [0, 1, 2] * 4 + [3, 3, 3, 0, 1, 2]          # three layouts in rotation, then a fourth evicts the oldest
    for step, li_ in enumerate(order):
        ns, nr, na = layouts[li_]
        d = to_cuda(dict(synthetic.make_batch(ns, nr, na, n_points=128, edge_mode='complete', seed=50 + step)), dev)
        got = cache(d)
        torch.cuda.synchronize()
        _, res, pos = _eager(model, d, 6)
        assert torch.equal(got['topk_idx'], res['topk_idx']) and torch.equal(got['anchor_pos'], pos), step
    assert cache.hits >= 6 and len(cache.graphs) <= 3
