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
