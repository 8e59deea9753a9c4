"""Throughput of the other BASELINE.json configurations on one GPU (bench.py's line is C2 / configs[1]): the
serving step (encoder forward + matching head + anchor positions, CUDA-graph replay) and the training step
(forward + OverallLoss + backward + Adam), device-resident inputs, L2 flushed between timed steps.  One JSON
line per configuration.    python tools/bench_configs.py [--steps 10]"""
import argparse, json, os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
import torch
from sgaligner_b200 import matching, ops, synthetic, to_cuda
from sgaligner_b200.losses import CustomMultiLossLayer, OverallLoss
from sgaligner_b200.serving import CapturedInference
from sgaligner_b200.sg_aligner import MultiModalEncoder
from sgaligner_b200.trainer import FlatAdam, train_step

ap = argparse.ArgumentParser()
ap.add_argument('--steps', type=int, default=10)
args = ap.parse_args()
dev = torch.device('cuda:0')
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
ALL = ['point', 'gat', 'rel', 'attr']
CONFIGS = [
    ('C2 (configs[1]) 32 pairs x (64+64) obj x 512 pts, point+gat', ['point', 'gat'], lambda: synthetic.config_c2(batch=32, seed=100), {}),
    ('C2 shapes with all four modalities (joint 400-d)', ALL, lambda: synthetic.config_c2(batch=32, seed=100), {}),
    ('C3 (configs[2]) 128 3RScan-shaped pairs, complete digraphs, P+S+R+A, 512 pts', ALL, lambda: synthetic.config_c3(batch=128, seed=1), {}),
    ('C5 (configs[4]) per-GPU share: 8 pairs x (256+256) obj x 1024 pts, pt_out 512, emb 128 (joint 512-d)', ALL,
     lambda: synthetic.config_c5(batch=8, seed=2), {'pt_out_dim': 512, 'emb_dim': 128}),
]


def timed(fn, steps, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(steps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


for name, mods, make, kw in CONFIGS:
    host = make()
    data = to_cuda(dict(host), dev)
    B = int(host['batch_size'])
    torch.manual_seed(0)
    model = MultiModalEncoder(modules=mods, rel_dim=41, attr_dim=164, **kw).to(dev)
    M = len(mods)
    li, lc = CustomMultiLossLayer(M).to(dev), CustomMultiLossLayer(M).to(dev)
    fn = OverallLoss(li, lc, dev, {'zoom': 0.1, 'wt_align_loss': 1.0, 'wt_contrastive_loss': 1.0, 'modules': mods})
    model.eval()
    e1 = torch.as_tensor(host['e1i']).to(dev); e2 = torch.as_tensor(host['e2i']).to(dev)

    def serve():
        with torch.no_grad():
            out = model(data)
            res = matching.match_batch(out['joint'], data, k=6, full_rank=False)
            return ops.match_anchor_pos(res['sim'], res['layout'], e1, e2)
    eager_ms = timed(serve, args.steps)
    cap = CapturedInference(model, data, k=6)
    graph_ms = timed(cap.replay, args.steps)
    pos = serve()
    del cap
    model.train()
    opt = FlatAdam(list(model.parameters()) + list(li.parameters()) + list(lc.parameters()), lr=1e-3, weight_decay=1e-6)
    l0 = float(train_step(model, fn, opt, data)['loss'])
    train_ms = timed(lambda: train_step(model, fn, opt, data), args.steps)
    l1 = float(train_step(model, fn, opt, data)['loss'])
    print(json.dumps({'config': name, 'pairs': B, 'objects': int(data['tot_obj_pts'].shape[0]), 'edges': int(data['edges'].shape[0]),
                      'anchors': int(len(host['e1i'])), 'serve_ms_graph': graph_ms, 'serve_ms_eager': eager_ms,
                      'serve_pairs_per_s': B / (graph_ms * 1e-3), 'train_ms': train_ms, 'train_pairs_per_s': B / (train_ms * 1e-3),
                      'hits_at_1_untrained': float((pos < 1).float().mean()), 'loss_first': l0, 'loss_after_%d_steps' % (args.steps + 5): l1}))
    del model, opt, data
    torch.cuda.empty_cache()
