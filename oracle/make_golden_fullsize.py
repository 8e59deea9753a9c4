"""Golden vectors at BASELINE.json's FULL sizes from the UNMODIFIED reference modules (``/root/reference`` through
:mod:`oracle.ref_import`): C2 = configs[1] (32 pairs x (64+64) objects x 512 points, PointNet + GAT) and C3 =
configs[2] (128 3RScan-shaped pairs, P+S+R+A, 512 points).  One training-mode forward + ``OverallLoss`` + backward
each, on the CPU of the build container (a few minutes).  TEST INFRASTRUCTURE ONLY.

    python -m oracle.make_golden_fullsize

The inputs are NOT stored: ``sgaligner_b200.synthetic.config_c2 / config_c3`` regenerate them from their seed on the
GPU box (same image, same numpy); a checksum of the inputs is stored so that a drifting generator fails loudly.
Stored per config (``tests/golden/full_c2.npz`` / ``full_c3.npz``): the parameters, every ``stride``-th row of every
embedding, fp64 checksums (sum, sum of squares) of the complete embeddings, the four loss values, every parameter
gradient, the BatchNorm running statistics after the step, Hits@1..5 and the reciprocal ranks.
"""
from __future__ import annotations

import os
import sys
import time
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_import, sgaligner_oracle as O          # noqa: E402
from sgaligner_b200 import synthetic                          # noqa: E402

GOLD = os.path.join(ROOT, 'tests', 'golden')
STRIDE = 8


def input_checksum(data: dict) -> np.ndarray:
    return np.array([float(data['tot_obj_pts'].double().sum()), float(data['tot_obj_pts'].double().pow(2).sum()),
                     float(data['edges'].double().sum()), float(data['tot_rel_pose'].double().sum()),
                     float(data['tot_bow_vec_object_attr_feats'].sum()), float(data['tot_bow_vec_object_edge_feats'].sum()),
                     float(np.asarray(data['e1i']).sum()), float(np.asarray(data['e2j']).sum())])


CONFIGS = {
    'full_c2': (lambda: synthetic.config_c2(batch=32, seed=0), ['point', 'gat'], 0),
    'full_c3': (lambda: synthetic.config_c3(batch=128, seed=1, train=True), ['point', 'gat', 'rel', 'attr'], 1),
}


def main():
    sg, ls, al = ref_import.load_reference()
    torch.set_num_threads(os.cpu_count())
    for name, (gen, modules, seed) in CONFIGS.items():
        t0 = time.time()
        data = gen()
        M = len(modules)
        torch.manual_seed(seed)
        model = sg.MultiModalEncoder(modules=modules, rel_dim=41, attr_dim=164)
        li, lc = ls.CustomMultiLossLayer(M), ls.CustomMultiLossLayer(M)
        params0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
        fn = ls.OverallLoss(li, lc, 'cpu', {'zoom': 0.1, 'wt_align_loss': 1.0, 'wt_contrastive_loss': 1.0, 'modules': modules})
        model.train()
        out = model(data)
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            ld = fn(out, data)
        ld['loss'].backward()
        blob = {'cfg/modules': np.array(modules), 'cfg/stride': np.array(STRIDE), 'in/checksum': input_checksum(data)}
        for k, v in params0.items():
            blob['p/' + k] = v.numpy()
        for k, v in out.items():
            v = v.detach()
            blob['out/' + k] = v[::STRIDE].numpy()
            blob['sum/' + k] = np.array([float(v.double().sum()), float(v.double().pow(2).sum())])
        for k, v in ld.items():
            blob['loss/' + k] = np.array(float(v))
        for n_, p_ in model.named_parameters():
            if p_.grad is not None:
                blob['grad/' + n_] = p_.grad.numpy()
        blob['grad/__lv_ial'] = li.log_vars.grad.numpy() if li.log_vars.grad is not None else np.zeros(M, np.float32)
        blob['grad/__lv_icl'] = lc.log_vars.grad.numpy() if lc.log_vars.grad is not None else np.zeros(M, np.float32)
        for k, v in model.state_dict().items():
            if 'running' in k or 'num_batches' in k:
                blob['bn/' + k] = v.numpy()
        emb = out['joint'].detach()
        ev = O.evaluate_batch(emb, data)            # restatement pinned to utils/alignment.py by make_golden.py
        # ... and once more through the reference's own functions on this batch
        offs = O.pair_offsets(data)
        hits, rr, a0 = {k: 0 for k in range(1, 6)}, [], 0
        for b in range(data['batch_size']):
            o0, o1 = int(offs[b]), int(offs[b + 1]); na = int(data['e1i_count'][b])
            e1 = data['e1i'][a0:a0 + na] - o0; e2 = data['e2i'][a0:a0 + na] - o0; a0 += na
            e = emb[o0:o1]; e = e / e.norm(dim=1)[:, None]
            rank = torch.argsort(1 - torch.mm(e, e.transpose(0, 1)), dim=1)
            rr = al.compute_mean_reciprocal_rank(rank, e1, e2, rr)
            for k in hits:
                hits[k] += al.compute_hits_k(rank, e1, e2, k)[0]
        assert ev['hits'] == hits, (ev['hits'], hits)
        blob['metric/hits'] = np.array([hits[k] for k in range(1, 6)])
        blob['metric/rr'] = np.sort(np.array(rr))
        # the restatement agrees with the reference at this size too
        with torch.no_grad():
            o_out = O.encoder_forward(params0, data, modules)
        worst = max(float((o_out[k] - out[k].detach()).abs().max() / out[k].detach().abs().max()) for k in out)
        assert worst < 5e-6, worst
        path = os.path.join(GOLD, name + '.npz')
        np.savez_compressed(path, **blob)
        print(f'{name}: N={data["tot_obj_pts"].shape[0]} loss={float(ld["loss"]):.6g} hits={hits} oracle-vs-ref {worst:.1e} '
              f'{os.path.getsize(path) // 1024} KiB  {time.time() - t0:.0f}s')


if __name__ == '__main__':
    main()
