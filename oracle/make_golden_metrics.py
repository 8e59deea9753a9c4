"""Pin the oracle's SGAR / node-correspondence / alignment-score restatements against the UNMODIFIED reference
functions (``utils/alignment.py:27-89``), run on the similarity matrices and rank lists already frozen in
``tests/golden/*.npz`` (which the reference itself produced, see make_golden.py).  Writes
``tests/golden/metrics_ref.npz``.

    python -m oracle.make_golden_metrics        (needs /root/reference; run in the build container only)
"""
from __future__ import annotations

import glob
import os

import numpy as np
import torch

from oracle import ref_import
from oracle import sgaligner_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')


def main():
    _sg, _ls, al = ref_import.load_reference()
    blob = {}
    for path in sorted(glob.glob(os.path.join(GOLD, '*.npz'))):
        name = os.path.basename(path)[:-4]
        if name == 'metrics_ref':
            continue
        g = np.load(path, allow_pickle=True)
        B = int(g['in/batch_size'])
        goc = g['in/graph_per_obj_count'].reshape(-1, 2)
        offs = np.concatenate([[0], np.cumsum(goc.sum(1))])
        a0 = 0
        for b in range(B):
            na = int(g['in/e1i_count'][b])
            e1 = g['in/e1i'][a0:a0 + na].astype(np.int64) - offs[b]
            e2 = g['in/e2i'][a0:a0 + na].astype(np.int64) - offs[b]
            a0 += na
            rank = torch.from_numpy(g[f'rank/{b}'].astype(np.int64))
            sim = torch.from_numpy(g[f'sim/{b}'])
            ns, nr = int(goc[b, 0]), int(goc[b, 1])
            corr = al.compute_node_corrs(rank, ns, 1)
            score = al.compute_alignment_score(rank, ns, nr)
            blob[f'{name}/{b}/node_corrs'] = np.asarray(corr, dtype=np.int32).reshape(-1, 2)
            blob[f'{name}/{b}/alignment_score'] = np.array(score)
            assert O.node_corrs(rank.numpy(), ns, 1) == [tuple(int(v) for v in c) for c in corr], (name, b)
            assert abs(O.alignment_score(rank.numpy(), ns, nr) - score) < 1e-12
            if na:
                sv = al.compute_sgar(sim, rank, e1, e2, ['2', '50', '100'])
                ov = O.sgar(sim.numpy(), rank.numpy(), e1, e2)
                assert sv == ov, (name, b, sv, ov)
                blob[f'{name}/{b}/sgar'] = np.array([sv['2'], sv['50'], sv['100']])
            print(f'{name}/{b}: anchors {na} sgar {blob.get(f"{name}/{b}/sgar")} align {score:.4f} corrs {len(corr)}')
    np.savez_compressed(os.path.join(GOLD, 'metrics_ref.npz'), **blob)


if __name__ == '__main__':
    main()
