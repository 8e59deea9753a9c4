"""Golden vectors for the batch collation: drive the UNMODIFIED reference
``Scan3RDataset.__getitem__`` + ``collate_fn`` (``src/datasets/scan3r.py``) on a small synthetic dataset tree
written in the reference's own on-disk formats (``scans/<id>/data.npy``, ``files/<mode>/data/<id>.pkl``,
``files/<mode>/anchors<type>_<split>.json``) and freeze inputs + collated outputs in
``tests/golden/collate_ref.npz``.  ``plyfile`` (imported by ``utils/scan3r.py:5``, unused on this path) is
stubbed.  TEST INFRASTRUCTURE ONLY; needs /root/reference.

    python -m oracle.make_golden_collate
"""
from __future__ import annotations

import json
import os
import pickle
import sys
import tempfile
import types

import numpy as np
import torch

from oracle.ref_import import REFERENCE_ROOT

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')
P, ATTR, REL = 16, 164, 41


def synth_scan(rng, scan_id: str, obj_ids):
    n = len(obj_ids)
    cloud = (rng.standard_normal((n * 40, 3)) + rng.uniform(-3, 3, (1, 3))).astype(np.float32)
    ply = np.zeros(cloud.shape[0], dtype=[('x', 'f4'), ('y', 'f4'), ('z', 'f4'), ('objectId', 'i4')])
    ply['x'], ply['y'], ply['z'] = cloud[:, 0], cloud[:, 1], cloud[:, 2]
    ply['objectId'] = np.repeat(obj_ids, 40)
    pts = np.stack([cloud[i * 40:(i + 1) * 40][rng.integers(0, 40, P)] for i in range(n)]).astype(np.float32)
    s, d = np.meshgrid(np.arange(n), np.arange(n), indexing='ij')
    keep = s != d
    edges = np.stack([s[keep], d[keep]], 1)                    # complete digraph (preprocess.py:176-182)
    perm = rng.permutation(edges.shape[0])
    data = {
        'scan_id': scan_id, 'objects_id': np.array(obj_ids), 'global_objects_id': np.array(obj_ids) + 100,
        'objects_cat': np.array(obj_ids) + 100, 'edges': edges[perm], 'obj_points': {P: pts},
        'objects_count': n, 'edges_count': edges.shape[0], 'object_id2idx': {int(v): i for i, v in enumerate(obj_ids)},
        'rel_trans': rng.standard_normal((n, 3)),
        'bow_vec_object_attr_feats': (rng.random((n, ATTR)) < 0.05).astype(np.float64),
        'bow_vec_object_edge_feats': rng.poisson(1.0, (n, REL)).astype(np.float64),
    }
    return ply, data


def main():
    if 'plyfile' not in sys.modules:
        m = types.ModuleType('plyfile')
        m.PlyData = object
        sys.modules['plyfile'] = m
    for pth in (os.path.join(REFERENCE_ROOT, 'src'), REFERENCE_ROOT):
        if pth not in sys.path:
            sys.path.insert(0, pth)
    from datasets.scan3r import Scan3RDataset      # the reference class, unmodified

    rng = np.random.default_rng(7)
    scans = {
        'scanA_0': [1, 2, 3, 5, 8, 13, 21], 'scanA_1': [2, 3, 5, 8, 34, 55], 'scanB_0': [4, 6, 7, 9, 10],
        'scanB_1': [6, 7, 9, 11, 12, 14, 15, 16], 'scanC_0': [1, 2, 3], 'scanC_1': [3, 2, 1, 17],
    }
    anchor_data = [
        {'src': 'scanA_0', 'ref': 'scanA_1', 'overlap': 0.55, 'anchorIds': [2, 3, 5, 8, 0, 99]},
        {'src': 'scanB_0', 'ref': 'scanB_1', 'overlap': 0.31, 'anchorIds': [6, 7, 9]},
        {'src': 'scanC_0', 'ref': 'scanC_1', 'overlap': 0.80, 'anchorIds': [1, 2, 3]},
        {'src': 'scanB_1', 'ref': 'scanB_0', 'overlap': 0.31, 'anchorIds': [9, 7, 6]},
    ]
    blob = {}
    with tempfile.TemporaryDirectory() as root:
        for mode in ('orig',):
            os.makedirs(os.path.join(root, 'files', mode, 'data'))
        for sid, ids in scans.items():
            ply, data = synth_scan(rng, sid, ids)
            os.makedirs(os.path.join(root, 'scans', sid))
            np.save(os.path.join(root, 'scans', sid, 'data.npy'), ply)
            with open(os.path.join(root, 'files', 'orig', 'data', sid + '.pkl'), 'wb') as h:
                pickle.dump(data, h)
            cloud = np.stack([ply['x'], ply['y'], ply['z']]).transpose((1, 0))
            blob[f'scan/{sid}/center'] = np.mean(cloud, axis=0)
            for k in ('objects_id', 'objects_cat', 'edges', 'rel_trans', 'bow_vec_object_attr_feats', 'bow_vec_object_edge_feats'):
                blob[f'scan/{sid}/{k}'] = np.asarray(data[k])
            blob[f'scan/{sid}/obj_points'] = data['obj_points'][P]
        for split in ('train', 'val'):
            with open(os.path.join(root, 'files', 'orig', f'anchors_gold_{split}.json'), 'w') as h:
                json.dump(anchor_data, h)
        ns = types.SimpleNamespace
        cfg = ns(val=ns(pc_res=P, data_mode='orig', overlap_low=0.0, overlap_high=0.0), train=ns(pc_res=P, use_augmentation=False,
                 rot_factor=1.0, augmentation_noise=0.005), preprocess=ns(anchor_type_name='_gold'), model_name='sgaligner',
                 scan_type='subscan', data=ns(root_dir=root, subscan_dir=root))
        for split in ('train', 'val'):
            ds = Scan3RDataset(cfg, split)
            np.random.seed(123)                        # scan3r.py:69 draws from the global numpy RNG
            out = ds.collate_fn([ds[i] for i in range(len(ds))])
            for k, v in out.items():
                blob[f'out/{split}/{k}'] = v.numpy() if torch.is_tensor(v) else np.asarray(v)
            print(split, {k: (tuple(v.shape) if hasattr(v, 'shape') else v) for k, v in out.items()})
    blob['anchor_data'] = np.array(json.dumps(anchor_data))
    blob['scan_ids'] = np.array(list(scans))
    np.savez_compressed(os.path.join(GOLD, 'collate_ref.npz'), **blob)
    print('wrote collate_ref.npz', os.path.getsize(os.path.join(GOLD, 'collate_ref.npz')) // 1024, 'KiB')


if __name__ == '__main__':
    main()
