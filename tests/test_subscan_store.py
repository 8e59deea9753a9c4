"""Packed sub-scan store + Scan3R collation against the collated batches the UNMODIFIED
reference ``Scan3RDataset`` produced on the same sub-scans (tests/golden/collate_ref.npz, written by
oracle/make_golden_collate.py)."""
import json
import os

import numpy as np
import pytest
import torch

from tests.util import GOLD

T_KEYS = ('tot_obj_pts', 'tot_bow_vec_object_attr_feats', 'tot_bow_vec_object_edge_feats', 'tot_rel_pose', 'edges')


@pytest.fixture(scope='module')
def gold():
    return np.load(os.path.join(GOLD, 'collate_ref.npz'), allow_pickle=False)


@pytest.fixture(scope='module')
def store(gold, tmp_path_factory):
    from sgaligner_b200.subscan_store import SubscanStore
    scans = []
    for sid in gold['scan_ids']:
        sid = str(sid)
        d = {k: gold[f'scan/{sid}/{k}'] for k in ('objects_id', 'objects_cat', 'edges', 'rel_trans', 'bow_vec_object_attr_feats',
                                                   'bow_vec_object_edge_feats', 'obj_points', 'center')}
        d['scan_id'] = sid
        scans.append(d)
    path = str(tmp_path_factory.mktemp('store') / 'subscans.sga')
    SubscanStore.pack(path, scans, n_points=int(scans[0]['obj_points'].shape[1]))
    return SubscanStore(path)


def _check(out, gold, split, centred_pts):
    for k in gold.files:
        if not k.startswith(f'out/{split}/'):
            continue
        key = k.split('/', 2)[2]
        ref = gold[k]
        if key == 'tot_obj_pts':
            got = centred_pts
        else:
            got = out[key]
        got = got.numpy() if torch.is_tensor(got) else np.asarray(got)
        assert got.shape == ref.shape, (key, got.shape, ref.shape)
        if ref.dtype.kind in 'US':
            assert (got.astype(str) == ref.astype(str)).all(), key
        else:
            assert got.dtype == ref.dtype, (key, got.dtype, ref.dtype)
            assert np.array_equal(got, ref), key


def test_store_roundtrip(store, gold):
    assert len(store) == len(gold['scan_ids'])
    for sid in store.ids():
        s = store[sid]
        assert np.array_equal(s['obj_points'], gold[f'scan/{sid}/obj_points'])
        assert np.array_equal(s['edges'], gold[f'scan/{sid}/edges'])
        assert np.array_equal(s['rel_trans'], gold[f'scan/{sid}/rel_trans'])
        assert s['bow_vec_object_attr_feats'].dtype == np.uint8            # packed 8x smaller, lossless
        assert np.array_equal(s['bow_vec_object_attr_feats'].astype(np.float64), gold[f'scan/{sid}/bow_vec_object_attr_feats'])
        assert np.array_equal(s['bow_vec_object_edge_feats'].astype(np.float64), gold[f'scan/{sid}/bow_vec_object_edge_feats'])
        assert np.array_equal(s['center'], gold[f'scan/{sid}/center'])
    with pytest.raises(ValueError):
        from sgaligner_b200.subscan_store import SubscanStore
        SubscanStore(os.path.join(GOLD, 'collate_ref.npz'))


@pytest.mark.parametrize('split', ['train', 'val'])
def test_dataloader_api_matches_reference(store, gold, split):
    """``__getitem__`` + ``collate_fn`` (the torch DataLoader path) == the reference's, bit for bit."""
    from sgaligner_b200.subscan_store import Scan3RPacked
    ds = Scan3RPacked(store, json.loads(str(gold['anchor_data'])), split=split, pinned=False)
    np.random.seed(123)
    out = ds.collate_fn([ds[i] for i in range(len(ds))])
    assert set(out) == {k.split('/', 2)[2] for k in gold.files if k.startswith(f'out/{split}/')}
    _check(out, gold, split, out['tot_obj_pts'])


@pytest.mark.parametrize('split', ['train', 'val'])
def test_collate_pairs_matches_reference(store, gold, split):
    """The one-pass collation: everything but the points is final on the host; the points are raw and
    ``raw - center[pair]`` (what ``sga_center_points`` computes on the device) equals the reference's."""
    from sgaligner_b200.subscan_store import Scan3RPacked
    ds = Scan3RPacked(store, json.loads(str(gold['anchor_data'])), split=split, pinned=False)
    np.random.seed(123)
    out = ds.collate_pairs(range(len(ds)))
    pair = np.repeat(np.arange(out['batch_size']), out['tot_obj_count'])
    centred = out['tot_obj_pts'].numpy() - out['_sga_center'].numpy()[pair][:, None, :]
    _check(out, gold, split, centred)
    # staging buffers are reused: a second batch of other pairs must not corrupt shapes
    out2 = ds.collate_pairs([1])
    assert out2['tot_obj_pts'].shape[0] == int(out2['tot_obj_count'].sum())


@pytest.mark.gpu
@pytest.mark.parametrize('split', ['train', 'val'])
def test_to_device_centres_on_gpu(store, gold, split):
    from sgaligner_b200.subscan_store import Scan3RPacked, to_device
    ds = Scan3RPacked(store, json.loads(str(gold['anchor_data'])), split=split)
    np.random.seed(123)
    d = to_device(ds.collate_pairs(range(len(ds))), torch.device('cuda:0'), n_chunks=3)
    torch.cuda.synchronize()
    assert np.array_equal(d['tot_obj_pts'].cpu().numpy(), gold[f'out/{split}/tot_obj_pts'])
    for k in T_KEYS[1:]:
        assert np.array_equal(d[k].cpu().numpy(), gold[f'out/{split}/{k}']), k
    # the batch feeds the encoder + loss unchanged
    from sgaligner_b200.losses import CustomMultiLossLayer, OverallLoss
    from sgaligner_b200.sg_aligner import MultiModalEncoder
    mods = ['point', 'gat', 'rel', 'attr']
    dev = torch.device('cuda:0')
    torch.manual_seed(0)
    model = MultiModalEncoder(modules=mods, rel_dim=41, attr_dim=164).to(dev)
    fn = OverallLoss(CustomMultiLossLayer(4).to(dev), CustomMultiLossLayer(4).to(dev), dev,
                     {'zoom': 0.1, 'wt_align_loss': 1.0, 'wt_contrastive_loss': 1.0, 'modules': mods})
    loss = fn(model(d), d)['loss']
    assert torch.isfinite(loss)


@pytest.mark.gpu
def test_staging_ring_is_safe_with_the_host_running_ahead(store, gold):
    """``to_device(ds.collate_pairs(idx))`` in a loop with NO synchronisation and a busy GPU: the pinned staging sets
    are event-guarded, so no batch is overwritten while its H2D copy is still queued (silent corruption)."""
    from sgaligner_b200.subscan_store import Scan3RPacked, to_device
    dev = torch.device('cuda:0')
    ds = Scan3RPacked(store, json.loads(str(gold['anchor_data'])), split='val', n_staging=2)
    n = len(ds)
    want = {}
    for i in range(n):
        b = Scan3RPacked(store, json.loads(str(gold['anchor_data'])), split='val', pinned=False).collate_pairs([i])
        pair = np.repeat(np.arange(b['batch_size']), b['tot_obj_count'])
        want[i] = (b['tot_obj_pts'].numpy() - b['_sga_center'].numpy()[pair][:, None, :], b['tot_rel_pose'].numpy().copy())
    big = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
    got = []
    order = [i % n for i in range(24)]
    for i in order:
        for _ in range(4):
            big.zero_()                               # keep the compute stream (which the copy stream waits for) busy
        got.append(to_device(ds.collate_pairs([i]), dev, n_chunks=2))
    torch.cuda.synchronize()
    for i, d in zip(order, got):
        assert np.array_equal(d['tot_obj_pts'].cpu().numpy(), want[i][0]), i
        assert np.array_equal(d['tot_rel_pose'].cpu().numpy(), want[i][1]), i


def test_pack_from_reference_files_and_dataloader(gold, tmp_path):
    """The converter reads the reference's own on-disk formats (``files/<mode>/data/<id>.pkl`` +
    ``scans/<id>/data.npy``, written here exactly as preprocess.py / the 3RScan export lay them out) and a stock
    ``torch.utils.data.DataLoader`` over the packed dataset yields the reference's batches."""
    import pickle
    from sgaligner_b200.subscan_store import Scan3RPacked, SubscanStore
    files, scans = tmp_path / 'files', tmp_path / 'scans'
    (files / 'orig' / 'data').mkdir(parents=True)
    ids = [str(s) for s in gold['scan_ids']]
    P = int(gold[f'scan/{ids[0]}/obj_points'].shape[1])
    for sid in ids:
        d = {k: gold[f'scan/{sid}/{k}'] for k in ('objects_id', 'objects_cat', 'edges', 'rel_trans', 'bow_vec_object_attr_feats',
                                                   'bow_vec_object_edge_feats')}
        d['obj_points'] = {P: gold[f'scan/{sid}/obj_points']}
        d['object_id2idx'] = {int(v): i for i, v in enumerate(d['objects_id'])}
        with open(files / 'orig' / 'data' / f'{sid}.pkl', 'wb') as h:
            pickle.dump(d, h)
        # a point cloud whose mean is the stored centre (the converter only takes the mean, scan3r.py:66-75)
        c = gold[f'scan/{sid}/center'].astype(np.float32)
        ply = np.zeros(1, dtype=[('x', 'f4'), ('y', 'f4'), ('z', 'f4'), ('objectId', 'i4')])
        ply['x'], ply['y'], ply['z'] = c[0], c[1], c[2]
        (scans / sid).mkdir(parents=True)
        np.save(scans / sid / 'data.npy', ply)
    path = SubscanStore.pack_from_reference_files(str(files), str(scans), 'orig', ids, str(tmp_path / 'packed.sga'), n_points=P)
    ds = Scan3RPacked(SubscanStore(path), json.loads(str(gold['anchor_data'])), split='val', pinned=False)
    loader = torch.utils.data.DataLoader(ds, batch_size=len(ds), shuffle=False, collate_fn=ds.collate_fn, num_workers=0)
    out = next(iter(loader))
    _check(out, gold, 'val', out['tot_obj_pts'])
    # two batches of two pairs: per-batch index offsets restart
    parts = list(torch.utils.data.DataLoader(ds, batch_size=2, shuffle=False, collate_fn=ds.collate_fn))
    assert [p['batch_size'] for p in parts] == [2, 2]
    assert int(parts[1]['e1i'].min()) < int(parts[1]['tot_obj_count'][0])
