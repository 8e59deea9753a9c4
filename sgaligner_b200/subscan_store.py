"""Packed on-disk store of preprocessed sub-scans + the Scan3R batch collation on top of it.

The reference keeps one pickle per sub-scan (``preprocessing/scan3r/preprocess.py:195-211,321,357``:
``obj_points{512: [n,512,3]}``, ``edges [e,2]``, ``rel_trans [n,3]``, ``bow_vec_object_{attr,edge}_feats``,
``objects_id``, ``objects_cat``, ``object_id2idx``) and, per sample, re-reads two pickles plus two
``scans/<id>/data.npy`` point clouds (only to take their mean), concatenates, centres and converts them
(``src/datasets/scan3r.py:59-140``); the collate then concatenates everything again (``:142-209``).

Here all sub-scans of a split live in ONE file of 64-byte aligned raw arrays behind a small JSON index
(``SubscanStore``, opened with ``np.memmap``), and a batch is assembled by ONE copy per array straight
from the mapping into pinned host buffers (``Scan3RPacked.collate_pairs``) -- no pickle parsing, no
intermediate ``torch.cat``; bag-of-words vectors are stored as ``uint8`` counts (8x smaller) and edges as
``int32`` and widened during that copy.  The per-pair centring of the points (``scan3r.py:99-100``) is
not done on the host at all: the raw points are uploaded and ``sga_center_points`` subtracts the centre
on the device (``to_device``).  The resulting dict obeys the reference's dataloader contract key for key
(SURVEY.md section 8b), so it feeds ``MultiModalEncoder`` / ``OverallLoss`` / the reference trainer unchanged.

File layout (little endian):
    0   8s   magic  b'SGASUBS1'
    8   u32  version (1)      12  u32  number of sub-scans
    16  u32  points per object 20 u32  attr_dim      24 u32 rel_dim      28 u32 reserved
    32  u64  index offset     40  u64  index bytes (UTF-8 JSON)
    64  ...  arrays, each 64-byte aligned; the index gives, per sub-scan, counts, centre and array offsets
"""
from __future__ import annotations

import json
import os
import pickle
import struct
from typing import Dict, Iterable, List, Optional, Sequence

import numpy as np
import torch

MAGIC = b'SGASUBS1'
_ALIGN = 64


def _bow_dtype(a: np.ndarray):
    """uint8 when the bag-of-words counts are small non-negative integers (they are word counts), else f64."""
    a = np.asarray(a)
    if a.size == 0 or (np.all(a == np.floor(a)) and a.min() >= 0 and a.max() <= 255):
        return np.uint8
    return np.float64


class SubscanStore:
    """Read side: ``store[scan_id]`` -> dict of zero-copy views into the mapping."""

    def __init__(self, path: str):
        self.path = path
        with open(path, 'rb') as f:
            head = f.read(64)
        if head[:8] != MAGIC:
            raise ValueError(f'{path}: not a packed sub-scan store (bad magic)')
        self.version, self.n_scans, self.n_points, self.attr_dim, self.rel_dim, _ = struct.unpack('<6I', head[8:32])
        if self.version != 1:
            raise ValueError(f'{path}: unsupported version {self.version}')
        idx_off, idx_len = struct.unpack('<2Q', head[32:48])
        self._mm = np.memmap(path, dtype=np.uint8, mode='r')
        self.index = json.loads(bytes(self._mm[idx_off:idx_off + idx_len]).decode('utf-8'))
        self._by_id = {e['id']: e for e in self.index}

    def __len__(self):
        return self.n_scans

    def __contains__(self, scan_id):
        return scan_id in self._by_id

    def ids(self) -> List[str]:
        return [e['id'] for e in self.index]

    def _view(self, off: int, dtype, shape):
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        return self._mm[off:off + n].view(dtype).reshape(shape)

    def __getitem__(self, scan_id: str) -> dict:
        e = self._by_id[scan_id]
        n, ne, o = e['n'], e['e'], e['off']
        return {
            'scan_id': scan_id, 'n': n, 'e': ne, 'center': np.asarray(e['center'], dtype=np.float32),
            'obj_points': self._view(o['points'], np.float32, (n, self.n_points, 3)),
            'edges': self._view(o['edges'], np.int32, (ne, 2)),
            'rel_trans': self._view(o['rel_trans'], np.float64, (n, 3)),
            'bow_vec_object_attr_feats': self._view(o['attr'], np.dtype(e['bow_dtype']), (n, self.attr_dim)),
            'bow_vec_object_edge_feats': self._view(o['rel'], np.dtype(e['bow_dtype']), (n, self.rel_dim)),
            'objects_id': self._view(o['objects_id'], np.int64, (n,)),
            'objects_cat': self._view(o['objects_cat'], np.int64, (n,)),
        }

    # ------------------------------------------------------------------ write side
    @staticmethod
    def pack(path: str, scans: Iterable[dict], n_points: int) -> str:
        """``scans``: dicts with the reference pickle's keys (``scan_id``, ``objects_id``, ``objects_cat``,
        ``edges``, ``obj_points`` (array or ``{res: array}``), ``rel_trans``, ``bow_vec_object_attr_feats``,
        ``bow_vec_object_edge_feats``) plus ``center`` = mean of the scan's full point cloud
        (``scan3r.py:66-75`` takes it from ``scans/<id>/data.npy``)."""
        index = []
        attr_dim = rel_dim = None
        with open(path, 'wb') as f:
            f.write(b'\0' * 64)

            def put(a: np.ndarray) -> int:
                pad = (-f.tell()) % _ALIGN
                f.write(b'\0' * pad)
                off = f.tell()
                f.write(np.ascontiguousarray(a).tobytes())
                return off

            for s in scans:
                pts = s['obj_points']
                pts = pts[n_points] if isinstance(pts, dict) else pts
                pts = np.asarray(pts)
                n = int(pts.shape[0])
                assert pts.shape == (n, n_points, 3), pts.shape
                attr = np.asarray(s['bow_vec_object_attr_feats'])
                rel = np.asarray(s['bow_vec_object_edge_feats'])
                attr_dim = attr.shape[1] if attr_dim is None else attr_dim
                rel_dim = rel.shape[1] if rel_dim is None else rel_dim
                assert attr.shape == (n, attr_dim) and rel.shape == (n, rel_dim)
                bdt = np.uint8 if (_bow_dtype(attr) == np.uint8 and _bow_dtype(rel) == np.uint8) else np.float64
                edges = np.asarray(s['edges']).reshape(-1, 2)
                assert edges.size == 0 or (edges.min() >= 0 and edges.max() < n)
                off = {
                    'points': put(pts.astype(np.float32, copy=False)),
                    'edges': put(edges.astype(np.int32)),
                    'rel_trans': put(np.asarray(s['rel_trans'], dtype=np.float64).reshape(n, 3)),
                    'attr': put(attr.astype(bdt)),
                    'rel': put(rel.astype(bdt)),
                    'objects_id': put(np.asarray(s['objects_id'], dtype=np.int64)),
                    'objects_cat': put(np.asarray(s['objects_cat'], dtype=np.int64)),
                }
                index.append({'id': str(s['scan_id']), 'n': n, 'e': int(edges.shape[0]),
                              'center': [float(v) for v in np.asarray(s['center'], dtype=np.float32)],
                              'bow_dtype': np.dtype(bdt).name, 'off': off,
                              'points_f64_source': bool(pts.dtype == np.float64)})
            pad = (-f.tell()) % _ALIGN
            f.write(b'\0' * pad)
            idx_off = f.tell()
            blob = json.dumps(index).encode('utf-8')
            f.write(blob)
            f.seek(0)
            f.write(MAGIC + struct.pack('<6I', 1, len(index), n_points, attr_dim or 0, rel_dim or 0, 0) +
                    struct.pack('<2Q', idx_off, len(blob)))
        return path

    @staticmethod
    def pack_from_reference_files(files_dir: str, scans_dir: str, mode: str, scan_ids: Sequence[str], out_path: str,
                                  n_points: int = 512) -> str:
        """Convert the reference's own preprocessing output (``<files_dir>/<mode>/data/<id>.pkl`` and
        ``<scans_dir>/<id>/data.npy``) into one packed store."""
        def gen():
            for sid in scan_ids:
                with open(os.path.join(files_dir, mode, 'data', f'{sid}.pkl'), 'rb') as h:
                    d = pickle.load(h)
                ply = np.load(os.path.join(scans_dir, sid, 'data.npy'))
                pts = np.stack([ply['x'], ply['y'], ply['z']]).transpose((1, 0))     # utils/scan3r.py:98-100
                d = dict(d)
                d['scan_id'] = sid
                d['center'] = np.mean(pts, axis=0)
                yield d
        return SubscanStore.pack(out_path, gen(), n_points)


class _Staging:
    """Reusable pinned host buffers of ONE batch (grown on demand).  ``busy`` is the CUDA event recorded after the
    last asynchronous H2D copy that reads these buffers (:func:`to_device` sets it); :meth:`wait` blocks the host
    until that copy has drained, so the buffers are never rewritten under a queued copy."""

    def __init__(self, pinned: bool):
        self.pinned = pinned
        self.buf: Dict[str, torch.Tensor] = {}
        self.busy = None

    def wait(self):
        if self.busy is not None:
            self.busy.synchronize()
            self.busy = None

    def get(self, key: str, shape, dtype) -> torch.Tensor:
        n = int(np.prod(shape))
        t = self.buf.get(key)
        if t is None or t.numel() < n or t.dtype != dtype:
            t = torch.empty(max(n, 1), dtype=dtype)
            if self.pinned:
                t = t.pin_memory()
            self.buf[key] = t
        return t[:n].view(*shape)


class Scan3RPacked(torch.utils.data.Dataset):
    """``Scan3RDataset`` (``src/datasets/scan3r.py``) over a :class:`SubscanStore`.

    ``anchor_data``: the list the reference reads from ``anchors<type>_<split>.json``
    (``[{'src','ref','overlap','anchorIds'}, ...]``, ``subgenscan3r.py:117-118``).
    ``__getitem__`` / ``collate_fn`` are kept for ``torch.utils.data.DataLoader`` users and return exactly
    what the reference returns; :meth:`collate_pairs` is the fast path (one pass into pinned memory, raw
    points + ``pcl_center``; pair with :func:`to_device`)."""

    def __init__(self, store: SubscanStore, anchor_data: Sequence[dict], split: str = 'train', pinned: Optional[bool] = None,
                 n_staging: int = 2):
        self.store = store
        self.anchor_data = list(anchor_data)
        self.split = split
        # ring of staging sets: batch k+1 is assembled while the H2D copy of batch k is still queued; a set is
        # only rewritten after the event of its last copy has fired (checked in collate_pairs, not left to the caller)
        pin_ = torch.cuda.is_available() if pinned is None else pinned
        self._stages = [_Staging(pin_) for _ in range(max(2, int(n_staging)))]
        self._stage_next = 0

    def __len__(self):
        return len(self.anchor_data)

    # -- per-sample logic shared by both paths (scan3r.py:59-107)
    def _sample_meta(self, idx: int) -> dict:
        g = self.anchor_data[idx]
        src, ref = self.store[g['src']], self.store[g['ref']]
        if self.split == 'train':       # scan3r.py:69-75 (same global-RNG call as the reference)
            center = src['center'] if np.random.rand(1)[0] > 0.5 else ref['center']
        else:
            center = src['center']
        src_ids, ref_ids = src['objects_id'], ref['objects_id']
        anchors = g['anchorIds'] if 'anchorIds' in g else list(src_ids)
        src_set, ref_set = set(int(v) for v in src_ids), set(int(v) for v in ref_ids)
        anchors = [a for a in anchors if a != 0 and a in src_set and a in ref_set]       # scan3r.py:85-86
        if self.split == 'train':                                                        # scan3r.py:88-90
            cnt = 2 if int(0.3 * len(anchors)) < 1 else int(0.3 * len(anchors))
            anchors = anchors[:cnt]
        s2i = {int(v): i for i, v in enumerate(src_ids)}
        r2i = {int(v): i for i, v in enumerate(ref_ids)}
        aset = set(anchors)
        ns = src['n']
        e1i = np.array([s2i[a] for a in anchors], dtype=np.int64)
        e2i = np.array([r2i[a] for a in anchors], dtype=np.int64) + ns
        e1j = np.array([s2i[int(o)] for o in src_ids if int(o) not in aset], dtype=np.int64)
        e2j = np.array([r2i[int(o)] for o in ref_ids if int(o) not in aset], dtype=np.int64) + ns
        return {'src': src, 'ref': ref, 'center': np.asarray(center), 'e1i': e1i, 'e2i': e2i, 'e1j': e1j, 'e2j': e2j,
                'overlap': g['overlap'] if 'overlap' in g else -1.0, 'scene_ids': [g['src'], g['ref']]}

    def __getitem__(self, idx):
        m = self._sample_meta(idx)
        src, ref = m['src'], m['ref']
        cat = np.concatenate
        f64 = lambda a: torch.from_numpy(np.asarray(a, dtype=np.float64))
        pts = torch.from_numpy(cat([src['obj_points'] - m['center'], ref['obj_points'] - m['center']])).type(torch.FloatTensor)
        d = {
            'obj_ids': cat([src['objects_id'], ref['objects_id']]),
            'tot_obj_pts': pts,
            'graph_per_obj_count': np.array([src['n'], ref['n']]),
            'graph_per_edge_count': np.array([src['e'], ref['e']]),
            'tot_obj_count': src['n'] + ref['n'],
            'tot_bow_vec_object_attr_feats': f64(cat([src['bow_vec_object_attr_feats'], ref['bow_vec_object_attr_feats']])),
            'tot_bow_vec_object_edge_feats': f64(cat([src['bow_vec_object_edge_feats'], ref['bow_vec_object_edge_feats']])),
            'tot_rel_pose': f64(cat([src['rel_trans'], ref['rel_trans']])),
            'edges': torch.from_numpy(cat([src['edges'], ref['edges']]).astype(np.int64)),
            'global_obj_ids': cat([src['objects_cat'], ref['objects_cat']]),
            'scene_ids': m['scene_ids'], 'pcl_center': m['center'], 'overlap': m['overlap'],
        }
        for k in ('e1i', 'e2i', 'e1j', 'e2j'):
            d[k] = m[k]
            d[k + '_count'] = m[k].shape[0]
        return d

    @staticmethod
    def collate_fn(batch):
        """``scan3r.py:142-209``."""
        cat_t = lambda k: torch.cat([b[k] for b in batch])
        out = {'tot_obj_pts': cat_t('tot_obj_pts')}
        prev = 0
        idx = {k: [] for k in ('e1i', 'e2i', 'e1j', 'e2j')}
        for b in batch:
            for k in idx:
                idx[k].append(np.asarray(b[k], dtype=np.int64) + prev)
            prev += b['tot_obj_count']
        for k in idx:
            out[k] = np.concatenate(idx[k]).astype(np.int32)
            out[k + '_count'] = np.stack([b[k + '_count'] for b in batch])
        out['tot_obj_count'] = np.stack([b['tot_obj_count'] for b in batch])
        out['global_obj_ids'] = np.concatenate([b['global_obj_ids'] for b in batch])
        out['tot_bow_vec_object_attr_feats'] = cat_t('tot_bow_vec_object_attr_feats').double()
        out['tot_bow_vec_object_edge_feats'] = cat_t('tot_bow_vec_object_edge_feats').double()
        out['tot_rel_pose'] = cat_t('tot_rel_pose').double()
        out['graph_per_obj_count'] = np.stack([b['graph_per_obj_count'] for b in batch])
        out['graph_per_edge_count'] = np.stack([b['graph_per_edge_count'] for b in batch])
        out['edges'] = cat_t('edges')
        out['scene_ids'] = np.stack([b['scene_ids'] for b in batch])
        out['obj_ids'] = np.concatenate([b['obj_ids'] for b in batch])
        out['pcl_center'] = np.stack([b['pcl_center'] for b in batch])
        out['overlap'] = np.stack([b['overlap'] for b in batch])
        out['batch_size'] = out['overlap'].shape[0]
        return out

    # -- fast path
    def collate_pairs(self, indices: Sequence[int]) -> dict:
        """One batch assembled straight from the mapping into (pinned) staging buffers.  Same dict as
        ``collate_fn([self[i] for i in indices])`` EXCEPT that ``tot_obj_pts`` holds the RAW (un-centred)
        points and ``'_sga_center'`` (f32 ``[B,3]``) + ``'_sga_raw_points'`` ask :func:`to_device` to do
        the centring on the GPU.  The staging buffers come from a ring of ``n_staging`` sets; a set is reused only
        after the H2D copies :func:`to_device` queued from it have completed (event-guarded), so
        ``to_device(ds.collate_pairs(idx))`` is safe with the host running ahead of the GPU.  A caller that reads
        the host tensors directly must be done with batch k before batch k + n_staging is collated."""
        metas = [self._sample_meta(i) for i in indices]
        B = len(metas)
        st = self.store
        P = st.n_points
        n_obj = [m['src']['n'] + m['ref']['n'] for m in metas]
        n_edge = [m['src']['e'] + m['ref']['e'] for m in metas]
        N, E = int(sum(n_obj)), int(sum(n_edge))
        S = self._stages[self._stage_next]
        self._stage_next = (self._stage_next + 1) % len(self._stages)
        S.wait()                                   # the H2D copies of the batch that last used this set
        pts = S.get('pts', (N, P, 3), torch.float32)
        attr = S.get('attr', (N, st.attr_dim), torch.float64)
        rel = S.get('rel', (N, st.rel_dim), torch.float64)
        pose = S.get('pose', (N, 3), torch.float64)
        edges = S.get('edges', (E, 2), torch.int64)
        pts_n, attr_n, rel_n, pose_n, edges_n = pts.numpy(), attr.numpy(), rel.numpy(), pose.numpy(), edges.numpy()
        o = e = 0
        idx = {k: [] for k in ('e1i', 'e2i', 'e1j', 'e2j')}
        obj_ids, gids = [], []
        for m in metas:
            for k in idx:
                idx[k].append(m[k] + o)
            for part in (m['src'], m['ref']):
                n, ne = part['n'], part['e']
                pts_n[o:o + n] = part['obj_points']
                attr_n[o:o + n] = part['bow_vec_object_attr_feats']       # uint8 -> f64 widening copy
                rel_n[o:o + n] = part['bow_vec_object_edge_feats']
                pose_n[o:o + n] = part['rel_trans']
                edges_n[e:e + ne] = part['edges']                         # int32 -> int64 widening copy
                obj_ids.append(part['objects_id'])
                gids.append(part['objects_cat'])
                o += n
                e += ne
        out = {'tot_obj_pts': pts, 'tot_bow_vec_object_attr_feats': attr, 'tot_bow_vec_object_edge_feats': rel,
               'tot_rel_pose': pose, 'edges': edges}
        for k in idx:
            out[k] = (np.concatenate(idx[k]) if idx[k] else np.zeros(0)).astype(np.int32)
            out[k + '_count'] = np.array([len(v) for v in idx[k]])
        out['tot_obj_count'] = np.array(n_obj)
        out['global_obj_ids'] = np.concatenate(gids)
        out['graph_per_obj_count'] = np.array([[m['src']['n'], m['ref']['n']] for m in metas])
        out['graph_per_edge_count'] = np.array([[m['src']['e'], m['ref']['e']] for m in metas])
        out['scene_ids'] = np.stack([m['scene_ids'] for m in metas])
        out['obj_ids'] = np.concatenate(obj_ids)
        out['pcl_center'] = np.stack([m['center'] for m in metas])
        out['overlap'] = np.stack([m['overlap'] for m in metas])
        out['batch_size'] = B
        out['_sga_center'] = torch.from_numpy(np.stack([m['center'] for m in metas]).astype(np.float32))
        out['_sga_raw_points'] = True
        out['_sga_stage'] = S
        return out


def to_device(batch: dict, device, n_chunks: int = 4, keys=None) -> dict:
    """H2D of a :meth:`Scan3RPacked.collate_pairs` batch (chunked, on the copy stream, see
    ``data.to_cuda_streamed``) followed by the on-device centring of the raw points
    (``sga_center_points``: ``pts[o] -= center[pair(o)]``, the subtraction of ``scan3r.py:99-100``)."""
    from . import ops
    from .data import to_cuda_streamed
    raw = batch.get('_sga_raw_points', False)
    host = {k: v for k, v in batch.items() if not k.startswith('_sga_')}
    d = to_cuda_streamed(host, device, n_chunks=n_chunks, keys=keys)
    stage = batch.get('_sga_stage')
    if stage is not None and stage.pinned:
        # every copy that reads the pinned staging set is queued on the copy stream by now: guard the set
        ev = torch.cuda.Event()
        ev.record(d['_sga_ready']['stream'])
        stage.busy = ev
    if raw:
        dev = d['tot_obj_pts'].device
        ready = d.pop('_sga_ready')
        cur = torch.cuda.current_stream(dev)
        for _s, _e, ev in ready['pts']:
            cur.wait_event(ev)
        cur.wait_event(ready['small'])
        lay = ops.PairLayout(np.asarray(batch['graph_per_obj_count']), dev)
        ops.center_points(d['tot_obj_pts'], batch['_sga_center'].to(dev, non_blocking=True), lay.node_pair)
    return d
