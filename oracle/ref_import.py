"""Import the UNMODIFIED reference hot-path modules from ``/root/reference`` (build container
only -- the path does not exist on the GPU box) so the oracle restatement can be pinned against
them and golden vectors generated.  TEST INFRASTRUCTURE ONLY.

Three import-time dependencies of ``src/aligner`` are absent from this image and are stubbed:

* ``torchsummary``     -- imported at ``networks/pointnet.py:11``, never called on the hot path.
* ``pointnet2_ops``    -- imported at ``networks/pct.py:6``; only ``PCT``/``SG`` use it.
* ``torch_geometric.nn`` -- ``networks/gat.py:4`` needs ``GATConv`` (``GCNConv`` for EVA only).
  ``torch-geometric==2.2.0`` (``req.yml:259``) is an un-vendored dependency, so ``GATConv`` is
  a restated ``nn.Module`` here with PyG's parameter names (``lin_src``/``lin_dst`` sharing one
  Linear, ``att_src``, ``att_dst``, ``bias``) and PyG's glorot init.  Its arithmetic delegates
  to :func:`oracle.sgaligner_oracle.gat_conv` -- consequently the GAT branch of the golden
  vectors pins the reference's *call structure* (per-graph slicing, layer stacking, ELU) but
  not PyG's arithmetic: "parity unpinned" at that boundary.
"""
from __future__ import annotations

import math
import os
import sys
import types

import torch
import torch.nn as nn

REFERENCE_ROOT = os.environ.get('SGA_REFERENCE_ROOT', '/root/reference')


class _GATConvStub(nn.Module):
    def __init__(self, in_channels, out_channels, heads=1, concat=True, negative_slope=0.2,
                 dropout=0.0, add_self_loops=True, bias=True, **kw):
        super().__init__()
        assert concat and add_self_loops and bias and dropout == 0.0
        self.in_channels, self.out_channels, self.heads = in_channels, out_channels, heads
        self.negative_slope = negative_slope
        self.lin_src = nn.Linear(in_channels, heads * out_channels, bias=False)
        self.lin_dst = self.lin_src
        self.att_src = nn.Parameter(torch.empty(1, heads, out_channels))
        self.att_dst = nn.Parameter(torch.empty(1, heads, out_channels))
        self.bias = nn.Parameter(torch.zeros(heads * out_channels))
        self.reset_parameters()

    def reset_parameters(self):
        def glorot(t):
            a = math.sqrt(6.0 / (t.size(-2) + t.size(-1)))
            with torch.no_grad():
                t.uniform_(-a, a)
        glorot(self.lin_src.weight)
        glorot(self.att_src)
        glorot(self.att_dst)
        with torch.no_grad():
            self.bias.zero_()

    def forward(self, x, edge_index):
        from oracle.sgaligner_oracle import gat_conv
        return gat_conv(x, edge_index, self.lin_src.weight, self.att_src, self.att_dst, self.bias,
                        self.heads, self.negative_slope)


class _GCNConvStub(nn.Module):
    """PyG 2.2.0 ``GCNConv`` parameter layout (``lin.weight`` [out, in] glorot, no bias inside ``lin``; ``bias`` zeros);
    arithmetic delegated to :func:`oracle.eva_oracle.gcn_conv` (restated, parity unpinned like ``GATConv``)."""

    def __init__(self, in_channels, out_channels, cached=False, **kw):
        super().__init__()
        self.lin = nn.Linear(in_channels, out_channels, bias=False)
        self.bias = nn.Parameter(torch.zeros(out_channels))
        a = math.sqrt(6.0 / (in_channels + out_channels))
        with torch.no_grad():
            self.lin.weight.uniform_(-a, a)

    def forward(self, x, edge_index):
        from oracle.eva_oracle import gcn_conv
        return gcn_conv(x, edge_index, self.lin.weight, self.bias)


def _install_stubs():
    if 'torchsummary' not in sys.modules:
        m = types.ModuleType('torchsummary')
        m.summary = lambda *a, **k: None
        sys.modules['torchsummary'] = m
    if 'pointnet2_ops' not in sys.modules:
        m = types.ModuleType('pointnet2_ops')
        mu = types.ModuleType('pointnet2_ops.pointnet2_utils')
        m.pointnet2_utils = mu
        sys.modules['pointnet2_ops'] = m
        sys.modules['pointnet2_ops.pointnet2_utils'] = mu
    if 'torch_geometric' not in sys.modules:
        tg = types.ModuleType('torch_geometric')
        tgn = types.ModuleType('torch_geometric.nn')
        tgn.GATConv = _GATConvStub
        tgn.GCNConv = _GCNConvStub
        tg.nn = tgn
        sys.modules['torch_geometric'] = tg
        sys.modules['torch_geometric.nn'] = tgn


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'src', 'aligner'))


def load_reference():
    """Returns ``(sg_aligner_module, losses_module, alignment_module)`` of the reference."""
    if not available():
        raise RuntimeError(f'reference not found under {REFERENCE_ROOT}')
    _install_stubs()
    for pth in (os.path.join(REFERENCE_ROOT, 'src'), REFERENCE_ROOT):
        if pth not in sys.path:
            sys.path.insert(0, pth)
    # a product shim named ``aligner`` may already be imported; drop it
    for k in [k for k in sys.modules if k == 'aligner' or k.startswith('aligner.')]:
        if REFERENCE_ROOT not in (getattr(sys.modules[k], '__file__', '') or ''):
            del sys.modules[k]
    import importlib
    sg = importlib.import_module('aligner.sg_aligner')
    ls = importlib.import_module('aligner.losses')
    import importlib.util
    spec = importlib.util.spec_from_file_location('_ref_alignment', os.path.join(REFERENCE_ROOT, 'utils', 'alignment.py'))
    al = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(al)
    return sg, ls, al
