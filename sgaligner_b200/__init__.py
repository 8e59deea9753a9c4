"""sgaligner_b200 -- B200-native (sm_100a) implementation of SGAligner's node-embedding, matching
and contrastive-loss hot path behind the reference's own Python module API.

    from sgaligner_b200.sg_aligner import MultiModalEncoder      # src/aligner/sg_aligner.py
    from sgaligner_b200.losses import OverallLoss, CustomMultiLossLayer   # src/aligner/losses.py
    from sgaligner_b200 import matching                            # inference matching head + utils/alignment.py

Importing the package does not need a GPU; running the hot path does (there is no CPU fallback).
"""
from .data import to_cuda  # noqa: F401

__version__ = '0.1.0'
