"""vcfdist_b200 — B200-native (sm_100a) precision/recall hot path of TimD1/vcfdist.

Only what the path needs lives here: `csrc/` (CUDA kernels + the C-ABI of
include/vcfdist_b200.h), `host/` (the C++ drop-in for the reference's
precision_recall_threads_wrapper) and thin Python plumbing (ctypes binding, batch layout,
synthetic inputs, multi-GPU sharding) used by tests/ and bench.py.
"""
from .batch import Batch, BatchBuilder, Final, Out  # noqa: F401

__all__ = ["Batch", "BatchBuilder", "Final", "Out"]
