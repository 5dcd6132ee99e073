"""ctypes bindings of the C-ABI (include/vcfdist_b200.h).

`Engine` is the product path: it loads vcfdist_b200/libvcfdist_b200.so (hand-written
sm_100a CUDA behind `extern "C"`) and fails loudly when the library or a GPU is missing —
there is no CPU fallback.  The checker libraries under oracle/_ref/ have their own
loader, `oracle/checkers.py` (test infrastructure: tests/, __graft_entry__.smoke() and
bench.py's CPU-baseline legs only); nothing in this package touches them.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Tuple

import numpy as np

from .batch import Batch, CompactIn, Final, Out, PackedOut, vd_batch_in, vd_batch_out, vd_compact_in, vd_final, vd_packed_out, vd_stats

_ROOT = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_ROOT, "libvcfdist_b200.so")

EXPORTS = ("vd_abi_version", "vd_create", "vd_destroy", "vd_run", "vd_run_device", "vd_run_device_slice",
           "vd_finalize", "vd_get_stats", "vd_last_error", "vd_stream", "vd_wf_batch", "vd_swg_align_batch",
           "vd_run_packed", "vd_finalize_packed", "vd_host_alloc", "vd_host_free", "vd_pack_device", "vd_packed_overflow",
           "vd_compact_pack", "vd_run_compact")

_lib = None


def load_library(path: Optional[str] = None) -> C.CDLL:
    """Load the product library; raises if it has not been built (see __graft_entry__.build)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("VD_LIB") or LIB_PATH      # VD_LIB: another build of the same library (kernel experiments)
    if not os.path.exists(p):
        raise RuntimeError(
            f"{p} is missing: the CUDA extension has not been built "
            "(run `python -c 'import __graft_entry__ as g; g.build()'`). "
            "vcfdist_b200 has no CPU fallback.")
    lib = C.CDLL(p)
    lib.vd_abi_version.restype = C.c_int
    lib.vd_create.argtypes = [C.c_int, C.c_int64, C.POINTER(C.c_void_p)]
    lib.vd_create.restype = C.c_int
    lib.vd_destroy.argtypes = [C.c_void_p]
    lib.vd_destroy.restype = None
    lib.vd_run.argtypes = [C.c_void_p, C.POINTER(vd_batch_in), C.POINTER(vd_batch_out)]
    lib.vd_run.restype = C.c_int
    lib.vd_run_packed.argtypes = [C.c_void_p, C.POINTER(vd_batch_in), C.POINTER(vd_packed_out)]
    lib.vd_run_packed.restype = C.c_int
    lib.vd_finalize_packed.argtypes = [C.POINTER(vd_batch_in), C.POINTER(vd_packed_out), C.c_double, C.c_double, C.POINTER(vd_final)]
    lib.vd_finalize_packed.restype = C.c_int
    lib.vd_host_alloc.argtypes = [C.c_int64]
    lib.vd_host_alloc.restype = C.c_void_p
    lib.vd_host_free.argtypes = [C.c_void_p]
    lib.vd_host_free.restype = None
    lib.vd_compact_pack.argtypes = [C.POINTER(vd_batch_in), C.POINTER(vd_compact_in)]
    lib.vd_compact_pack.restype = C.c_int
    lib.vd_run_compact.argtypes = [C.c_void_p, C.POINTER(vd_compact_in), C.POINTER(vd_packed_out)]
    lib.vd_run_compact.restype = C.c_int
    lib.vd_pack_device.argtypes = [C.c_void_p, C.POINTER(vd_batch_out), C.c_int64, C.c_int64, C.POINTER(vd_packed_out)]
    lib.vd_pack_device.restype = C.c_int
    lib.vd_packed_overflow.argtypes = [C.c_void_p]
    lib.vd_packed_overflow.restype = C.c_int
    lib.vd_run_device.argtypes = [C.c_void_p, C.POINTER(vd_batch_in), C.POINTER(vd_batch_out),
                                  C.c_int64, C.c_int64, C.c_int64]
    lib.vd_run_device.restype = C.c_int
    lib.vd_run_device_slice.argtypes = [C.c_void_p, C.POINTER(vd_batch_in), C.POINTER(vd_batch_out),
                                        C.c_int64, C.c_int64, C.c_int64, C.c_int64]
    lib.vd_run_device_slice.restype = C.c_int
    lib.vd_finalize.argtypes = [C.POINTER(vd_batch_in), C.POINTER(vd_batch_out), C.c_double, C.c_double,
                                C.POINTER(vd_final)]
    lib.vd_finalize.restype = C.c_int
    lib.vd_get_stats.argtypes = [C.c_void_p, C.POINTER(vd_stats)]
    lib.vd_get_stats.restype = C.c_int
    lib.vd_last_error.argtypes = [C.c_void_p]
    lib.vd_last_error.restype = C.c_char_p
    lib.vd_stream.argtypes = [C.c_void_p]
    lib.vd_stream.restype = C.c_void_p
    lib.vd_wf_batch.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 8 + [C.c_int, C.c_int, C.c_int, C.c_void_p]
    lib.vd_wf_batch.restype = C.c_int
    lib.vd_swg_align_batch.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 4 + [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    lib.vd_swg_align_batch.restype = C.c_int
    if path is None:
        _lib = lib
    return lib


class VdError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"vcfdist_b200 error {code}: {msg}")
        self.code = code


def finalize(batch: Batch, out, phase_threshold: float = 0.6, credit_threshold: float = 0.7, lib=None) -> Final:
    """Host float step (store_phase + credit thresholds) over wide (Out) or 16-bit (PackedOut) records; needs no GPU."""
    lib = lib or load_library()
    fin = Final(batch.n_sc, batch.n_var)
    cin, cout, cfin = batch.as_c(), out.as_c(), fin.as_c()
    fn = lib.vd_finalize_packed if isinstance(out, PackedOut) else lib.vd_finalize
    rc = fn(C.byref(cin), C.byref(cout), phase_threshold, credit_threshold, C.byref(cfin))
    if rc != 0:
        raise VdError(rc, "vd_finalize")
    return fin


def compact(batch: Batch, lib=None) -> CompactIn:
    """vd_compact_pack: the batch in compact form (host only; raises VdError -7 when a value does not fit 16 / 8 bits)."""
    lib = lib or load_library()
    ci = CompactIn(batch)
    ci.refresh_pointers()
    cin = batch.as_c()
    rc = lib.vd_compact_pack(C.byref(cin), C.byref(ci.c))
    if rc != 0:
        raise VdError(rc, "vd_compact_pack: a value does not fit the compact form")
    return ci


class Engine:
    """One handle per GPU (vd_create / vd_run / vd_destroy)."""

    def __init__(self, device: int = 0, scratch_bytes: int = 0):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.vd_create(device, scratch_bytes, C.byref(h))
        if rc != 0:
            raise VdError(rc, "vd_create failed: no usable CUDA device (there is no CPU fallback)")
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.lib.vd_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, allow_align: bool = True):
        if rc != 0 and not (allow_align and rc == -6):
            raise VdError(rc, (self.lib.vd_last_error(self.h) or b"").decode())

    def run(self, batch: Batch, out: Optional[Out] = None) -> Out:
        """Host buffers in, host buffers out (H2D, kernels, D2H inside)."""
        out = out or Out(batch.n_sc, batch.n_var)
        cin, cout = batch.as_c(), out.as_c()
        self._check(self.lib.vd_run(self.h, C.byref(cin), C.byref(cout)))
        return out

    def run_packed(self, batch: Batch, out: Optional[PackedOut] = None) -> PackedOut:
        """vd_run_packed: the same results in 16-bit records (raises VdError -7 when a value does not fit)."""
        out = out or PackedOut(batch.n_sc, batch.n_var)
        cin, cout = batch.as_c(), out.as_c()
        self._check(self.lib.vd_run_packed(self.h, C.byref(cin), C.byref(cout)))
        return out

    def run_compact(self, ci: CompactIn, out: Optional[PackedOut] = None) -> PackedOut:
        """vd_run_compact: compact batch in (less than half the bytes over PCIe), 16-bit records out."""
        out = out or PackedOut(ci.batch.n_sc, ci.batch.n_var)
        cout = out.as_c()
        self._check(self.lib.vd_run_compact(self.h, C.byref(ci.c), C.byref(cout)))
        return out

    def pack_device(self, dwide: vd_batch_out, n_sc: int, n_var: int, dpacked: vd_packed_out):
        """Narrow device-resident wide records into 16-bit ones on the handle's stream (asynchronous)."""
        self._check(self.lib.vd_pack_device(self.h, C.byref(dwide), n_sc, n_var, C.byref(dpacked)))

    def packed_overflow(self) -> bool:
        return bool(self.lib.vd_packed_overflow(self.h))

    def run_device(self, din: vd_batch_in, dout: vd_batch_out, n_var: int, ref_bytes: int, alt_bytes: int):
        """All pointers already resident in this GPU's HBM."""
        self._check(self.lib.vd_run_device(self.h, C.byref(din), C.byref(dout), n_var, ref_bytes, alt_bytes))

    def run_device_slice(self, din: vd_batch_in, dout: vd_batch_out, first_var: int, n_var: int,
                         ref_bytes: int, alt_bytes: int):
        """A slice of a resident batch (vd_run_device_slice): ref_off / var_off advanced to the slice."""
        self._check(self.lib.vd_run_device_slice(self.h, C.byref(din), C.byref(dout), first_var, n_var,
                                                 ref_bytes, alt_bytes))

    def wf_batch(self, mode: int, queries, truths, sub: int, open_: int, extend: int, main_diag=None, main_diag_start=None,
                 max_score=None, reverse=None) -> np.ndarray:
        """Batch of affine-gap wavefront problems of the cluster-growing stage (vd_wf_batch): mode 0 =
        wf_swg_max_reach (needs main_diag, main_diag_start, max_score, reverse per problem), mode 1 = score of
        wf_swg_align.  queries / truths: sequences of bytes."""
        n = len(queries)
        q_off = np.zeros(n + 1, np.int64); q_off[1:] = np.cumsum([len(q) for q in queries])
        t_off = np.zeros(n + 1, np.int64); t_off[1:] = np.cumsum([len(t) for t in truths])
        q_seq = np.frombuffer(b"".join(queries) or b"\0", np.uint8)
        t_seq = np.frombuffer(b"".join(truths) or b"\0", np.uint8)
        res = np.zeros(max(n, 1), np.int32)
        arr = lambda a, dt: None if a is None else np.ascontiguousarray(a, dt)
        md, mds, ms, rv = arr(main_diag, np.int32), arr(main_diag_start, np.int32), arr(max_score, np.int32), arr(reverse, np.uint8)
        p = lambda a: None if a is None else a.ctypes.data
        self._check(self.lib.vd_wf_batch(self.h, mode, n, p(q_off), p(q_seq), p(t_off), p(t_seq), p(md), p(mds), p(ms), p(rv),
                                         sub, open_, extend, p(res)), allow_align=False)
        return res[:n]

    def swg_align_batch(self, queries, truths, sub: int, open_: int, extend: int):
        """`--distance` alignments (vd_swg_align_batch): -> (scores [n], list of CIGAR arrays as the reference fills them)."""
        n = len(queries)
        q_off = np.zeros(n + 1, np.int64); q_off[1:] = np.cumsum([len(q) for q in queries])
        t_off = np.zeros(n + 1, np.int64); t_off[1:] = np.cumsum([len(t) for t in truths])
        q_seq = np.frombuffer(b"".join(queries) or b"\0", np.uint8)
        t_seq = np.frombuffer(b"".join(truths) or b"\0", np.uint8)
        score = np.zeros(max(n, 1), np.int32)
        cig = np.zeros(max(int(q_off[n] + t_off[n]), 1), np.int32)
        self._check(self.lib.vd_swg_align_batch(self.h, n, q_off.ctypes.data, q_seq.ctypes.data, t_off.ctypes.data, t_seq.ctypes.data,
                                                sub, open_, extend, score.ctypes.data, cig.ctypes.data), allow_align=False)
        co = q_off + t_off
        return score[:n], [cig[co[i]: co[i + 1]] for i in range(n)]

    def stats(self) -> dict:
        st = vd_stats()
        self.lib.vd_get_stats(self.h, C.byref(st))
        return st.as_dict()

    @property
    def stream(self) -> int:
        return int(self.lib.vd_stream(self.h) or 0)
