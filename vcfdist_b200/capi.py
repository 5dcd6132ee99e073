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

from .batch import Batch, Final, Out, vd_batch_in, vd_batch_out, vd_final, vd_stats

_ROOT = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_ROOT, "libvcfdist_b200.so")

EXPORTS = ("vd_abi_version", "vd_create", "vd_destroy", "vd_run", "vd_run_device", "vd_run_device_slice",
           "vd_finalize", "vd_get_stats", "vd_last_error", "vd_stream")

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
    if path is None:
        _lib = lib
    return lib


class VdError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"vcfdist_b200 error {code}: {msg}")
        self.code = code


def finalize(batch: Batch, out: Out, phase_threshold: float = 0.6, credit_threshold: float = 0.7) -> Final:
    """Host float step (store_phase + credit thresholds); needs no GPU."""
    lib = load_library()
    fin = Final(batch.n_sc, batch.n_var)
    cin, cout, cfin = batch.as_c(), out.as_c(), fin.as_c()
    rc = lib.vd_finalize(C.byref(cin), C.byref(cout), phase_threshold, credit_threshold, C.byref(cfin))
    if rc != 0:
        raise VdError(rc, "vd_finalize")
    return fin


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

    def run_device(self, din: vd_batch_in, dout: vd_batch_out, n_var: int, ref_bytes: int, alt_bytes: int):
        """All pointers already resident in this GPU's HBM."""
        self._check(self.lib.vd_run_device(self.h, C.byref(din), C.byref(dout), n_var, ref_bytes, alt_bytes))

    def run_device_slice(self, din: vd_batch_in, dout: vd_batch_out, first_var: int, n_var: int,
                         ref_bytes: int, alt_bytes: int):
        """A slice of a resident batch (vd_run_device_slice): ref_off / var_off advanced to the slice."""
        self._check(self.lib.vd_run_device_slice(self.h, C.byref(din), C.byref(dout), first_var, n_var,
                                                 ref_bytes, alt_bytes))

    def stats(self) -> dict:
        st = vd_stats()
        self.lib.vd_get_stats(self.h, C.byref(st))
        return st.as_dict()

    @property
    def stream(self) -> int:
        return int(self.lib.vd_stream(self.h) or 0)
