"""TEST INFRASTRUCTURE - loaders of the checker libraries under oracle/_ref/.

  liboracle.so          the plain-C restatement of the path (oracle/vd_oracle.c)
  libvdref.so / B       the reference's own object code behind a function-level harness
                        (oracle/ref_harness.cpp; B = with the canonical tie-break patch)
  libvdseam.so          that harness with the hot-path call resolved to the GPU drop-in (seam timing and parity of
                        vcfdist_b200/host/pr_dropin.cpp; here the harness is the test bench, the drop-in the product)

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module;
it is the checker, never the thing measured or shipped.  The product path
(vcfdist_b200.capi.Engine) does not know it exists.
"""
from __future__ import annotations

import ctypes as C
import os
import sys
from typing import Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))

from vcfdist_b200.batch import Batch, Out, vd_batch_in, vd_batch_out  # noqa: E402
from vcfdist_b200.capi import VdError  # noqa: E402

ORACLE_DIR = os.path.join(_HERE, "_ref")


def load_oracle() -> C.CDLL:
    p = os.path.join(ORACLE_DIR, "liboracle.so")
    if not os.path.exists(p):
        raise RuntimeError(f"{p} missing: run `make -C oracle port`")
    lib = C.CDLL(p)
    lib.vdo_run.argtypes = [C.POINTER(vd_batch_in), C.POINTER(vd_batch_out)]
    lib.vdo_run.restype = C.c_int
    return lib


def oracle_run(batch: Batch) -> Out:
    lib = load_oracle()
    out = Out(batch.n_sc, batch.n_var)
    cin, cout = batch.as_c(), out.as_c()
    rc = lib.vdo_run(C.byref(cin), C.byref(cout))
    if rc != 0:
        raise VdError(rc, "oracle rejected the batch")
    return out


class vdref_out(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in
                ("errtypes", "sync_group", "ref_ed", "query_ed", "callq", "credit",
                 "sc_phase", "orig_dist", "swap_dist")]


def reference_available(canonical: bool = False) -> bool:
    return os.path.exists(os.path.join(ORACLE_DIR, "libvdrefB.so" if canonical else "libvdref.so"))


_ref_libs = {}


def seam_available() -> bool:
    return os.path.exists(os.path.join(ORACLE_DIR, "libvdseam.so"))


def seam_breakdown() -> dict:
    """Host-side times (ms) of the drop-in's last call through libvdseam.so (reference_run(..., seam=True))."""
    lib = _ref_libs["libvdseam.so"]
    a = (C.c_double * 10)()
    lib.vd_dropin_last_times(a)
    keys = ("wait_for_handle", "pack", "vd_run_wall", "vd_run_device", "status_scan", "float_step", "scatter")
    d = {k: float(a[i]) for i, k in enumerate(keys)}
    d.update(host_threads=int(a[7]), page_locked=bool(a[8]), tie_superclusters=int(a[9]))
    return d


def reference_run(batch: Batch, canonical: bool = False, threads: int = 1, max_ram: float = 64.0,
                  phase_threshold: float = 0.6, credit_threshold: float = 0.7,
                  max_qual: int = 60, seam: bool = False) -> Tuple[dict, float]:
    """Run the REFERENCE's own object code (oracle/_ref/libvdref[B].so) on the batch.
    Returns (results dict like Final.trimmed(), seconds inside precision_recall_threads_wrapper).
    seam=True: the same harness and reference data structures with precision_recall_threads_wrapper resolved to the
    GPU drop-in (libvdseam.so) - the product measured where the reference's own timer sits (needs a GPU)."""
    name = "libvdseam.so" if seam else "libvdrefB.so" if canonical else "libvdref.so"
    if name not in _ref_libs:
        lib = C.CDLL(os.path.join(ORACLE_DIR, name))
        lib.vdref_run.argtypes = [C.POINTER(vd_batch_in), C.POINTER(vdref_out), C.c_int, C.c_double,
                                  C.c_double, C.c_double, C.c_int, C.POINTER(C.c_double)]
        lib.vdref_run.restype = C.c_int
        _ref_libs[name] = lib
    lib = _ref_libs[name]
    v, s = max(2 * batch.n_var, 1), max(batch.n_sc, 1)
    arrs = dict(errtypes=np.zeros(v, np.uint8), sync_group=np.zeros(v, np.int32),
                ref_ed=np.zeros(v, np.int32), query_ed=np.zeros(v, np.int32),
                callq=np.zeros(v, np.float32), credit=np.zeros(v, np.float32),
                sc_phase=np.zeros(s, np.int32), orig_dist=np.zeros(s, np.int32),
                swap_dist=np.zeros(s, np.int32))
    ro = vdref_out()
    for k, a in arrs.items():
        setattr(ro, k, a.ctypes.data)
    sec = C.c_double(0)
    cin = batch.as_c()
    rc = lib.vdref_run(C.byref(cin), C.byref(ro), threads, max_ram, phase_threshold, credit_threshold,
                       max_qual, C.byref(sec))
    if rc != 0:
        raise VdError(rc, "reference harness")
    nv2 = 2 * batch.n_var
    res = {k: (a[:nv2] if k not in ("sc_phase", "orig_dist", "swap_dist") else a[: batch.n_sc])
           for k, a in arrs.items()}
    return res, sec.value


# --------------------------------------------------------------------------------------
# SURVEY 8f-1 groundwork: wf_swg_max_reach (src/dist.cpp:2150-2333)
# --------------------------------------------------------------------------------------
_REACH_ARGS = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]


def reach_oracle(query: bytes, truth: bytes, main_diag: int, main_diag_start: int, max_score: int,
                 sub: int, open_: int, extend: int, reverse: bool) -> int:
    """The C restatement (oracle/vd_reach.c)."""
    lib = load_oracle()
    lib.vdo_max_reach.argtypes = _REACH_ARGS
    lib.vdo_max_reach.restype = C.c_int
    return lib.vdo_max_reach(query, len(query), truth, len(truth), main_diag, main_diag_start, max_score,
                             sub, open_, extend, int(reverse))


def reach_reference(query: bytes, truth: bytes, main_diag: int, main_diag_start: int, max_score: int,
                    sub: int, open_: int, extend: int, reverse: bool) -> int:
    """The reference's own object code (oracle/ref_harness.cpp: vdref_max_reach)."""
    key = "reach:libvdref.so"
    if key not in _ref_libs:
        _ref_libs[key] = C.CDLL(os.path.join(ORACLE_DIR, "libvdref.so"))
    lib = _ref_libs[key]
    lib.vdref_max_reach.argtypes = _REACH_ARGS
    lib.vdref_max_reach.restype = C.c_int
    return lib.vdref_max_reach(query, len(query), truth, len(truth), main_diag, main_diag_start, max_score,
                               sub, open_, extend, int(reverse))


_SWG_ARGS = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int]


def swg_score_oracle(query: bytes, truth: bytes, sub: int, open_: int, extend: int) -> int:
    """Affine-gap alignment score, C restatement of wf_swg_align's score (oracle/vd_reach.c)."""
    lib = load_oracle()
    lib.vdo_swg_score.argtypes = _SWG_ARGS
    lib.vdo_swg_score.restype = C.c_int
    return lib.vdo_swg_score(query, len(query), truth, len(truth), sub, open_, extend)


def swg_score_reference(query: bytes, truth: bytes, sub: int, open_: int, extend: int) -> int:
    """The reference's own wf_swg_align (oracle/ref_harness.cpp: vdref_swg_score)."""
    key = "reach:libvdref.so"
    if key not in _ref_libs:
        _ref_libs[key] = C.CDLL(os.path.join(ORACLE_DIR, "libvdref.so"))
    lib = _ref_libs[key]
    lib.vdref_swg_score.argtypes = _SWG_ARGS
    lib.vdref_swg_score.restype = C.c_int
    return lib.vdref_swg_score(query, len(query), truth, len(truth), sub, open_, extend)


def cluster_reference(fasta: bytes, var, sub: int, open_: int, extend: int, max_iters: int = 4, reach_min_gap: int = 10):
    """The reference's own wf_swg_cluster (oracle/ref_harness.cpp: vdref_cluster) on one haplotype.
    var[i] = (pos, rlen, type, alt bytes), sorted.  -> (clusters, left_reaches, right_reaches)."""
    key = "reach:libvdref.so"
    if key not in _ref_libs:
        _ref_libs[key] = C.CDLL(os.path.join(ORACLE_DIR, "libvdref.so"))
    lib = _ref_libs[key]
    n = len(var)
    pos = np.array([v[0] for v in var], np.int32)
    rlen = np.array([v[1] for v in var], np.int32)
    typ = np.array([v[2] for v in var], np.uint8)
    alt_off = np.zeros(n + 1, np.int64)
    alt_off[1:] = np.cumsum([len(v[3]) for v in var])
    alt = b"".join(v[3] for v in var) or b"\0"
    out = [np.zeros(n + 1, np.int32) for _ in range(3)]
    lib.vdref_cluster.restype = C.c_int
    lib.vdref_cluster.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_char_p,
                                  C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    nc = lib.vdref_cluster(fasta, len(fasta), n, pos.ctypes.data, rlen.ctypes.data, typ.ctypes.data, alt_off.ctypes.data, alt,
                           sub, open_, extend, max_iters, reach_min_gap, out[0].ctypes.data, out[1].ctypes.data, out[2].ctypes.data)
    return [list(map(int, o[:nc])) for o in out]


def swg_cigar_oracle(query: bytes, truth: bytes, sub: int, open_: int, extend: int):
    """Affine-gap alignment + walk back, C restatement (oracle/vd_reach.c: vdo_swg_cigar) -> (score, cigar)."""
    lib = load_oracle()
    lib.vdo_swg_cigar.argtypes = _SWG_ARGS + [C.c_void_p]
    lib.vdo_swg_cigar.restype = C.c_int
    cig = np.zeros(len(query) + len(truth), np.int32)
    s = lib.vdo_swg_cigar(query, len(query), truth, len(truth), sub, open_, extend, cig.ctypes.data)
    return s, cig


def swg_cigar_reference(query: bytes, truth: bytes, sub: int, open_: int, extend: int):
    """The reference's wf_swg_align + wf_swg_backtrack (oracle/ref_harness.cpp: vdref_swg_cigar)."""
    key = "reach:libvdref.so"
    if key not in _ref_libs:
        _ref_libs[key] = C.CDLL(os.path.join(ORACLE_DIR, "libvdref.so"))
    lib = _ref_libs[key]
    lib.vdref_swg_cigar.argtypes = _SWG_ARGS + [C.c_void_p]
    lib.vdref_swg_cigar.restype = C.c_int
    cig = np.zeros(len(query) + len(truth), np.int32)
    s = lib.vdref_swg_cigar(query, len(query), truth, len(truth), sub, open_, extend, cig.ctypes.data)
    return s, cig
