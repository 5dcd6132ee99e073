"""Batches of superclusters in the compact boundary layout of include/vcfdist_b200.h.

A `Batch` mirrors `vd_batch_in`: the reference window of every supercluster
(fasta[ctg][begs..ends], src/dist.cpp:163,232) plus the four haplotypes' variant lists
in the order q1, q2, t1, t2 (src/dist.cpp:1786-1822).  All arrays are numpy, laid out
exactly as the C-ABI expects, so the ctypes view is zero-copy.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

TYPE_SUB, TYPE_INS, TYPE_DEL = 1, 2, 3          # src/defs.h:31-35
ERRTYPE_TP, ERRTYPE_FP, ERRTYPE_FN, ERRTYPE_UN = 0, 1, 2, 5   # src/defs.h:66-72
PHASE_ORIG, PHASE_SWAP, PHASE_NONE = 0, 1, 2    # src/defs.h:131-134

ST_TIE = 0x0001
ST_WARN_MASK = 0x001E
ST_ERR_MASK = 0xFF00


class vd_batch_in(C.Structure):
    _fields_ = [
        ("n_sc", C.c_int32),
        ("ref_off", C.c_void_p),
        ("ref_seq", C.c_void_p),
        ("rplane_seq", C.c_void_p),
        ("var_off", C.c_void_p),
        ("var_pos", C.c_void_p),
        ("var_rlen", C.c_void_p),
        ("var_type", C.c_void_p),
        ("alt_off", C.c_void_p),
        ("alt_seq", C.c_void_p),
        ("var_qual", C.c_void_p),
        ("max_qual", C.c_float),
    ]


class vd_packed_out(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("aln_score", "aln_planes", "status", "sync_group", "ref_ed", "query_ed", "callq")]


class vd_compact_in(C.Structure):
    _fields_ = [("n_sc", C.c_int32), ("max_qual", C.c_float), ("n_var", C.c_int64), ("ref_bytes", C.c_int64), ("alt_bytes", C.c_int64)] + \
               [(n, C.c_void_p) for n in ("ref_len", "ref_seq", "rplane_seq", "hap_nvar", "var_pos", "var_rlen", "alt_len", "var_type",
                                          "alt_seq", "var_qual", "blk_var", "blk_ref", "blk_alt")]


COMPACT_BLOCK = 65536


class vd_batch_out(C.Structure):
    _fields_ = [
        ("aln_score", C.c_void_p),
        ("aln_end_plane", C.c_void_p),
        ("aln_beg_plane", C.c_void_p),
        ("status", C.c_void_p),
        ("assigned", C.c_void_p),
        ("sync_group", C.c_void_p),
        ("ref_ed", C.c_void_p),
        ("query_ed", C.c_void_p),
        ("callq", C.c_void_p),
    ]


class vd_final(C.Structure):
    _fields_ = [
        ("errtypes", C.c_void_p),
        ("credit", C.c_void_p),
        ("callq", C.c_void_p),
        ("sync_group", C.c_void_p),
        ("ref_ed", C.c_void_p),
        ("query_ed", C.c_void_p),
        ("sc_phase", C.c_void_p),
        ("orig_dist", C.c_void_p),
        ("swap_dist", C.c_void_p),
    ]


class vd_stats(C.Structure):
    _fields_ = [
        ("n_sc", C.c_int64), ("n_var", C.c_int64), ("cells", C.c_int64),
        ("io_bytes", C.c_int64), ("spill_bytes", C.c_int64),
        ("n_short", C.c_int64), ("n_long", C.c_int64), ("n_launches", C.c_int64),
        ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64),
        ("ms_total", C.c_float), ("ms_short", C.c_float),
        ("ms_long_fwd", C.c_float), ("ms_long_bwd", C.c_float), ("ms_long_walk", C.c_float),
        ("ms_plan", C.c_float), ("ms_long_wall", C.c_float), ("ms_small", C.c_float * 3),
        ("n_small", C.c_int64 * 3), ("io_small", C.c_int64 * 3), ("n_hom", C.c_int64),
        ("n_dense", C.c_int64), ("ms_band", C.c_float), ("pad_", C.c_float),
        ("band_cells", C.c_int64), ("band_rows", C.c_int64), ("band_cols", C.c_int64),
    ]

    def as_dict(self):
        return {k: (list(getattr(self, k)) if k in ("ms_small", "n_small", "io_small") else getattr(self, k))
                for k, _ in self._fields_}


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data


@dataclass
class Batch:
    ref_off: np.ndarray            # int64 [n_sc+1]
    ref_seq: np.ndarray            # uint8
    var_off: np.ndarray            # int64 [4*n_sc+1]
    var_pos: np.ndarray            # int32 [n_var]
    var_rlen: np.ndarray           # int32 [n_var]
    var_type: np.ndarray           # uint8 [n_var]
    alt_off: np.ndarray            # int64 [n_var+1]
    alt_seq: np.ndarray            # uint8
    var_qual: np.ndarray           # float32 [n_var]
    max_qual: float = 60.0
    rplane_seq: Optional[np.ndarray] = None

    def __post_init__(self):
        self.ref_off = np.ascontiguousarray(self.ref_off, np.int64)
        self.ref_seq = np.ascontiguousarray(self.ref_seq, np.uint8)
        self.var_off = np.ascontiguousarray(self.var_off, np.int64)
        self.var_pos = np.ascontiguousarray(self.var_pos, np.int32)
        self.var_rlen = np.ascontiguousarray(self.var_rlen, np.int32)
        self.var_type = np.ascontiguousarray(self.var_type, np.uint8)
        self.alt_off = np.ascontiguousarray(self.alt_off, np.int64)
        self.alt_seq = np.ascontiguousarray(self.alt_seq, np.uint8)
        self.var_qual = np.ascontiguousarray(self.var_qual, np.float32)
        if self.rplane_seq is not None:
            self.rplane_seq = np.ascontiguousarray(self.rplane_seq, np.uint8)
        # keep one spare element so that pointers of empty arrays are still valid
        for name in ("ref_seq", "alt_seq", "var_pos", "var_rlen", "var_type", "var_qual"):
            a = getattr(self, name)
            if a.size == 0:
                setattr(self, name, np.zeros(1, a.dtype))

    # ---- sizes -------------------------------------------------------------------
    @property
    def n_sc(self) -> int:
        return len(self.ref_off) - 1

    @property
    def n_var(self) -> int:
        return int(self.var_off[-1])

    @property
    def ref_bytes(self) -> int:
        return int(self.ref_off[-1])

    @property
    def alt_bytes(self) -> int:
        return int(self.alt_off[-1])

    def window_len(self) -> np.ndarray:
        return np.diff(self.ref_off).astype(np.int64)

    def hap_len(self) -> np.ndarray:
        """[n_sc, 4] haplotype string lengths after applying the variants."""
        n_var = self.n_var
        delta = (np.diff(self.alt_off)[:n_var] - self.var_rlen[:n_var]).astype(np.int64)
        csum = np.concatenate([[0], np.cumsum(delta)])
        per = csum[self.var_off[1:]] - csum[self.var_off[:-1]]
        return self.window_len()[:, None] + per.reshape(self.n_sc, 4)

    def cells(self) -> np.ndarray:
        """[n_sc] DP cells per supercluster: sum over the 4 alignments of (Lq+Lr)*Lt,
        the eight flag matrices of src/dist.cpp:1828-1844."""
        L = self.hap_len()
        lr = self.window_len()
        lq = L[:, [0, 0, 1, 1]]
        lt = L[:, [2, 3, 2, 3]]
        return ((lq + lr[:, None]) * lt).sum(axis=1)

    def io_bytes(self) -> int:
        """Algorithmic input + output bytes of one pass (compact layout)."""
        n_sc, n_var = self.n_sc, self.n_var
        inp = (self.ref_bytes + 8 * (n_sc + 1) + 8 * (4 * n_sc + 1) + n_var * (4 + 4 + 1 + 4)
               + 8 * (n_var + 1) + self.alt_bytes)
        out = 4 * n_sc * (4 + 1 + 1 + 4) + 2 * n_var * (1 + 4 + 4 + 4 + 4)
        return int(inp + out)

    # ---- C view ------------------------------------------------------------------
    def as_c(self) -> vd_batch_in:
        s = vd_batch_in()
        s.n_sc = self.n_sc
        s.ref_off = _ptr(self.ref_off)
        s.ref_seq = _ptr(self.ref_seq)
        s.rplane_seq = _ptr(self.rplane_seq)
        s.var_off = _ptr(self.var_off)
        s.var_pos = _ptr(self.var_pos)
        s.var_rlen = _ptr(self.var_rlen)
        s.var_type = _ptr(self.var_type)
        s.alt_off = _ptr(self.alt_off)
        s.alt_seq = _ptr(self.alt_seq)
        s.var_qual = _ptr(self.var_qual)
        s.max_qual = float(self.max_qual)
        s._keep = self
        return s

    # ---- (de)serialisation ---------------------------------------------------------
    def save(self, path: str) -> None:
        d = dict(ref_off=self.ref_off, ref_seq=self.ref_seq[: self.ref_bytes], var_off=self.var_off,
                 var_pos=self.var_pos[: self.n_var], var_rlen=self.var_rlen[: self.n_var],
                 var_type=self.var_type[: self.n_var], alt_off=self.alt_off,
                 alt_seq=self.alt_seq[: self.alt_bytes], var_qual=self.var_qual[: self.n_var],
                 max_qual=np.float32(self.max_qual))
        if self.rplane_seq is not None:
            d["rplane_seq"] = self.rplane_seq
        np.savez_compressed(path, **d)

    @staticmethod
    def load(path: str) -> "Batch":
        z = np.load(path)
        return Batch(ref_off=z["ref_off"], ref_seq=z["ref_seq"], var_off=z["var_off"],
                     var_pos=z["var_pos"], var_rlen=z["var_rlen"], var_type=z["var_type"],
                     alt_off=z["alt_off"], alt_seq=z["alt_seq"], var_qual=z["var_qual"],
                     max_qual=float(z["max_qual"]),
                     rplane_seq=z["rplane_seq"] if "rplane_seq" in z.files else None)

    # ---- slicing (sharding across GPUs) --------------------------------------------
    def take(self, idx: Sequence[int]) -> "Batch":
        """Sub-batch with the superclusters `idx`, in that order."""
        idx = np.asarray(idx, np.int64)
        n_var = self.n_var
        wl = self.window_len()[idx]
        ref_off = np.concatenate([[0], np.cumsum(wl)])
        ref_seq = _gather_ranges(self.ref_seq, self.ref_off[idx], wl)
        rplane = None if self.rplane_seq is None else _gather_ranges(self.rplane_seq, self.ref_off[idx], wl)
        vo = self.var_off
        hb = (idx[:, None] * 4 + np.arange(4)[None, :]).reshape(-1)
        cnt = vo[hb + 1] - vo[hb]
        var_off = np.concatenate([[0], np.cumsum(cnt)])
        vsel = _range_index(vo[hb], cnt)
        al = np.diff(self.alt_off)[:n_var][vsel] if n_var else np.zeros(0, np.int64)
        alt_off = np.concatenate([[0], np.cumsum(al)])
        alt_seq = _gather_ranges(self.alt_seq, self.alt_off[:-1][vsel], al) if n_var else np.zeros(0, np.uint8)
        return Batch(ref_off=ref_off, ref_seq=ref_seq, var_off=var_off,
                     var_pos=self.var_pos[:n_var][vsel], var_rlen=self.var_rlen[:n_var][vsel],
                     var_type=self.var_type[:n_var][vsel], alt_off=alt_off, alt_seq=alt_seq,
                     var_qual=self.var_qual[:n_var][vsel], max_qual=self.max_qual, rplane_seq=rplane)

    def var_index_of(self, idx: Sequence[int]) -> np.ndarray:
        """Batch-global variant indices of the variants of superclusters `idx` (order of take())."""
        idx = np.asarray(idx, np.int64)
        hb = (idx[:, None] * 4 + np.arange(4)[None, :]).reshape(-1)
        cnt = self.var_off[hb + 1] - self.var_off[hb]
        return _range_index(self.var_off[hb], cnt)

    @staticmethod
    def concat(parts: Sequence["Batch"]) -> "Batch":
        ref_off = [np.zeros(1, np.int64)]
        var_off = [np.zeros(1, np.int64)]
        alt_off = [np.zeros(1, np.int64)]
        rb = vb = ab = 0
        for p in parts:
            ref_off.append(p.ref_off[1:] + rb)
            var_off.append(p.var_off[1:] + vb)
            alt_off.append(p.alt_off[1:] + ab)
            rb += p.ref_bytes; vb += p.n_var; ab += p.alt_bytes
        any_rp = any(p.rplane_seq is not None for p in parts)
        return Batch(
            ref_off=np.concatenate(ref_off),
            ref_seq=np.concatenate([p.ref_seq[: p.ref_bytes] for p in parts]),
            var_off=np.concatenate(var_off),
            var_pos=np.concatenate([p.var_pos[: p.n_var] for p in parts]),
            var_rlen=np.concatenate([p.var_rlen[: p.n_var] for p in parts]),
            var_type=np.concatenate([p.var_type[: p.n_var] for p in parts]),
            alt_off=np.concatenate(alt_off),
            alt_seq=np.concatenate([p.alt_seq[: p.alt_bytes] for p in parts]),
            var_qual=np.concatenate([p.var_qual[: p.n_var] for p in parts]),
            max_qual=parts[0].max_qual,
            rplane_seq=(np.concatenate([(p.rplane_seq if p.rplane_seq is not None else p.ref_seq)[: p.ref_bytes]
                                        for p in parts]) if any_rp else None))


def _range_index(starts: np.ndarray, counts: np.ndarray) -> np.ndarray:
    """Concatenation of arange(starts[i], starts[i]+counts[i])."""
    counts = np.asarray(counts, np.int64)
    total = int(counts.sum())
    if total == 0:
        return np.zeros(0, np.int64)
    ends = np.cumsum(counts)
    base = np.repeat(np.asarray(starts, np.int64) - (ends - counts), counts)
    return base + np.arange(total, dtype=np.int64)


def _gather_ranges(src: np.ndarray, starts: np.ndarray, counts: np.ndarray) -> np.ndarray:
    return src[_range_index(starts, counts)]


Variant = Tuple[int, int, int, bytes, float]   # (pos, type, rlen, alt, qual)


class BatchBuilder:
    """Append superclusters one at a time (tests, fixtures, small synthetic sets)."""

    def __init__(self, max_qual: float = 60.0):
        self.max_qual = max_qual
        self._ref: List[bytes] = []
        self._rplane: List[bytes] = []
        self._any_rplane = False
        self._var_cnt: List[int] = []
        self._pos: List[int] = []
        self._rlen: List[int] = []
        self._type: List[int] = []
        self._alt: List[bytes] = []
        self._qual: List[float] = []

    def add(self, ref_window: bytes, haps: Sequence[Sequence[Variant]], rplane: Optional[bytes] = None):
        assert len(haps) == 4
        self._ref.append(ref_window)
        self._rplane.append(rplane if rplane is not None else ref_window)
        self._any_rplane |= rplane is not None
        for h in haps:
            self._var_cnt.append(len(h))
            for pos, ty, rlen, alt, qual in h:
                self._pos.append(pos); self._type.append(ty); self._rlen.append(rlen)
                self._alt.append(alt); self._qual.append(qual)

    def build(self) -> Batch:
        ref_off = np.concatenate([[0], np.cumsum([len(r) for r in self._ref])]).astype(np.int64)
        alt_off = np.concatenate([[0], np.cumsum([len(a) for a in self._alt])]).astype(np.int64)
        var_off = np.concatenate([[0], np.cumsum(self._var_cnt)]).astype(np.int64)
        return Batch(
            ref_off=ref_off,
            ref_seq=np.frombuffer(b"".join(self._ref), np.uint8).copy(),
            var_off=var_off,
            var_pos=np.array(self._pos, np.int32), var_rlen=np.array(self._rlen, np.int32),
            var_type=np.array(self._type, np.uint8), alt_off=alt_off,
            alt_seq=np.frombuffer(b"".join(self._alt), np.uint8).copy(),
            var_qual=np.array(self._qual, np.float32), max_qual=self.max_qual,
            rplane_seq=(np.frombuffer(b"".join(self._rplane), np.uint8).copy() if self._any_rplane else None))


@dataclass
class Out:
    """Host-side `vd_batch_out` buffers."""
    n_sc: int
    n_var: int
    aln_score: np.ndarray = field(init=False)
    aln_end_plane: np.ndarray = field(init=False)
    aln_beg_plane: np.ndarray = field(init=False)
    status: np.ndarray = field(init=False)
    assigned: np.ndarray = field(init=False)
    sync_group: np.ndarray = field(init=False)
    ref_ed: np.ndarray = field(init=False)
    query_ed: np.ndarray = field(init=False)
    callq: np.ndarray = field(init=False)

    def __post_init__(self):
        a, v = 4 * self.n_sc, max(2 * self.n_var, 1)
        self.aln_score = np.full(max(a, 1), -1, np.int32)
        self.aln_end_plane = np.full(max(a, 1), 255, np.uint8)
        self.aln_beg_plane = np.full(max(a, 1), 255, np.uint8)
        self.status = np.zeros(max(a, 1), np.uint32)
        self.assigned = np.zeros(v, np.uint8)
        self.sync_group = np.zeros(v, np.int32)
        self.ref_ed = np.zeros(v, np.int32)
        self.query_ed = np.zeros(v, np.int32)
        self.callq = np.zeros(v, np.float32)

    FIELDS = ("aln_score", "aln_end_plane", "aln_beg_plane", "status", "assigned",
              "sync_group", "ref_ed", "query_ed", "callq")

    def as_c(self) -> vd_batch_out:
        s = vd_batch_out()
        for f in self.FIELDS:
            setattr(s, f, _ptr(getattr(self, f)))
        s._keep = self
        return s

    def trimmed(self) -> dict:
        a, v = 4 * self.n_sc, 2 * self.n_var
        d = {}
        for f in self.FIELDS[:4]:
            d[f] = getattr(self, f)[:a]
        for f in self.FIELDS[4:]:
            d[f] = getattr(self, f)[:v]
        return d


class CompactIn:
    """Host-side `vd_compact_in`: the batch with lengths instead of 64-bit offsets and 16-bit positions (filled by
    vd_compact_pack, capi.compact).  ref_seq / rplane_seq / var_type / alt_seq / var_qual are the batch's own arrays."""
    OWN = (("ref_len", np.uint16, "sc"), ("hap_nvar", np.uint8, "sc4"), ("var_pos", np.uint16, "var"), ("var_rlen", np.uint16, "var"),
           ("alt_len", np.uint16, "var"), ("blk_var", np.int64, "blk"), ("blk_ref", np.int64, "blk"), ("blk_alt", np.int64, "blk"))
    SHARED = ("ref_seq", "rplane_seq", "var_type", "alt_seq", "var_qual")

    def __init__(self, batch: "Batch"):
        self.batch = batch
        n_blk = (batch.n_sc + COMPACT_BLOCK - 1) // COMPACT_BLOCK
        size = {"sc": batch.n_sc, "sc4": 4 * batch.n_sc, "var": batch.n_var, "blk": n_blk + 1}
        for name, dt, per in self.OWN:
            setattr(self, name, np.zeros(max(size[per], 1), dt))
        for name in self.SHARED:
            setattr(self, name, getattr(batch, name))
        self.c = vd_compact_in()

    def refresh_pointers(self):
        """After replacing arrays (e.g. by page-locked copies): point the C struct at them."""
        for name, _, _ in self.OWN:
            setattr(self.c, name, _ptr(getattr(self, name)))
        for name in self.SHARED:
            setattr(self.c, name, _ptr(getattr(self, name)))

    def nbytes(self) -> int:
        b = self.batch
        return (sum(getattr(self, n).nbytes for n, _, _ in self.OWN[:5]) + b.ref_bytes * (2 if b.rplane_seq is not None else 1)
                + b.n_var + b.alt_bytes + 4 * b.n_var)


@dataclass
class PackedOut:
    """Host-side `vd_packed_out` buffers (16-bit records)."""
    n_sc: int
    n_var: int

    FIELDS = ("aln_score", "aln_planes", "status", "sync_group", "ref_ed", "query_ed", "callq")

    def __post_init__(self):
        a, v = max(4 * self.n_sc, 1), max(2 * self.n_var, 1)
        self.aln_score = np.full(a, 0xFFFF, np.uint16)
        self.aln_planes = np.zeros(a, np.uint8)
        self.status = np.zeros(a, np.uint16)
        self.sync_group = np.zeros(v, np.uint16)
        self.ref_ed = np.zeros(v, np.uint16)
        self.query_ed = np.zeros(v, np.uint16)
        self.callq = np.zeros(v, np.float32)

    def as_c(self) -> vd_packed_out:
        s = vd_packed_out()
        for f in self.FIELDS:
            setattr(s, f, _ptr(getattr(self, f)))
        s._keep = self
        return s

    def widened(self) -> dict:
        """The same results in the field layout of Out.trimmed()."""
        a, v = 4 * self.n_sc, 2 * self.n_var
        sc = self.aln_score[:a].astype(np.int32)
        sc[self.aln_score[:a] == 0xFFFF] = -1
        return {"aln_score": sc, "aln_end_plane": self.aln_planes[:a] & 1, "aln_beg_plane": (self.aln_planes[:a] >> 1) & 1,
                "status": self.status[:a].astype(np.uint32), "assigned": (self.sync_group[:v] >> 14).astype(np.uint8),
                "sync_group": (self.sync_group[:v] & 0x3FFF).astype(np.int32), "ref_ed": self.ref_ed[:v].astype(np.int32),
                "query_ed": self.query_ed[:v].astype(np.int32), "callq": self.callq[:v]}

    def nbytes(self) -> int:
        return sum(getattr(self, f).nbytes for f in self.FIELDS)


@dataclass
class Final:
    """Per-variant / per-supercluster results in the reference's own terms
    (src/variant.h:49-60, src/cluster.h:36-42)."""
    n_sc: int
    n_var: int

    def __post_init__(self):
        v, s = max(2 * self.n_var, 1), max(self.n_sc, 1)
        self.errtypes = np.full(v, ERRTYPE_UN, np.uint8)
        self.credit = np.zeros(v, np.float32)
        self.callq = np.zeros(v, np.float32)
        self.sync_group = np.zeros(v, np.int32)
        self.ref_ed = np.zeros(v, np.int32)
        self.query_ed = np.zeros(v, np.int32)
        self.sc_phase = np.full(s, PHASE_NONE, np.int32)
        self.orig_dist = np.full(s, -1, np.int32)
        self.swap_dist = np.full(s, -1, np.int32)

    FIELDS = ("errtypes", "credit", "callq", "sync_group", "ref_ed", "query_ed",
              "sc_phase", "orig_dist", "swap_dist")

    def as_c(self) -> vd_final:
        s = vd_final()
        for f in self.FIELDS:
            setattr(s, f, _ptr(getattr(self, f)))
        s._keep = self
        return s

    def trimmed(self) -> dict:
        v = 2 * self.n_var
        d = {f: getattr(self, f)[:v] for f in self.FIELDS[:6]}
        d.update({f: getattr(self, f)[: self.n_sc] for f in self.FIELDS[6:]})
        return d
