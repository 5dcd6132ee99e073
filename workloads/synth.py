"""Seeded synthetic supercluster batches (there is no HG002 data and no network here).

* `adversarial`  — short windows over 2-4 letter alphabets with independent random SUB/INS/DEL
                   on all four haplotypes: dense ties, repeats, partial credit, swap edges.
* `wgs_like`     — the measured demo (HG002 chr1:1-5Mb vs GIAB v4.2.1) supercluster-size
                   mixture of SURVEY.md 8d: query = truth with a few FN / FP / representation
                   differences; optional SV tail (one INS or DEL, length log-uniform 50..10k).
* `bootstrap`    — resample superclusters of an existing batch (e.g. the real demo batch in
                   tests/golden/) with replacement up to WGS scale.
* `read_vdarr`   — reader for the array container written by the fixture tool.
"""
from __future__ import annotations

import struct
from typing import List, Optional, Sequence, Tuple

import numpy as np

from vcfdist_b200.batch import Batch, BatchBuilder, TYPE_DEL, TYPE_INS, TYPE_SUB, Variant

_DT = {"q": np.int64, "i": np.int32, "B": np.uint8, "f": np.float32}


def read_vdarr(path: str) -> dict:
    """Arrays written by vdhost::put_arr (vcfdist_b200/host/pr_dropin.cpp)."""
    out = {}
    with open(path, "rb") as f:
        assert f.read(8) == b"VDARR001", "not a VDARR001 file"
        while True:
            hdr = f.read(33)
            if len(hdr) < 33:
                break
            name = hdr[:24].split(b"\0")[0].decode()
            dt = _DT[chr(hdr[24])]
            (n,) = struct.unpack("<q", hdr[25:33])
            out[name] = np.frombuffer(f.read(n * np.dtype(dt).itemsize), dt).copy()
    return out


def batch_from_vdarr(path: str) -> Batch:
    d = read_vdarr(path)
    return Batch(ref_off=d["ref_off"], ref_seq=d["ref_seq"], var_off=d["var_off"], var_pos=d["var_pos"],
                 var_rlen=d["var_rlen"], var_type=d["var_type"], alt_off=d["alt_off"], alt_seq=d["alt_seq"],
                 var_qual=d["var_qual"], max_qual=float(d["max_qual"][0]),
                 rplane_seq=d.get("rplane_seq"))


def _rand_seq(rng: np.random.Generator, n: int, alphabet: bytes) -> bytes:
    a = np.frombuffer(alphabet, np.uint8)
    return a[rng.integers(0, len(a), n)].tobytes()


def random_hap(rng: np.random.Generator, ref: bytes, p_var: float, max_indel: int, alphabet: bytes,
               qual_lo: float = 3.0, qual_hi: float = 50.0, first_pos: int = 1) -> List[Variant]:
    """Sorted, non-overlapping variants as the reference's parser admits them
    (src/variant.cpp:852-861): next.pos >= prev.pos + prev.rlen, never two INS at one position;
    INS followed by SUB/DEL at the same position is allowed (a split CPX, :866-871)."""
    W = len(ref)
    out: List[Variant] = []
    pos = first_pos
    last_ins_pos = -1
    while pos <= W - 2:
        if rng.random() < p_var:
            ty = int(rng.integers(1, 4))
            q = float(np.float32(rng.uniform(qual_lo, qual_hi)))
            if ty == TYPE_SUB:
                choices = [c for c in alphabet if c != ref[pos]]
                if not choices:
                    pos += 1
                    continue
                out.append((pos, TYPE_SUB, 1, bytes([choices[int(rng.integers(0, len(choices)))]]), q))
                pos += 1
            elif ty == TYPE_INS:
                if last_ins_pos == pos:
                    pos += 1
                    continue
                n = int(rng.integers(1, max_indel + 1))
                out.append((pos, TYPE_INS, 0, _rand_seq(rng, n, alphabet), q))
                last_ins_pos = pos
                if rng.random() < 0.5:
                    pos += 1          # else: allow SUB/DEL at this very position next
            else:
                n = int(rng.integers(1, max_indel + 1))
                if pos + n > W - 1:
                    pos += 1
                    continue
                out.append((pos, TYPE_DEL, n, b"", q))
                pos += n
        else:
            pos += 1
    return out


def adversarial(seed: int, n_sc: int, min_len: int = 6, max_len: int = 40, p_var: float = 0.12,
                max_indel: int = 4, max_qual: float = 60.0) -> Batch:
    rng = np.random.default_rng(seed)
    bb = BatchBuilder(max_qual)
    for _ in range(n_sc):
        alphabet = [b"AC", b"ACG", b"ACGT", b"AT"][int(rng.integers(0, 4))]
        W = int(rng.integers(min_len, max_len + 1))
        mode = rng.random()
        if mode < 0.3:       # tandem repeat window: many equivalent representations
            unit = _rand_seq(rng, int(rng.integers(1, 4)), alphabet)
            ref = (unit * (W // len(unit) + 1))[:W]
        else:
            ref = _rand_seq(rng, W, alphabet)
        if rng.random() < 0.5:       # independent haplotypes
            haps = [random_hap(rng, ref, p_var, max_indel, alphabet) for _ in range(4)]
        else:                        # query derived from truth with drop-outs and extras
            t1 = random_hap(rng, ref, p_var, max_indel, alphabet)
            t2 = random_hap(rng, ref, p_var, max_indel, alphabet) if rng.random() < 0.7 else list(t1)
            def derive(t):
                if rng.random() < 0.3:
                    return random_hap(rng, ref, p_var, max_indel, alphabet)
                q = float(np.float32(rng.uniform(3, 50)))
                return [(v[0], v[1], v[2], v[3], q) for v in t if rng.random() > 0.15]
            q1, q2 = derive(t1), derive(t2)
            if rng.random() < 0.3:
                q1, q2 = q2, q1      # swapped phasing
            haps = [q1, q2, t1, t2]
        bb.add(ref, haps)
    return bb.build()


# demo histogram of SIZE = end - beg: class upper bounds and probabilities
_SIZE_CLASSES = [(2, 3, 0.929), (4, 7, 0.032), (8, 15, 0.030), (16, 31, 0.0076), (32, 63, 0.0014)]


def wgs_like(seed: int, n_sc: int, sv_frac: float = 0.0, sv_min: int = 50, sv_max: int = 10000,
             max_indel: int = 50, fn_rate: float = 0.006, fp_rate: float = 0.0015,
             max_qual: float = 60.0) -> Batch:
    """Small-variant superclusters with the demo's size mixture; `sv_frac` of them carry one
    INS or DEL of log-uniform length (half matched between truth and query, half 1 % divergent)."""
    rng = np.random.default_rng(seed)
    bb = BatchBuilder(max_qual)
    probs = np.array([c[2] for c in _SIZE_CLASSES])
    probs = probs / probs.sum()
    alphabet = b"ACGT"
    for _ in range(n_sc):
        q = lambda: float(np.float32(rng.uniform(3, 50)))
        if rng.random() < sv_frac:
            L = int(round(float(np.exp(rng.uniform(np.log(sv_min), np.log(sv_max))))))
            flank = int(rng.integers(2, 30))
            if rng.random() < 0.5:                       # insertion
                ref = _rand_seq(rng, 2 * flank + 1, alphabet)
                ins = bytearray(_rand_seq(rng, L, alphabet))
                tv = (flank, TYPE_INS, 0, bytes(ins), q())
                if rng.random() < 0.5:                   # 1 % divergent copy in the query
                    k = max(1, L // 100)
                    for p in rng.integers(0, L, k):
                        ins[p] = alphabet[(alphabet.index(ins[p]) + 1) % 4]
                qv = (flank, TYPE_INS, 0, bytes(ins), q())
            else:                                        # deletion
                ref = _rand_seq(rng, L + 2 * flank + 1, alphabet)
                tv = (flank, TYPE_DEL, L, b"", q())
                if rng.random() < 0.5 and L > 100:       # query deletes a slightly different span
                    d = max(1, L // 100)
                    qv = (flank + d, TYPE_DEL, L - d, b"", q())
                else:
                    qv = (flank, TYPE_DEL, L, b"", q())
            hom = rng.random() < 0.35
            t1, t2 = [tv], ([tv] if hom else [])
            q1, q2 = [qv], ([qv] if hom else [])
            if rng.random() < 0.3:
                q1, q2 = q2, q1
            bb.add(ref, [q1, q2, t1, t2])
            continue
        lo, hi, _p = _SIZE_CLASSES[int(rng.choice(len(_SIZE_CLASSES), p=probs))]
        W = int(rng.integers(lo, hi + 1)) + 1            # Lr = SIZE + 1
        ref = _rand_seq(rng, W, alphabet)
        if W <= 4:                                       # the 93 % class: one SNP or 1-bp indel
            pos = 1
            r = rng.random()
            if r < 0.86:
                alt = bytes([alphabet[(alphabet.index(ref[pos]) + 1 + int(rng.integers(0, 3))) % 4]])
                tv = [(pos, TYPE_SUB, 1, alt, q())]
            elif r < 0.93:
                tv = [(pos, TYPE_INS, 0, _rand_seq(rng, int(rng.integers(1, 4)), alphabet), q())]
            else:
                tv = [(pos, TYPE_DEL, min(W - 2, 1 + int(rng.integers(0, 2))), b"", q())]
        else:
            tv = random_hap(rng, ref, 3.0 / W, min(max_indel, max(1, W // 2)), alphabet)
        hom = rng.random() < 0.4
        t1, t2 = list(tv), (list(tv) if hom else [])
        def derive(t):
            out = [(v[0], v[1], v[2], v[3], q()) for v in t if rng.random() > fn_rate]
            if rng.random() < fp_rate * max(1, len(t)):
                extra = random_hap(rng, ref, 1.5 / W, 2, alphabet)
                taken = {v[0] for v in out}
                out = sorted(out + [e for e in extra if e[0] not in taken and e[1] == TYPE_SUB],
                             key=lambda v: v[0])
            return out
        q1, q2 = derive(t1), derive(t2)
        if rng.random() < 0.5:                           # unphased truth/query orientation
            t1, t2 = t2, t1
        if rng.random() < 0.5:
            q1, q2 = q2, q1
        bb.add(ref, [q1, q2, t1, t2])
    return bb.build()


def sv_pairs(seed: int, n_sc: int, length: int, divergence: float = 0.01, flank: int = 20,
             max_qual: float = 60.0) -> Batch:
    """Kernel-sweep batch: every supercluster carries one homozygous insertion of `length`
    bases in truth and a `divergence`-mutated copy in the query (Lq ~ Lt ~ length)."""
    rng = np.random.default_rng(seed)
    bb = BatchBuilder(max_qual)
    alphabet = b"ACGT"
    for _ in range(n_sc):
        ref = _rand_seq(rng, 2 * flank + 1, alphabet)
        ins = bytearray(_rand_seq(rng, length, alphabet))
        tv = (flank, TYPE_INS, 0, bytes(ins), 30.0)
        k = int(round(length * divergence))
        for p in rng.integers(0, length, k):
            ins[p] = alphabet[(alphabet.index(ins[p]) + 1) % 4]
        qv = (flank, TYPE_INS, 0, bytes(ins), 25.0)
        bb.add(ref, [[qv], [qv], [tv], [tv]])
    return bb.build()


def bootstrap(seed: int, base: Batch, n_sc: int) -> Batch:
    """`n_sc` superclusters drawn with replacement from `base` (vectorised)."""
    rng = np.random.default_rng(seed)
    return base.take(rng.integers(0, base.n_sc, n_sc))


SV_KINDS = ("ins", "del", "ins_truth_only", "ins_query_only", "del_truth_only", "del_query_only")
SV_ZYG = ("hom", "het", "cross", "mixed")


def sv_case(seed: int, length: int, kind: str = "ins", zyg: str = "hom", divergence: float = 0.0,
            flank: int = 25, flank_snps: bool = True, max_qual: float = 60.0) -> Batch:
    """One supercluster around one structural variant of `length` bases (BASELINE configs[3]: SV to 10 kb).

    kind  ins / del            both sides carry it; the query's copy is `divergence`-mutated (INS) or deletes a
                               span shifted and shortened by length*divergence bases (DEL)
          *_truth_only / *_query_only   only one side carries it (a false negative / false positive SV)
    zyg   hom    both haplotypes of a side carry the side's variants (four identical alignments)
          het    haplotype 1 only
          cross  query on haplotype 2, truth on haplotype 1 (swapped phasing)
          mixed  query homozygous, truth heterozygous
    flank_snps: one SNP in each flank (shared by truth and query, different qualities) so that the path also has
    small-variant sync points around the SV."""
    rng = np.random.default_rng(seed)
    A = b"ACGT"
    d = int(round(length * divergence))
    q = lambda: float(np.float32(rng.uniform(3, 50)))
    is_del = kind.startswith("del")
    W = 2 * flank + 1 + (length if is_del else 0)
    ref = bytearray(_rand_seq(rng, W, A))
    tv: List[Variant] = []
    qv: List[Variant] = []
    if is_del:
        t_sv = (flank, TYPE_DEL, length, b"", q())
        q_sv = (flank + d, TYPE_DEL, max(length - d, 1), b"", q()) if d else (flank, TYPE_DEL, length, b"", q())
    else:
        ins = bytearray(_rand_seq(rng, length, A))
        t_sv = (flank, TYPE_INS, 0, bytes(ins), q())
        for p in rng.integers(0, length, d):
            ins[p] = A[(A.index(ins[p]) + 1) % 4]
        q_sv = (flank, TYPE_INS, 0, bytes(ins), q())
    if not kind.endswith("query_only"):
        tv.append(t_sv)
    if not kind.endswith("truth_only"):
        qv.append(q_sv)
    if flank_snps and flank >= 8:
        for pos in (3, W - 4):
            alt = bytes([A[(A.index(ref[pos]) + 1 + int(rng.integers(0, 3))) % 4]])
            tv.append((pos, TYPE_SUB, 1, alt, q()))
            qv.append((pos, TYPE_SUB, 1, alt, q()))
    tv.sort(key=lambda v: (v[0], v[1] != TYPE_INS))
    qv.sort(key=lambda v: (v[0], v[1] != TYPE_INS))
    if zyg == "hom":
        haps = [qv, list(qv), tv, list(tv)]
    elif zyg == "het":
        haps = [qv, [], tv, []]
    elif zyg == "cross":
        haps = [[], qv, tv, []]
    elif zyg == "mixed":
        haps = [qv, list(qv), tv, []]
    else:
        raise ValueError(zyg)
    bb = BatchBuilder(max_qual)
    bb.add(bytes(ref), haps)
    return bb.build()
