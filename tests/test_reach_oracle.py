"""SURVEY 8f-1 groundwork: the C restatement of wf_swg_max_reach (oracle/vd_reach.c) against the
reference's own object code (src/dist.cpp:2150-2333 behind oracle/ref_harness.cpp) on fresh random
cases, and against known answers recorded from the reference (tests/golden/reach_kat.npz) where the
reference objects are not available.  CPU only; no CUDA path exists for this row yet."""
import os

import numpy as np
import pytest

import json

from conftest import ROOT
from oracle import checkers, vd_cluster

KAT = os.path.join(ROOT, "tests", "golden", "reach_kat.npz")


def random_case(rng):
    """A cluster-like pair: truth = reference window, query = the window with a few variants applied
    (so that main_diag = net length change, as src/cluster.cpp:1059-1064 computes it), optionally
    reversed as the leftward search does (src/cluster.cpp:1077-1078)."""
    alphabet = b"ACGT"[: int(rng.integers(2, 5))]
    tlen = int(rng.integers(2, 90))
    truth = bytes(rng.choice(list(alphabet), tlen).tolist())
    q = bytearray()
    main_diag = 0
    pos = 0
    last_var_end = 0
    while pos < tlen:
        r = rng.random()
        if r < 0.06:                                   # substitution
            q.append(alphabet[(alphabet.index(truth[pos]) + 1) % len(alphabet)]); pos += 1; last_var_end = pos
        elif r < 0.10:                                 # insertion
            n = int(rng.integers(1, 6)); q.extend(rng.choice(list(alphabet), n).tolist()); main_diag -= n; last_var_end = pos
        elif r < 0.14:                                 # deletion
            n = int(rng.integers(1, 6)); pos += n; main_diag += min(n, tlen - (pos - n)); last_var_end = min(pos, tlen)
        else:
            q.append(truth[pos]); pos += 1
    if not q:
        q.append(alphabet[0])
    query = bytes(q)
    reverse = bool(rng.integers(0, 2))
    if reverse:
        query, truth = query[::-1], truth[::-1]
    sub, open_, extend = int(rng.integers(0, 6)), int(rng.integers(0, 7)), int(rng.integers(1, 4))
    max_score = int(rng.integers(0, 40))
    main_diag_start = int(rng.integers(0, tlen + 3)) if rng.random() < 0.5 else last_var_end
    return query, truth, main_diag, main_diag_start, max_score, sub, open_, extend, reverse


@pytest.mark.skipif(not checkers.reference_available(False), reason="reference objects not built (no /root/reference)")
def test_restatement_matches_reference_object_code():
    rng = np.random.default_rng()          # fresh cases every run
    for _ in range(4000):
        case = random_case(rng)
        assert checkers.reach_oracle(*case) == checkers.reach_reference(*case), case


def test_restatement_matches_recorded_reference_answers():
    z = np.load(KAT, allow_pickle=False)
    n = len(z["answer"])
    assert n >= 1000
    for i in range(n):
        q = z["query"][z["q_off"][i]: z["q_off"][i + 1]].tobytes()
        t = z["truth"][z["t_off"][i]: z["t_off"][i + 1]].tobytes()
        p = [int(x) for x in z["params"][i]]
        assert checkers.reach_oracle(q, t, p[0], p[1], p[2], p[3], p[4], p[5], bool(p[6])) == int(z["answer"][i]), i


def random_pair(rng):
    alphabet = b"ACGT"[: int(rng.integers(2, 5))]
    q = bytes(rng.choice(list(alphabet), int(rng.integers(1, 60))).tolist())
    if rng.random() < 0.6:                      # truth = query with a few edits (the clustering use case)
        t = bytearray(q)
        for _ in range(int(rng.integers(0, 5))):
            p = int(rng.integers(0, len(t) + 1))
            r = rng.random()
            if r < 0.4 and len(t) > 1:
                del t[p: p + int(rng.integers(1, 5))]
            elif r < 0.8:
                t[p:p] = bytes(rng.choice(list(alphabet), int(rng.integers(1, 5))).tolist())
            elif p < len(t):
                t[p] = alphabet[(alphabet.index(t[p]) + 1) % len(alphabet)]
        t = bytes(t) if len(t) else alphabet[:1]
    else:
        t = bytes(rng.choice(list(alphabet), int(rng.integers(1, 60))).tolist())
    return q, t, int(rng.integers(1, 7)), int(rng.integers(0, 8)), int(rng.integers(1, 4))


@pytest.mark.skipif(not checkers.reference_available(False), reason="reference objects not built (no /root/reference)")
def test_affine_score_matches_reference_object_code():
    """wf_swg_align's score (src/dist.cpp:1510-1652): the budget of the reach searches."""
    rng = np.random.default_rng()
    for _ in range(4000):
        case = random_pair(rng)
        assert checkers.swg_score_oracle(*case) == checkers.swg_score_reference(*case), case


def test_affine_score_known_answers():
    """Recorded from the reference (first bases are always paired: no leading gap)."""
    for q, t, x, o, e, want in [(b"CA", b"A", 1, 1, 1, 3), (b"CA", b"A", 3, 1, 1, 5), (b"A", b"CA", 1, 1, 1, 3),
                               (b"AC", b"A", 3, 1, 1, 2), (b"A", b"AC", 1, 1, 1, 2), (b"GCA", b"A", 3, 1, 1, 6),
                               (b"CCCA", b"A", 1, 1, 1, 5), (b"ACCC", b"A", 3, 1, 1, 4)]:
        assert checkers.swg_score_oracle(q, t, x, o, e) == want
    z = np.load(KAT, allow_pickle=False)
    if "swg_answer" in z.files:
        for i in range(len(z["swg_answer"])):
            q = z["query"][z["q_off"][i]: z["q_off"][i + 1]].tobytes()
            t = z["truth"][z["t_off"][i]: z["t_off"][i + 1]].tobytes()
            p = [int(v) for v in z["swg_params"][i]]
            assert checkers.swg_score_oracle(q, t, p[0], p[1], p[2]) == int(z["swg_answer"][i]), i


CLUSTER_KAT = os.path.join(ROOT, "tests", "golden", "cluster_kat.json")


def random_cluster_case(rng):
    """One haplotype of one contig: random reference with short tandem repeats (so that reaches wander),
    sorted non-overlapping SUB / INS / DEL variants away from the contig ends, random penalties."""
    A = b"ACGT"[: int(rng.integers(2, 5))]
    L = int(rng.integers(60, 400))
    unit = bytes(rng.choice(list(A), int(rng.integers(1, 4))).tolist())
    fasta = bytearray(rng.choice(list(A), L).tolist())
    for _ in range(int(rng.integers(0, 4))):
        p = int(rng.integers(0, L - 20)); k = int(rng.integers(5, 20))
        fasta[p: p + k] = (unit * k)[:k]
    fasta = bytes(fasta)
    var, pos = [], int(rng.integers(3, 15))
    while pos < L - 12:
        r = rng.random()
        if r < 0.4:
            var.append((pos, 1, vd_cluster.TYPE_SUB, bytes([A[(A.index(fasta[pos]) + 1) % len(A)]]))); pos += 1
        elif r < 0.7:
            var.append((pos, 0, vd_cluster.TYPE_INS, bytes(rng.choice(list(A), int(rng.integers(1, 6))).tolist()))); pos += 1
        else:
            n = int(rng.integers(1, 6)); var.append((pos, n, vd_cluster.TYPE_DEL, b"")); pos += n
        pos += int(rng.integers(1, 40))
    return fasta, var, int(rng.integers(1, 6)), int(rng.integers(0, 7)), int(rng.integers(1, 4))


@pytest.mark.skipif(not checkers.reference_available(False), reason="reference objects not built (no /root/reference)")
def test_cluster_growth_matches_reference_object_code():
    """wf_swg_cluster (src/cluster.cpp:954-1263): cluster boundaries and reaches."""
    rng = np.random.default_rng()
    merged = 0
    for _ in range(600):
        fasta, var, x, o, e = random_cluster_case(rng)
        if not var:
            continue
        got = vd_cluster.wf_swg_cluster(fasta, var, x, o, e)
        want = checkers.cluster_reference(fasta, var, x, o, e)
        assert [list(v) for v in got] == want, (fasta, var, x, o, e)
        merged += len(got[0]) - 1 < len(var)
    assert merged > 100


def test_cluster_growth_known_answers():
    kat = json.load(open(CLUSTER_KAT))
    assert len(kat) >= 150
    for c in kat:
        var = [(v[0], v[1], v[2], v[3].encode()) for v in c["var"]]
        got = vd_cluster.wf_swg_cluster(c["fasta"].encode(), var, *c["penalties"])
        assert [list(v) for v in got] == c["answer"]


@pytest.mark.skipif(not checkers.reference_available(False), reason="reference objects not built (no /root/reference)")
def test_affine_alignment_cigar_matches_reference_object_code():
    """wf_swg_align + wf_swg_backtrack (src/dist.cpp:1510-1652, :2625-2757), the `--distance` kernel pair
    (SURVEY 8f-2): score and CIGAR.  The reference is only called where the restatement walks back cleanly
    (its own failure mode is ERROR() + exit, which would take the test process with it)."""
    rng = np.random.default_rng()
    n = 0
    for _ in range(3000):
        case = random_pair(rng)
        so, co = checkers.swg_cigar_oracle(*case)
        assert so == checkers.swg_score_oracle(*case)
        if so < 0:
            continue
        sr, cr = checkers.swg_cigar_reference(*case)
        assert so == sr and (co == cr).all(), case
        n += 1
    assert n > 2900


def test_affine_alignment_cigar_known_answers():
    z = np.load(KAT, allow_pickle=False)
    m = len(z["cig_score"])
    assert m >= 500
    for i in range(m):
        q = z["query"][z["q_off"][i]: z["q_off"][i + 1]].tobytes()
        t = z["truth"][z["t_off"][i]: z["t_off"][i + 1]].tobytes()
        p = [int(v) for v in z["swg_params"][i]]
        s, cig = checkers.swg_cigar_oracle(q, t, p[0], p[1], p[2])
        assert s == int(z["cig_score"][i]) == int(z["swg_answer"][i])
        assert (cig == z["cigar"][z["cig_off"][i]: z["cig_off"][i + 1]]).all(), i
