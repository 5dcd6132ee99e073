"""SURVEY 8f-1 on the GPU: the affine-gap wavefront kernels of the cluster-growing stage (vd_wf_batch:
wf_swg_max_reach and the score of wf_swg_align, a warp or a block per problem) and the batched cluster-growing driver
over them (vcfdist_b200.cluster.wf_swg_cluster), through the C-ABI, against the known answers recorded from the
reference's own object code (tests/golden/reach_kat.npz, cluster_kat.json), the C restatement and - where the
reference objects travelled to the box - the reference itself on fresh random cases."""
import json
import os
from collections import defaultdict

import numpy as np
import pytest

from conftest import ROOT
from oracle import checkers
from vcfdist_b200 import capi, cluster
import test_reach_oracle as TR

pytestmark = pytest.mark.gpu
KAT = os.path.join(ROOT, "tests", "golden", "reach_kat.npz")
CLUSTER_KAT = os.path.join(ROOT, "tests", "golden", "cluster_kat.json")


@pytest.fixture(scope="module")
def engine():
    e = capi.Engine(0)
    yield e
    e.close()


def by_penalties(cases, key):
    g = defaultdict(list)
    for i, c in enumerate(cases):
        g[key(c)].append(i)
    return g


def test_reach_known_answers(engine):
    """1000+ (query, truth, main_diag, main_diag_start, max_score, penalties, direction) cases recorded from the
    reference's wf_swg_max_reach."""
    z = np.load(KAT, allow_pickle=False)
    n = len(z["answer"])
    cases = []
    for i in range(n):
        q = z["query"][z["q_off"][i]: z["q_off"][i + 1]].tobytes()
        t = z["truth"][z["t_off"][i]: z["t_off"][i + 1]].tobytes()
        cases.append((q, t) + tuple(int(x) for x in z["params"][i]))
    done = 0
    for (x, o, e), idx in by_penalties(cases, lambda c: c[5:8]).items():
        cs = [cases[i] for i in idx]
        got = engine.wf_batch(0, [c[0] for c in cs], [c[1] for c in cs], x, o, e, [c[2] for c in cs], [c[3] for c in cs],
                              [c[4] for c in cs], [c[8] for c in cs])
        assert (got == z["answer"][idx]).all(), (x, o, e)
        done += len(idx)
    assert done == n >= 1000


def test_reach_and_score_random_cases(engine):
    """Fresh random cluster-like pairs: reach against the C restatement (and the reference's object code where it
    is present), alignment score likewise; strings up to a few hundred bases so that diagonals span several lane
    chunks."""
    rng = np.random.default_rng(11)
    cases = [TR.random_case(rng) for _ in range(3000)]
    for (x, o, e), idx in by_penalties(cases, lambda c: c[5:8]).items():
        cs = [cases[i] for i in idx]
        got = engine.wf_batch(0, [c[0] for c in cs], [c[1] for c in cs], x, o, e, [c[2] for c in cs], [c[3] for c in cs],
                              [c[4] for c in cs], [int(c[8]) for c in cs])
        want = [checkers.reach_oracle(*c) for c in cs]
        assert (got == np.array(want)).all(), (x, o, e)
        if checkers.reference_available(False):
            assert (got == np.array([checkers.reach_reference(*c) for c in cs])).all(), (x, o, e)
    pairs = [TR.random_pair(rng) for _ in range(3000)]
    big = b"ACGT" * 90
    pairs += [(big[:300 + k], big[3:290] + b"GG" + big[10:20 + k], 2, 3, 1) for k in range(8)]
    for (x, o, e), idx in by_penalties(pairs, lambda c: c[2:5]).items():
        cs = [pairs[i] for i in idx]
        got = engine.wf_batch(1, [c[0] for c in cs], [c[1] for c in cs], x, o, e)
        assert (got == np.array([checkers.swg_score_oracle(*c) for c in cs])).all(), (x, o, e)


def sv_like_cases(rng, n):
    """Wide problems: a window of one to three thousand bases whose query carries one long insertion or deletion
    (and a few small edits), forwards and reversed - the reached range grows to hundreds of diagonals."""
    out = []
    for _ in range(n):
        tlen = int(rng.integers(900, 2600))
        truth = bytes(rng.choice(list(b"ACGT"), tlen).tolist())
        at = int(rng.integers(50, tlen - 400))
        size = int(rng.integers(150, 700))
        if rng.random() < 0.5:
            q = truth[:at] + bytes(rng.choice(list(b"ACGT"), size).tolist()) + truth[at:]; main_diag = -size
        else:
            q = truth[:at] + truth[at + size:]; main_diag = size
        q = bytearray(q)
        for _ in range(int(rng.integers(0, 4))):
            j = int(rng.integers(0, len(q))); q[j] = b"ACGT"[(b"ACGT".index(q[j]) + 1) % 4]
        q = bytes(q)
        x, o, e = [(3, 2, 1), (2, 3, 1), (1, 1, 1), (4, 6, 2)][int(rng.integers(0, 4))]
        rev = int(rng.random() < 0.5)
        if rev:
            q, truth = q[::-1], truth[::-1]
        budget = int(rng.integers(20, o + e * size + 40))
        out.append((q, truth, main_diag, int(rng.integers(0, len(q))), budget, x, o, e, rev))
    return out


@pytest.mark.parametrize("block_min,cluster_min", [(1, 10 ** 9), (40, 600), (100000, 10 ** 9), (16, 64), (100000, 1)])
def test_block_and_warp_forms_agree(monkeypatch, block_min, cluster_min):
    """vd_wf_batch gives a problem a warp, a 256- or 1024-thread block or a cluster of eight blocks by the width its
    wavefront can reach (VD_WF_BLOCK_MIN, VD_WF_CLUSTER_MIN): all through the block forms, mixed, all through the warp
    form, everything on clusters - same answers as the C restatement, on cluster-sized and on structural-variant-sized
    problems, reach and score."""
    monkeypatch.setenv("VD_WF_BLOCK_MIN", str(block_min))
    monkeypatch.setenv("VD_WF_CLUSTER_MIN", str(cluster_min))
    e = capi.Engine(0)
    rng = np.random.default_rng(23)
    cases = [TR.random_case(rng) for _ in range(600)] + sv_like_cases(rng, 60)
    for (x, o, ex), idx in by_penalties(cases, lambda c: c[5:8]).items():
        cs = [cases[i] for i in idx]
        got = e.wf_batch(0, [c[0] for c in cs], [c[1] for c in cs], x, o, ex, [c[2] for c in cs], [c[3] for c in cs],
                         [c[4] for c in cs], [int(c[8]) for c in cs])
        assert (got == np.array([checkers.reach_oracle(*c) for c in cs])).all(), (x, o, ex)
    pairs = [TR.random_pair(rng) for _ in range(600)] + [(c[0], c[1], c[5], c[6], c[7]) for c in sv_like_cases(rng, 12)]
    for (x, o, ex), idx in by_penalties(pairs, lambda c: c[2:5]).items():
        cs = [pairs[i] for i in idx]
        got = e.wf_batch(1, [c[0] for c in cs], [c[1] for c in cs], x, o, ex)
        assert (got == np.array([checkers.swg_score_oracle(*c) for c in cs])).all(), (x, o, ex)
    # the --distance alignment (scores, CIGARs): this engine's mix of forms against an all-warp engine
    monkeypatch.setenv("VD_WF_BLOCK_MIN", "100000")
    w = capi.Engine(0)
    for (x, o, ex), idx in by_penalties(pairs, lambda c: c[2:5]).items():
        cs = [pairs[i] for i in idx]
        sa, ca = e.swg_align_batch([c[0] for c in cs], [c[1] for c in cs], x, o, ex)
        sb, cb = w.swg_align_batch([c[0] for c in cs], [c[1] for c in cs], x, o, ex)
        assert (np.asarray(sa) == np.asarray(sb)).all() and all((a == b).all() for a, b in zip(ca, cb)), (x, o, ex)
        assert (np.asarray(sa) == np.array([checkers.swg_score_oracle(*c) for c in cs])).all()
    w.close()
    e.close()


def test_cluster_growth_known_answers(engine):
    """wf_swg_cluster (src/cluster.cpp:954-1263) over the GPU kernels: cluster boundaries and reaches recorded from the
    reference's object code."""
    kat = json.load(open(CLUSTER_KAT))
    assert len(kat) >= 150
    for c in kat:
        var = [(v[0], v[1], v[2], v[3].encode()) for v in c["var"]]
        got = cluster.wf_swg_cluster(engine, c["fasta"].encode(), var, *c["penalties"])
        assert [list(v) for v in got] == c["answer"]


def test_cluster_growth_random_cases(engine):
    rng = np.random.default_rng(12)
    merged = 0
    for _ in range(150):
        fasta, var, x, o, e = TR.random_cluster_case(rng)
        if not var:
            continue
        got = cluster.wf_swg_cluster(engine, fasta, var, x, o, e)
        from oracle import vd_cluster
        want = vd_cluster.wf_swg_cluster(fasta, var, x, o, e)
        assert [list(v) for v in got] == [list(v) for v in want], (fasta, var, x, o, e)
        merged += len(got[0]) - 1 < len(var)
    assert merged > 20


def test_distance_alignment_known_answers(engine):
    """wf_swg_align + wf_swg_backtrack on the GPU (vd_swg_align_batch) against the scores and CIGARs
    recorded from the reference's object code, and against the C restatement on fresh random pairs."""
    z = np.load(KAT, allow_pickle=False)
    m = len(z["cig_score"])
    assert m >= 500
    cases = []
    for i in range(m):
        q = z["query"][z["q_off"][i]: z["q_off"][i + 1]].tobytes()
        t = z["truth"][z["t_off"][i]: z["t_off"][i + 1]].tobytes()
        cases.append((q, t) + tuple(int(v) for v in z["swg_params"][i]))
    for (x, o, e), idx in by_penalties(cases, lambda c: c[2:5]).items():
        cs = [cases[i] for i in idx]
        sc, cigs = engine.swg_align_batch([c[0] for c in cs], [c[1] for c in cs], x, o, e)
        for i, s, cg in zip(idx, sc, cigs):
            assert s == int(z["cig_score"][i])
            assert (cg == z["cigar"][z["cig_off"][i]: z["cig_off"][i + 1]]).all(), i
    rng = np.random.default_rng(13)
    pairs = [TR.random_pair(rng) for _ in range(2000)]
    for (x, o, e), idx in by_penalties(pairs, lambda c: c[2:5]).items():
        cs = [pairs[i] for i in idx]
        sc, cigs = engine.swg_align_batch([c[0] for c in cs], [c[1] for c in cs], x, o, e)
        for c, s, cg in zip(cs, sc, cigs):
            so, co = checkers.swg_cigar_oracle(*c)
            assert so == s and (co[: len(cg)] == cg).all(), c


def test_distance_alignment_in_rounds(engine):
    """vd_swg_align_batch with a storage budget far below what the batch needs: the alignments run in rounds over the same
    storage; scores and CIGARs equal to those of an engine with room for everything, scores equal to the C restatement.  An
    alignment that alone exceeds the budget is refused (VD_E_NOMEM), nothing is computed on the CPU instead."""
    small = capi.Engine(0, scratch_bytes=3_000_000)
    rng = np.random.default_rng(17)
    pairs = [TR.random_pair(rng)[:2] for _ in range(400)]
    q, t = [c[0] for c in pairs], [c[1] for c in pairs]
    sa, ca = small.swg_align_batch(q, t, 2, 3, 1)
    sb, cb = engine.swg_align_batch(q, t, 2, 3, 1)
    assert (np.asarray(sa) == np.asarray(sb)).all() and all((a == b).all() for a, b in zip(ca, cb))
    assert (np.asarray(sa) == np.array([checkers.swg_score_oracle(a, b, 2, 3, 1) for a, b in pairs])).all()
    need = sum(5 * (int(s) + 1) * 3 * (len(a) + len(b) - 1) for s, (a, b) in zip(sa, pairs))
    assert need > 2 * 3_000_000                                   # so there were at least three rounds
    big = b"ACGT" * 200
    with pytest.raises(capi.VdError):
        small.swg_align_batch([big], [big[1:400] + b"T" * 300], 2, 3, 1)
    small.close()
