"""SURVEY 8f-1 on the GPU: the affine-gap wavefront kernels of the cluster-growing stage (vd_wf_batch:
wf_swg_max_reach and the score of wf_swg_align, one warp per problem) and the batched cluster-growing driver
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
