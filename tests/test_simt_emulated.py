"""Kernel LOGIC on the CPU: the sources of vcfdist_b200/csrc compiled with g++ under the SIMT emulator of
tests/simt/simt_emu.h (one fiber per CUDA thread, warp collectives and barriers as rendezvous) and run through
the same C-ABI against the oracle.  A debugging aid for this GPU-less container and a regression net for
indexing / collective / scheduler mistakes; it proves nothing about the sm_100a build, races or performance -
the -m gpu tests do that - and the product path never loads the emulated library."""
import os
import sys

import numpy as np
import pytest

from conftest import OUT_KEYS, load_golden, mismatches

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "simt"))
from emu import EmuEngine  # noqa: E402
from oracle import checkers  # noqa: E402
from vcfdist_b200.batch import Batch  # noqa: E402
from workloads import synth  # noqa: E402


def check(b, **env):
    e = EmuEngine(**env)
    got = e.run(b)
    st = e.stats()
    e.close()
    want = checkers.oracle_run(b)
    assert mismatches(got.trimmed(), want.trimmed(), OUT_KEYS) == {}
    return got, st


def test_short_kernels_on_golden_subsets():
    for name, n in (("demo", 1200), ("adv_11", 250)):
        b, _, _ = load_golden(name)
        check(b.take(np.arange(n)))


def test_banded_warp_kernels_unbounded_rungs():
    """Every supercluster forced onto the long path: small shapes run the banded warp kernels with the window
    covering both planes (swap edges, ties, several swap sources per row)."""
    b, _, _ = load_golden("adv_11")
    _, st = check(b.take(np.arange(300)), VD_FORCE_CLASS=1)
    assert st["n_dense"] == 0
    b, _, _ = load_golden("adv_12")
    check(b.take(np.arange(150)), VD_FORCE_CLASS=1)


def test_banded_warp_kernels_sliding_windows():
    """Structural variants of 400 bases: sliding windows, score-bound ladder, lower-bound pruning, homozygous
    replication on the long path, fall-through to the dense block kernels."""
    cases = [("ins", "hom", 0.01), ("ins", "het", 0.05), ("ins", "cross", 0.15), ("del", "het", 0.01),
             ("ins_truth_only", "het", 0.0), ("del_query_only", "mixed", 0.0)]
    b = Batch.concat([synth.sv_case(300 + i, 400, k, z, d) for i, (k, z, d) in enumerate(cases)]
                     + [synth.sv_case(320, 600, "ins", "het", 0.01), synth.sv_case(321, 560, "ins", "het", 0.5)])
    got, st = check(b)
    assert st["n_long"] == 4 * b.n_sc and 0 < st["n_dense"] < st["n_long"] // 2


def test_dense_block_kernels_alone():
    b = Batch.concat([synth.sv_case(310, 260, "ins", "het", 0.02), synth.sv_case(311, 200, "del", "hom", 0.0)])
    _, st = check(b, VD_BAND=0)
    assert st["n_dense"] > 0


@pytest.mark.parametrize("block_min", [None, "1", "8"])
def test_cluster_wavefront_kernels_and_driver(monkeypatch, block_min):
    """the reach / score wavefront kernel (a warp per problem; with VD_WF_BLOCK_MIN a 256- or 1024-thread block per problem from that width on)
    and the batched cluster-growing driver over it, against the recorded reference answers."""
    import json
    if block_min:
        monkeypatch.setenv("VD_WF_BLOCK_MIN", block_min)
    from vcfdist_b200 import cluster
    from conftest import ROOT
    z = np.load(os.path.join(ROOT, "tests", "golden", "reach_kat.npz"), allow_pickle=False)
    e = EmuEngine()
    idx = [i for i in range(300)]
    groups = {}
    for i in idx:
        groups.setdefault(tuple(int(x) for x in z["params"][i][3:6]), []).append(i)
    for (x, o, ex), ii in groups.items():
        q = [z["query"][z["q_off"][i]: z["q_off"][i + 1]].tobytes() for i in ii]
        t = [z["truth"][z["t_off"][i]: z["t_off"][i + 1]].tobytes() for i in ii]
        p = z["params"][ii]
        got = e.wf_batch(0, q, t, x, o, ex, p[:, 0], p[:, 1], p[:, 2], p[:, 6])
        assert (got == z["answer"][ii]).all()
    kat = json.load(open(os.path.join(ROOT, "tests", "golden", "cluster_kat.json")))
    for c in kat[:40]:
        var = [(v[0], v[1], v[2], v[3].encode()) for v in c["var"]]
        got = cluster.wf_swg_cluster(e, c["fasta"].encode(), var, *c["penalties"])
        assert [list(v) for v in got] == c["answer"]
    e.close()


@pytest.mark.parametrize("block_min", [None, "8"])
def test_distance_alignment_kernel(monkeypatch, block_min):
    """affine-gap alignment with CIGAR (wf_swg_align + wf_swg_backtrack) against recorded reference answers,
    a warp per problem and (VD_WF_BLOCK_MIN=1) a block per problem."""
    from conftest import ROOT
    if block_min:
        monkeypatch.setenv("VD_WF_BLOCK_MIN", block_min)
    z = np.load(os.path.join(ROOT, "tests", "golden", "reach_kat.npz"), allow_pickle=False)
    e = EmuEngine()
    groups = {}
    for i in range(200):
        groups.setdefault(tuple(int(v) for v in z["swg_params"][i]), []).append(i)
    for (x, o, ex), ii in groups.items():
        q = [z["query"][z["q_off"][i]: z["q_off"][i + 1]].tobytes() for i in ii]
        t = [z["truth"][z["t_off"][i]: z["t_off"][i + 1]].tobytes() for i in ii]
        sc, cigs = e.swg_align_batch(q, t, x, o, ex)
        for i, s, cg in zip(ii, sc, cigs):
            assert s == int(z["cig_score"][i]) and (cg == z["cigar"][z["cig_off"][i]: z["cig_off"][i + 1]]).all()
    e.close()


def test_compact_input_and_packed_records():
    """vd_run_compact (offsets rebuilt on the device) and vd_run_packed against vd_run, several pipeline chunks."""
    from vcfdist_b200 import capi
    b0, _, _ = load_golden("demo")
    b = b0.take(np.resize(np.arange(b0.n_sc), 70_000))
    e = EmuEngine(VD_CHUNK_SC=65536)
    wide = e.run(b).trimmed()
    pk = e.run_packed(b).widened()
    ck = e.run_compact(capi.compact(b)).widened()
    e.close()
    assert mismatches(pk, wide, OUT_KEYS) == {}
    assert mismatches(ck, wide, OUT_KEYS) == {}


def test_bit_parallel_levenshtein_against_the_row_dp():
    """lev_myers64 (csrc/vd_scalar.cuh: what wf_ed's `s`, src/dist.cpp:1406-1506, is computed with when the shorter string
    has at most 64 symbols) against the textbook DP: random pairs over ACGT, with N / IUPAC symbols (the per-symbol
    fallback of the match mask), pattern lengths 1..64, one-sided long texts, identical and disjoint strings."""
    import ctypes as C
    from emu import EMU_LIB, build
    build()
    lib = C.CDLL(EMU_LIB)
    lib.vd_emu_lev_myers64.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int]
    lib.vd_emu_lev_myers64.restype = C.c_int

    def dp(a, b):
        prev = list(range(len(b) + 1))
        for i, x in enumerate(a, 1):
            cur = [i]
            for j, y in enumerate(b, 1):
                cur.append(min(prev[j - 1] + (x != y), prev[j] + 1, cur[j - 1] + 1))
            prev = cur
        return prev[-1]

    rng = np.random.default_rng(11)
    alph = [b"ACGT", b"ACGTN", b"ACGTNRYKM", b"AC"]
    cases = [(b"A", b"A"), (b"A", b"C"), (b"ACGT" * 16, b"ACGT" * 16), (b"C" * 300, b"A" * 64), (b"N" * 10, b"ACGTN" * 7)]
    for _ in range(1500):
        al = alph[int(rng.integers(0, len(alph)))]
        n = int(rng.integers(1, 65))
        m = int(rng.integers(1, 400 if rng.random() < 0.2 else 90))
        b = bytes(al[i] for i in rng.integers(0, len(al), n))
        if rng.random() < 0.5:                                   # a mutated copy: small distances
            a = bytearray(b * (m // n + 1))[:m]
            for k in rng.integers(0, max(len(a), 1), int(rng.integers(0, 6))):
                a[int(k)] = al[int(rng.integers(0, len(al)))]
            a = bytes(a)
        else:
            a = bytes(al[i] for i in rng.integers(0, len(al), m))
        cases.append((a, b))
    for a, b in cases:
        assert lib.vd_emu_lev_myers64(a, len(a), b, len(b)) == dp(a, b), (a, b)
