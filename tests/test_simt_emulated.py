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


def test_cluster_wavefront_kernels_and_driver():
    """the reach / score wavefront kernel (one warp per problem) and the batched cluster-growing
    driver over it, against the recorded reference answers."""
    import json
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


def test_distance_alignment_kernel():
    """affine-gap alignment with CIGAR (wf_swg_align + wf_swg_backtrack) against recorded reference answers."""
    from conftest import ROOT
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
