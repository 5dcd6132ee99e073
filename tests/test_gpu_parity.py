"""Parity of the CUDA path (through the C-ABI, vd_run with host buffers) against the oracle
and the committed reference outputs.  Bit-exact: integer scores, planes, status bits, credit
integers, the exact float minima, and after the host float step every reference field."""
import os

import numpy as np
import pytest

from conftest import FINAL_KEYS, OUT_KEYS, load_golden, mismatches, non_tie_var_mask
from vcfdist_b200 import capi
from workloads import synth
from oracle import checkers
from vcfdist_b200.batch import Batch, BatchBuilder, TYPE_DEL, TYPE_INS, TYPE_SUB

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    e = capi.Engine(0)
    yield e
    e.close()


def forced_engine(cls):
    os.environ["VD_FORCE_CLASS"] = str(cls)
    try:
        return capi.Engine(0)
    finally:
        del os.environ["VD_FORCE_CLASS"]


def engine_with(**env):
    os.environ.update({k: str(v) for k, v in env.items()})
    try:
        return capi.Engine(0)
    finally:
        for k in env:
            del os.environ[k]


def check_vs_oracle(engine, b):
    got = engine.run(b)
    want = checkers.oracle_run(b)
    assert mismatches(got.trimmed(), want.trimmed(), OUT_KEYS) == {}
    return got


@pytest.mark.parametrize("name", ["demo", "adv_11", "adv_12", "sv_21"])
def test_golden_through_c_abi(engine, name):
    b, refA, refB = load_golden(name)
    got = check_vs_oracle(engine, b)
    fin = capi.finalize(b, got).trimmed()
    assert mismatches(fin, refB, FINAL_KEYS) == {}
    mask, _ = non_tie_var_mask(b, got.status)
    assert mismatches(fin, refA, FINAL_KEYS, var_mask=mask) == {}


@pytest.mark.parametrize("cls,band", [(1, 1), (1, 0), (2, 1)])
@pytest.mark.parametrize("name", ["demo", "adv_11", "adv_12", "sv_21"])
def test_every_kernel_family_on_golden(name, cls, band):
    """Force all superclusters through the long path (1: banded warp kernels first, then the dense block
    kernels; with VD_BAND=0 the dense block kernels alone) or the scalar-slab kernel (2)."""
    b, _, refB = load_golden(name)
    e = engine_with(VD_FORCE_CLASS=cls, VD_BAND=band)
    got = check_vs_oracle(e, b)
    assert mismatches(capi.finalize(b, got).trimmed(), refB, FINAL_KEYS) == {}
    e.close()


@pytest.mark.parametrize("lo,hi,wsc", [(0, 0, 1), (1, 1, 1), (0, 1, 0), (0, -1, 1), (0, -1, 0)])
@pytest.mark.parametrize("name", ["demo", "adv_11", "adv_12"])
def test_every_short_kernel_on_golden(name, lo, hi, wsc):
    """Restrict the thread-per-alignment kernels to footprint classes lo..hi (hi < lo: none) and switch
    the warp-per-supercluster kernel on/off, so that every kernel also sees the superclusters a cheaper
    one would normally take; whatever is left falls through to the HBM-slab wavefront kernels."""
    b, _, refB = load_golden(name)
    e = engine_with(VD_SMALL_MIN=lo, VD_SMALL_MAX=hi, VD_WSC=wsc)
    got = check_vs_oracle(e, b)
    assert mismatches(capi.finalize(b, got).trimmed(), refB, FINAL_KEYS) == {}
    st = e.stats()
    assert sum(st["n_small"]) * 4 == st["n_short"] and st["n_short"] + st["n_long"] == 4 * b.n_sc
    assert st["n_small"][0] == 0 or lo == 0
    assert st["n_small"][1] == 0 or lo <= 1 <= hi
    assert (st["n_small"][2] > 0) == bool(wsc)
    if name == "demo" and hi >= lo:
        assert sum(st["n_small"][lo:hi + 1]) > 0
    e.close()


def test_homozygous_replication_matches_full_computation():
    """Homozygous superclusters (both query haplotypes identical, both truth haplotypes identical) are
    solved once and replicated; VD_HOM=0 runs all four alignments.  Both must equal the oracle; the
    batch mixes real demo superclusters (43 % homozygous) with synthetic homozygous ones of every size."""
    demo, _, _ = load_golden("demo")
    rng = np.random.default_rng(7)
    bb = BatchBuilder()
    for i in range(400):
        L = int(rng.integers(3, 70))
        ref = bytes(rng.choice(list(b"ACGT"), L).tolist())
        qv = synth.random_hap(rng, ref, 0.15, 4, b"ACGT")
        tv = synth.random_hap(rng, ref, 0.15, 4, b"ACGT") if i % 3 else qv
        bb.add(ref, [qv, qv, tv, tv])
    b = Batch.concat([demo, bb.build()])
    want = checkers.oracle_run(b).trimmed()
    for hom in (1, 0):
        e = engine_with(VD_HOM=hom)
        got = e.run(b).trimmed()
        assert mismatches(got, want, OUT_KEYS) == {}
        st = e.stats()
        assert (st["n_hom"] > 2000) == bool(hom)
        e.close()


def test_rplane_string_differs_from_fasta(engine):
    """REF-plane string given separately (query-hap-1 REF alleles that differ from the FASTA,
    src/dist.cpp:1784-1792): every kernel family must read it, and the no-variant shortcut of the
    warp kernel must not be taken."""
    b = synth.adversarial(41, 1500, max_len=40)
    rng = np.random.default_rng(3)
    rp = b.ref_seq.copy()
    idx = rng.integers(0, len(rp), len(rp) // 15)
    rp[idx] = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, len(idx))]
    b2 = Batch(ref_off=b.ref_off, ref_seq=b.ref_seq, var_off=b.var_off, var_pos=b.var_pos, var_rlen=b.var_rlen,
               var_type=b.var_type, alt_off=b.alt_off, alt_seq=b.alt_seq, var_qual=b.var_qual, max_qual=b.max_qual,
               rplane_seq=rp)
    check_vs_oracle(engine, b2)
    e = engine_with(VD_SMALL_MAX=-1)
    check_vs_oracle(e, b2)
    e.close()


def test_warp_kernel_shapes():
    """Warp-per-supercluster kernel on its own: one to four register slots (<= 32 .. <= 128 rows),
    several swap sources per row (insertions, adjacent deletions), every shared-memory bin."""
    b = Batch.concat([synth.adversarial(31, 600, max_len=14), synth.adversarial(32, 600, max_len=28),
                      synth.adversarial(33, 300, max_len=45), synth.adversarial(35, 200, max_len=60),
                      synth.wgs_like(34, 3000)])
    e = engine_with(VD_SMALL_MAX=-1)
    check_vs_oracle(e, b)
    st = e.stats()
    assert st["n_small"][2] > 1000
    e.close()


@pytest.mark.parametrize("seed", [201, 202, 203])
def test_adversarial_seeds(engine, seed):
    check_vs_oracle(engine, synth.adversarial(seed, 800, max_len=20 + 15 * (seed - 200)))


@pytest.mark.parametrize("seed,max_len", [(301, 8), (302, 12), (303, 18), (304, 26), (305, 40), (306, 58)])
def test_adversarial_sweep_over_kernel_boundaries(engine, seed, max_len):
    """Window lengths straddling every kernel boundary (thread-per-alignment classes 0/1, warp kernel with
    1-4 register slots and every shared-memory bin, HBM-slab path), 1500 superclusters each."""
    check_vs_oracle(engine, synth.adversarial(seed, 1500, max_len=max_len))


def test_wgs_like_mixture(engine):
    b = synth.wgs_like(5, 20000, sv_frac=0.003, sv_max=1500)
    got = check_vs_oracle(engine, b)
    st = engine.stats()
    assert st["cells"] == int(b.cells().sum())
    assert st["n_short"] + st["n_long"] == 4 * b.n_sc


@pytest.mark.parametrize("length", [70, 300, 1200])
def test_sv_lengths(engine, length):
    check_vs_oracle(engine, synth.sv_pairs(3, 3, length, divergence=0.03))


@pytest.mark.parametrize("length,div", [(1400, 0.3), (2300, 0.9), (700, 1.0)])
def test_divergent_sv_pairs(engine, length, div):
    """Long alignments whose score exceeds the banded sweep's bounds (80 ... 1280): retries and
    the dense fallback must give the same bits."""
    check_vs_oracle(engine, synth.sv_pairs(5, 1, length, divergence=div))


def _long_mixed_batch(seed, n_sc):
    """Long alignments (hundreds to thousands of rows) in which the banded forward sweep succeeds but the
    optimal paths carry long runs: a big insertion shared by truth and query (large matrices), plus an
    insertion only the query has (a run of insertion moves longer than the windowed backward sweep's
    margin, so its chain follower has to continue below the window), a deletion only the truth has, and
    scattered substitutions.  Mixed zygosity, so that all four alignments differ."""
    rng = np.random.default_rng(seed)
    A = b"ACGT"
    bb = BatchBuilder()
    for i in range(n_sc):
        W = int(rng.integers(260, 420))
        ref = bytes(rng.choice(list(A), W).tolist())
        big = bytes(rng.choice(list(A), int(rng.integers(300, 1500))).tolist())
        shared = (40, TYPE_INS, 0, big, 30.0)
        q_only = (120, TYPE_INS, 0, bytes(rng.choice(list(A), int(rng.integers(30, 260))).tolist()), 22.0)
        t_only = (180, TYPE_DEL, int(rng.integers(30, 70)), b"", 40.0)
        def subs(lo, hi, n, qual):
            out, used = [], set()
            for p_ in sorted(set(int(x) for x in rng.integers(lo, hi, n))):
                alt = A[(A.index(ref[p_]) + 1 + int(rng.integers(0, 3))) % 4]
                out.append((p_, TYPE_SUB, 1, bytes([alt]), qual))
            return out
        q1 = [shared, q_only] + subs(W - 60, W - 2, 3, 15.0)
        q2 = [shared] if i % 2 else []
        t1 = [shared, t_only] + subs(W - 60, W - 2, 2, 33.0)
        t2 = [shared] if i % 3 else [t_only]
        for lst in (q1, q2, t1, t2):
            lst.sort(key=lambda v: (v[0], v[1] != TYPE_INS))
        bb.add(ref, [q1, q2, t1, t2])
    return bb.build()


def test_long_alignments_with_one_sided_runs(engine):
    """Windowed backward sweep (chain follower, swap targets across planes), dense backward sweep for the
    alignments above the last score bound, walk over long insertion / deletion runs."""
    b = _long_mixed_batch(17, 24)
    check_vs_oracle(engine, b)
    st = engine.stats()
    assert st["n_long"] == 4 * b.n_sc
    for env in (dict(VD_BAND=0), dict(VD_BAND=0, VD_SPARSE_BWD=1)):     # dense block kernels alone; frontier kernel
        e = engine_with(**env)
        check_vs_oracle(e, b)
        assert e.stats()["n_dense"] == 4 * b.n_sc
        e.close()


def test_dense_paths_match_banded_and_sparse(engine):
    """VD_DENSE_FWD / VD_DENSE_BWD select the dense register-blocked sweeps for every alignment,
    VD_SPARSE_BWD the frontier backward kernel instead of the banded one."""
    b = Batch.concat([synth.sv_pairs(7, 2, 800, divergence=0.05), synth.wgs_like(8, 300, sv_frac=0.1, sv_max=900),
                      synth.sv_pairs(9, 2, 2600, divergence=0.02)])
    want = checkers.oracle_run(b)
    for env in ({"VD_DENSE_FWD": "1"}, {"VD_DENSE_BWD": "1"}, {"VD_DENSE_FWD": "1", "VD_DENSE_BWD": "1"},
                {"VD_SPARSE_BWD": "1"}, {"VD_BAND": "0"}, {"VD_BAND": "0", "VD_DENSE_FWD": "1"}, {"VD_BAND": "0", "VD_DENSE_BWD": "1"},
                {"VD_BAND": "0", "VD_DENSE_FWD": "1", "VD_DENSE_BWD": "1"}, {"VD_BAND": "0", "VD_SPARSE_BWD": "1"}):
        os.environ.update(env)
        try:
            e = capi.Engine(0)
        finally:
            for k in env:
                del os.environ[k]
        assert mismatches(e.run(b).trimmed(), want.trimmed(), OUT_KEYS) == {}
        e.close()


def test_edge_cases(engine):
    bb = BatchBuilder()
    # no variants at all (window of one base and of several)
    bb.add(b"A", [[], [], [], []])
    bb.add(b"ACGTAC", [[], [], [], []])
    # truth only / query only
    bb.add(b"ACGTAC", [[], [], [(2, TYPE_SUB, 1, b"T", 30.0)], []])
    bb.add(b"ACGTAC", [[(2, TYPE_SUB, 1, b"T", 30.0)], [], [], []])
    # INS followed by DEL at the same position (a split CPX), deletion up to the last base but one
    bb.add(b"ACGTACGT", [[(2, TYPE_INS, 0, b"TT", 20.0), (2, TYPE_DEL, 3, b"", 20.0)], [],
                         [(2, TYPE_INS, 0, b"TT", 25.0), (2, TYPE_DEL, 3, b"", 25.0)], []])
    # same allele, different representation in a homopolymer
    bb.add(b"CAAAAAT", [[(1, TYPE_DEL, 1, b"", 12.0)], [], [(4, TYPE_DEL, 1, b"", 40.0)], []])
    # swapped phasing
    bb.add(b"ACGTACGTAC", [[(6, TYPE_SUB, 1, b"A", 9.0)], [(2, TYPE_SUB, 1, b"T", 8.0)],
                           [(2, TYPE_SUB, 1, b"T", 30.0)], [(6, TYPE_SUB, 1, b"A", 30.0)]])
    b = bb.build()
    check_vs_oracle(engine, b)


def test_empty_batch(engine):
    b = BatchBuilder().build()
    out = engine.run(b)
    assert out.n_sc == 0


def test_malformed_input_is_reported(engine):
    bb = BatchBuilder()
    bb.add(b"ACGTAC", [[(4, TYPE_SUB, 1, b"T", 30.0), (2, TYPE_SUB, 1, b"G", 30.0)], [], [], []])   # unsorted
    with pytest.raises(capi.VdError) as ei:
        engine.run(bb.build())
    assert ei.value.code == -3


def test_input_the_kernels_cannot_process_is_reported(engine):
    """More than 8 swap sources for one row (9 adjacent deletions): the kernels flag the supercluster
    (VD_ST_ERR_BADINPUT on its alignments) and vd_run returns VD_E_BADINPUT, for a short and for a long
    supercluster; the rest of the batch is computed and a second call on the same handle is clean."""
    rng = np.random.default_rng(4)
    for W in (40, 700):
        ref = bytes(rng.choice(list(b"ACGT"), W).tolist())
        dels = [(5 + k, TYPE_DEL, 1, b"", 30.0) for k in range(9)]
        bb = BatchBuilder()
        bb.add(b"ACGTACGT", [[(2, TYPE_SUB, 1, b"T", 30.0)], [], [(2, TYPE_SUB, 1, b"T", 30.0)], []])
        bb.add(ref, [dels, [], dels, []])
        b = bb.build()
        out = capi.Out(b.n_sc, b.n_var) if hasattr(capi, "Out") else None
        from vcfdist_b200.batch import Out
        out = Out(b.n_sc, b.n_var)
        with pytest.raises(capi.VdError) as ei:
            engine.run(b, out)
        assert ei.value.code == -3
        assert (out.status[4:8] & 0x0800).all() and not (out.status[0:4] & 0xff00).any()
        assert out.aln_score[0] == 0
    check_vs_oracle(engine, synth.adversarial(77, 200, max_len=30))


def test_idempotent_and_order_independent(engine):
    """Size-independent properties: same results on a second run, and per-supercluster results do
    not depend on batch order (superclusters are independent units)."""
    b = synth.wgs_like(9, 5000, sv_frac=0.002, sv_max=800)
    a1 = engine.run(b).trimmed()
    a2 = engine.run(b).trimmed()
    assert mismatches(a1, a2, OUT_KEYS) == {}
    perm = np.random.default_rng(1).permutation(b.n_sc)
    bp = b.take(perm)
    ap = engine.run(bp).trimmed()
    assert (ap["aln_score"].reshape(-1, 4) == a1["aln_score"].reshape(-1, 4)[perm]).all()
    vidx = b.var_index_of(perm)
    for k in ("assigned", "sync_group", "ref_ed", "query_ed", "callq"):
        x = a1[k].reshape(2, -1)[:, vidx]
        assert (ap[k].reshape(2, -1) == x).all(), k


def test_device_slices_match_whole_batch(engine):
    """vd_run_device_slice on four slices of a device-resident batch writes the same records as one
    vd_run_device over the whole batch (and as the oracle)."""
    import torch
    from vcfdist_b200.batch import vd_batch_in, vd_batch_out
    b = synth.wgs_like(21, 6000, sv_frac=0.003, sv_max=500)
    want = checkers.oracle_run(b).trimmed()
    dev = torch.device("cuda", 0)
    names = ("ref_off", "ref_seq", "var_off", "var_pos", "var_rlen", "var_type", "alt_off", "alt_seq", "var_qual")
    d_in = {k: torch.from_numpy(getattr(b, k)).to(dev) for k in names}
    fields = {"aln_score": (torch.int32, 4, "sc"), "aln_end_plane": (torch.uint8, 4, "sc"), "aln_beg_plane": (torch.uint8, 4, "sc"),
              "status": (torch.int32, 4, "sc"), "assigned": (torch.uint8, 2, "var"), "sync_group": (torch.int32, 2, "var"),
              "ref_ed": (torch.int32, 2, "var"), "query_ed": (torch.int32, 2, "var"), "callq": (torch.float32, 2, "var")}
    K = 4
    cut = [b.n_sc * k // K for k in range(K + 1)]
    vcut = [int(b.var_off[4 * c]) for c in cut]
    got = {k: [] for k in fields}
    for k in range(K):
        ns, nv = cut[k + 1] - cut[k], vcut[k + 1] - vcut[k]
        din = vd_batch_in()
        din.n_sc = ns
        for name, t in d_in.items():
            setattr(din, name, t.data_ptr())
        din.ref_off = d_in["ref_off"].data_ptr() + 8 * cut[k]
        din.var_off = d_in["var_off"].data_ptr() + 32 * cut[k]
        din.rplane_seq = None
        din.max_qual = b.max_qual
        outs = {name: torch.full((mult * (ns if kind == "sc" else max(nv, 1)),), 77, dtype=dt, device=dev)
                for name, (dt, mult, kind) in fields.items()}
        dout = vd_batch_out()
        for name, t in outs.items():
            setattr(dout, name, t.data_ptr())
        engine.run_device_slice(din, dout, vcut[k], nv, 0, 0)
        for name, (dt, mult, kind) in fields.items():
            a = outs[name].cpu().numpy()
            got[name].append(a if kind == "sc" else a[: 2 * nv].reshape(2, nv))
    for name, (dt, mult, kind) in fields.items():
        full = np.concatenate(got[name]) if kind == "sc" else np.concatenate(got[name], axis=1).reshape(-1)
        w = want[name]
        if w.dtype.kind == "f":
            assert (full.view(np.uint32) == w.view(np.uint32)).all(), name
        else:
            assert (full.astype(np.int64) == w.astype(np.int64)).all(), name


def test_chunked_pipeline_matches_single_pass():
    """vd_run splits big batches into double-buffered chunks (H2D / kernels / D2H overlapped);
    force tiny chunks so that a small batch crosses many chunk boundaries."""
    b = synth.wgs_like(13, 7000, sv_frac=0.004, sv_max=600)
    os.environ["VD_CHUNK_SC"] = "613"
    try:
        e = capi.Engine(0)
    finally:
        del os.environ["VD_CHUNK_SC"]
    got = e.run(b)
    want = checkers.oracle_run(b)
    assert mismatches(got.trimmed(), want.trimmed(), OUT_KEYS) == {}
    st = e.stats()
    assert st["cells"] == int(b.cells().sum()) and st["n_short"] + st["n_long"] == 4 * b.n_sc
    e.close()


def test_packed_records_equal_wide_records(engine):
    """vd_run_packed (16-bit records) against vd_run on the demo golden + SV pairs, and through the host step."""
    from vcfdist_b200.batch import PackedOut
    b0, _, refB = load_golden("demo")
    for b in (b0, Batch.concat([synth.wgs_like(31, 3000), synth.sv_pairs(32, 2, 700, divergence=0.02)])):
        wide = engine.run(b).trimmed()
        pk = engine.run_packed(b)
        w = pk.widened()
        assert mismatches(w, wide, OUT_KEYS) == {}
        fin_w, fin_p = capi.finalize(b, engine.run(b)).trimmed(), capi.finalize(b, pk).trimmed()
        assert mismatches(fin_p, fin_w, FINAL_KEYS) == {}
    assert mismatches(capi.finalize(b0, engine.run_packed(b0)).trimmed(), refB, FINAL_KEYS) == {}
    assert pk.nbytes() < 0.6 * sum(getattr(engine.run(b), f).nbytes for f in OUT_KEYS)


def test_packed_records_refuse_values_that_do_not_fit(engine):
    """A supercluster with more than 16383 sync groups... is not constructible within the size limits, but a ref_ed
    of 65535 or more is: a 70 kb deletion on the truth side only."""
    rng = np.random.default_rng(3)
    ref = rng.integers(0, 4, 70_100).astype(np.uint8)
    ref = np.frombuffer(b"ACGT", np.uint8)[ref].tobytes()
    bb = BatchBuilder(max_qual=60)
    bb.add(ref, [[], [], [(20, TYPE_DEL, 70_000, b"", 30.0)], [(20, TYPE_DEL, 70_000, b"", 30.0)]])
    b = bb.build()
    try:
        wide = engine.run(b)
    except capi.VdError as e:          # shapes beyond the long path's limits are refused outright
        assert e.code in (-4, -5)
        return
    assert int(wide.aln_score[:4].max()) >= 65535 or int(wide.ref_ed.max()) >= 65535
    with pytest.raises(capi.VdError) as ei:
        engine.run_packed(b)
    assert ei.value.code == -7


def test_compact_input_gives_the_same_records(engine):
    """vd_run_compact (offsets rebuilt on the GPU from lengths, chunk cuts at the block index) against vd_run_packed, on
    the demo golden, on a batch with long superclusters, on one that needs the REF-plane string, and on a batch big enough
    for several pipeline chunks."""
    b0, _, _ = load_golden("demo")
    big = b0.take(np.resize(np.arange(b0.n_sc), 300_000))
    a = synth.adversarial(41, 1500, max_len=40)
    rp = a.ref_seq.copy()
    rp[::13] = ord("A")
    a_rp = Batch(ref_off=a.ref_off, ref_seq=a.ref_seq, var_off=a.var_off, var_pos=a.var_pos, var_rlen=a.var_rlen, var_type=a.var_type,
                 alt_off=a.alt_off, alt_seq=a.alt_seq, var_qual=a.var_qual, max_qual=a.max_qual, rplane_seq=rp)
    for b, env in ((b0, {}), (Batch.concat([synth.wgs_like(31, 3000), synth.sv_pairs(32, 2, 700, divergence=0.02)]), {}),
                   (a_rp, {}), (big, {"VD_CHUNK_SC": 65536})):
        e = engine_with(**env) if env else engine
        want = e.run_packed(b)
        got = e.run_compact(capi.compact(b))
        for f in want.FIELDS:
            x, y = getattr(got, f), getattr(want, f)
            assert (x.view(np.uint8) == y.view(np.uint8)).all(), f
        if env:
            e.close()


def test_device_records_narrowed_on_the_gpu(engine):
    """The exchange record of the multi-GPU path: vd_run_device into wide device arrays, vd_pack_device into a
    shard.PackedRecord, parsed back - against the oracle; and the overflow flag on a value that does not fit."""
    import torch
    from vcfdist_b200 import shard
    from vcfdist_b200.batch import vd_batch_in, vd_batch_out, vd_packed_out
    b = Batch.concat([synth.wgs_like(23, 5000, sv_frac=0.002, sv_max=400), synth.adversarial(24, 300, max_len=40)])
    want = checkers.oracle_run(b).trimmed()
    dev = torch.device("cuda", 0)
    d_in = {k: torch.from_numpy(getattr(b, k)).to(dev) for k in ("ref_off", "ref_seq", "var_off", "var_pos", "var_rlen", "var_type",
                                                                 "alt_off", "alt_seq", "var_qual")}
    din = vd_batch_in()
    din.n_sc = b.n_sc
    for name, t in d_in.items():
        setattr(din, name, t.data_ptr())
    din.rplane_seq = None
    din.max_qual = b.max_qual
    nv = 2 * b.n_var
    wide = {"aln_score": torch.zeros(4 * b.n_sc, dtype=torch.int32, device=dev), "status": torch.zeros(4 * b.n_sc, dtype=torch.int32, device=dev),
            "aln_end_plane": torch.zeros(4 * b.n_sc, dtype=torch.uint8, device=dev), "aln_beg_plane": torch.zeros(4 * b.n_sc, dtype=torch.uint8, device=dev),
            "assigned": torch.zeros(nv, dtype=torch.uint8, device=dev), "callq": torch.zeros(nv, dtype=torch.float32, device=dev),
            "sync_group": torch.zeros(nv, dtype=torch.int32, device=dev), "ref_ed": torch.zeros(nv, dtype=torch.int32, device=dev),
            "query_ed": torch.zeros(nv, dtype=torch.int32, device=dev)}
    dout = vd_batch_out()
    for name, t in wide.items():
        setattr(dout, name, t.data_ptr())
    rec = shard.PackedRecord(b.n_sc + 7, b.n_var + 5, dev)          # sized for a larger shard, as on a rank that is not the biggest
    rec.set_counts(b.n_sc, b.n_var)
    dp = vd_packed_out()
    for name in ("aln_score", "aln_planes", "status", "sync_group", "ref_ed", "query_ed", "callq"):
        setattr(dp, name, rec.views[name].data_ptr())
    engine.run_device(din, dout, b.n_var, b.ref_bytes, b.alt_bytes)
    engine.pack_device(dout, b.n_sc, b.n_var, dp)
    torch.cuda.synchronize()
    assert not engine.packed_overflow()
    got = rec.parse(rec.buf.view(1, -1), 0)
    for name in ("aln_score", "aln_end_plane", "aln_beg_plane", "status"):
        assert (got[name].cpu().numpy().astype(np.int64) == want[name].astype(np.int64)).all(), name
    for name in ("assigned", "sync_group", "ref_ed", "query_ed"):
        assert (got[name].cpu().numpy().reshape(-1).astype(np.int64) == want[name].astype(np.int64)).all(), name
    assert (got["callq"].cpu().numpy().reshape(-1).view(np.uint32) == want["callq"].view(np.uint32)).all()
    wide["ref_ed"][3] = 70000                                         # does not fit 16 bits
    engine.pack_device(dout, b.n_sc, b.n_var, dp)
    torch.cuda.synchronize()
    assert engine.packed_overflow()
