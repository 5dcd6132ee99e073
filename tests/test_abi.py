"""The C-ABI library: builds, loads, exports every symbol include/vcfdist_b200.h declares,
and fails loudly (no CPU fallback) when there is no GPU.  No compute calls.  CPU only."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, has_gpu
from vcfdist_b200 import capi
from vcfdist_b200.batch import BatchBuilder, Out, TYPE_SUB


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "vcfdist_b200.h")).read()
    return sorted(set(re.findall(r"\b(vd_[a-z_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    lib = capi.load_library()
    syms = declared_symbols()
    assert set(capi.EXPORTS) == set(syms)
    for s in syms:
        assert getattr(lib, s) is not None
    assert lib.vd_abi_version() == 2


def test_struct_layouts_match_header(tmp_path):
    """The ctypes mirrors have the sizes and field offsets gcc gives the structs of include/vcfdist_b200.h."""
    import subprocess
    from vcfdist_b200 import batch
    names = ["vd_batch_in", "vd_batch_out", "vd_packed_out", "vd_compact_in", "vd_final", "vd_stats"]
    prog = ['#include <stdio.h>', '#include <stddef.h>', '#include "vcfdist_b200.h"', 'int main(void) {']
    for n in names:
        prog.append(f'printf("{n} %zu\\n", sizeof({n}));')
        for f, _ in getattr(batch, n)._fields_:
            prog.append(f'printf("{n}.{f} %zu\\n", offsetof({n}, {f}));')
    prog.append('return 0; }')
    src = tmp_path / "layout.c"
    src.write_text("\n".join(prog))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    want = dict(line.split() for line in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for n in names:
        cls = getattr(batch, n)
        assert C.sizeof(cls) == int(want[n]), n
        for f, _ in cls._fields_:
            assert getattr(cls, f).offset == int(want[f"{n}.{f}"]), (n, f)


@pytest.mark.skipif(has_gpu(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback_without_gpu():
    with pytest.raises(capi.VdError):
        capi.Engine(0)


def test_finalize_host_step():
    """store_phase and the credit thresholds (src/dist.cpp:449-475, :1293-1352) from integers."""
    bb = BatchBuilder(max_qual=60)
    bb.add(b"ACGT", [[(1, TYPE_SUB, 1, b"T", 10.0)], [], [(1, TYPE_SUB, 1, b"T", 20.0)], []])
    b = bb.build()
    out = Out(b.n_sc, b.n_var)
    out.aln_score[:4] = [0, 1, 1, 0]              # orig = 0, swap = 2 -> PHASE_ORIG
    out.assigned[:] = [2, 2, 1, 2]                # slot0: q sync, t sync; slot1: q ref-FP, t sync
    out.ref_ed[:] = [1, 1, 0, 3]
    out.query_ed[:] = [0, 0, 0, 1]                # credit 1.0, 1.0, -, 1-1/3 = 0.667 < 0.7 -> FN
    out.callq[:] = [10, 10, 10, 60]
    out.sync_group[:] = [0, 0, 0, 0]
    fin = capi.finalize(b, out).trimmed()
    assert list(fin["sc_phase"]) == [0] and list(fin["orig_dist"]) == [0] and list(fin["swap_dist"]) == [2]
    assert list(fin["errtypes"]) == [0, 0, 1, 2]
    assert fin["credit"][3] == np.float32(1) - np.float32(1) / np.float32(3)
    assert fin["callq"][3] == 60.0 and fin["callq"][2] == 10.0


def test_finalize_packed_equals_finalize():
    """vd_finalize_packed over the 16-bit records = vd_finalize over the wide ones (random records, multi-threaded
    split included: > 200 k superclusters)."""
    from vcfdist_b200.batch import PackedOut
    from workloads import synth
    b = synth.wgs_like(5, 210_000)
    rng = np.random.default_rng(7)
    out, pk = Out(b.n_sc, b.n_var), PackedOut(b.n_sc, b.n_var)
    a, v = 4 * b.n_sc, 2 * b.n_var
    out.aln_score[:a] = rng.integers(-1, 40, a)
    out.assigned[:v] = rng.integers(0, 3, v)
    out.sync_group[:v] = rng.integers(0, 9, v)
    out.ref_ed[:v] = rng.integers(1, 12, v)
    out.query_ed[:v] = rng.integers(0, 12, v)
    out.callq[:v] = rng.integers(0, 60, v).astype(np.float32)
    pk.aln_score[:a] = np.where(out.aln_score[:a] < 0, 0xFFFF, out.aln_score[:a]).astype(np.uint16)
    pk.sync_group[:v] = (out.assigned[:v].astype(np.uint16) << 14) | out.sync_group[:v].astype(np.uint16)
    pk.ref_ed[:v] = out.ref_ed[:v]; pk.query_ed[:v] = out.query_ed[:v]; pk.callq[:v] = out.callq[:v]
    f1, f2 = capi.finalize(b, out).trimmed(), capi.finalize(b, pk).trimmed()
    for k in f1:
        assert (f1[k].view(np.uint8) == f2[k].view(np.uint8)).all(), k
    w = pk.widened()
    assert (w["aln_score"] == out.aln_score[:a]).all() and (w["assigned"] == out.assigned[:v]).all()


def test_compact_form_of_a_batch():
    """vd_compact_pack (host only): lengths whose prefix sums are the batch's offsets, 16-bit positions, the block index the
    host pipeline cuts at; VD_E_RANGE when a window does not fit 16 bits."""
    from workloads import synth
    from vcfdist_b200.batch import COMPACT_BLOCK
    b = synth.wgs_like(9, 150_000, sv_frac=0.001, sv_max=3000)      # (the packer does not validate: plan_kernel does)
    ci = capi.compact(b)
    assert (np.concatenate([[0], np.cumsum(ci.ref_len[: b.n_sc].astype(np.int64))]) == b.ref_off).all()
    assert (np.concatenate([[0], np.cumsum(ci.hap_nvar[: 4 * b.n_sc].astype(np.int64))]) == b.var_off).all()
    assert (np.concatenate([[0], np.cumsum(ci.alt_len[: b.n_var].astype(np.int64))]) == b.alt_off).all()
    assert (ci.var_pos[: b.n_var] == b.var_pos[: b.n_var]).all() and (ci.var_rlen[: b.n_var] == b.var_rlen[: b.n_var]).all()
    blk = np.arange(0, b.n_sc, COMPACT_BLOCK)
    assert (ci.blk_ref[: len(blk)] == b.ref_off[blk]).all() and ci.blk_ref[len(blk)] == b.ref_bytes
    assert (ci.blk_var[: len(blk)] == b.var_off[4 * blk]).all() and ci.blk_var[len(blk)] == b.n_var
    assert (ci.blk_alt[: len(blk)] == b.alt_off[b.var_off[4 * blk]]).all() and ci.blk_alt[len(blk)] == b.alt_bytes
    assert ci.c.n_sc == b.n_sc and ci.c.n_var == b.n_var and ci.nbytes() < 0.55 * b.io_bytes()
    bb = BatchBuilder(max_qual=60)
    bb.add(b"ACGT" * 20000, [[(5, TYPE_SUB, 1, b"T", 10.0)], [], [], []])
    with pytest.raises(capi.VdError) as ei:
        capi.compact(bb.build())
    assert ei.value.code == -7


def test_product_package_never_touches_the_oracle():
    """The checkers under oracle/ are test infrastructure: no module of the product package may import,
    load or name them (docstrings that say so excepted), and importing the package must not pull
    oracle.checkers in."""
    import ast
    import subprocess
    import sys
    pkg = os.path.join(ROOT, "vcfdist_b200")
    for fn in sorted(os.listdir(pkg)):
        if not fn.endswith(".py"):
            continue
        tree = ast.parse(open(os.path.join(pkg, fn)).read())
        for node in ast.walk(tree):
            if isinstance(node, (ast.Import, ast.ImportFrom)):
                names = [a.name for a in node.names] + [getattr(node, "module", "") or ""]
                assert not any(n.split(".")[0] == "oracle" for n in names), (fn, names)
            if isinstance(node, ast.Constant) and isinstance(node.value, str) and node.value is not ast.get_docstring(tree, clean=False):
                assert "liboracle" not in node.value and "libvdref" not in node.value, (fn, node.value[:60])
    code = "import sys; import vcfdist_b200, vcfdist_b200.capi, vcfdist_b200.shard, vcfdist_b200.batch; " \
           "sys.exit(1 if any(m.startswith('oracle') for m in sys.modules) else 0)"
    assert subprocess.run([sys.executable, "-c", code], cwd=ROOT).returncode == 0
