"""End to end through the reference's own CLI: oracle/_ref/vcfdist_ref (unmodified reference) and
oracle/_ref/vcfdist_b200cli (same object code; the hot-path call replaced by the drop-in of
vcfdist_b200/host/pr_dropin.cpp -> vd_run_packed on the GPU, and the cluster-growing stage wf_swg_cluster by
vcfdist_b200/host/cluster_dropin.cpp -> vd_wf_batch, the --distance pass edits_wrapper by
vcfdist_b200/host/edits_dropin.cpp -> vd_swg_align_batch) must write byte-identical output files.  The default clustering
method is biwfa, so the cases without -c go through the GPU clustering; superclusters.tsv then pins its result.
Both binaries are prebuilt by oracle/Makefile and travel to the GPU box."""
import filecmp
import os
import subprocess

import pytest

from conftest import ROOT
from workloads import vcfgen

REF = os.path.join(ROOT, "oracle", "_ref", "vcfdist_ref")
REFB = os.path.join(ROOT, "oracle", "_ref", "vcfdist_refB")
CLI = os.path.join(ROOT, "oracle", "_ref", "vcfdist_b200cli")

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not (os.path.exists(REF) and os.path.exists(CLI)), reason="CLI binaries not built")]

FILES = ["precision-recall.tsv", "precision-recall-summary.tsv", "query.tsv", "truth.tsv", "superclusters.tsv",
         "phase-blocks.tsv", "phasing-summary.tsv", "switchflips.tsv"]


def run(binary, q, t, fa, out, extra):
    os.makedirs(out, exist_ok=True)
    r = subprocess.run([binary, q, t, fa, "-p", out + "/", "-v", "0", *extra],
                       capture_output=True, text=True, cwd=out, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    return r


def vcf_body(path):
    return [l for l in open(path) if not l.startswith("##")]


@pytest.mark.parametrize("seed,extra", [(1, ["-c", "gap", "50"]), (2, ["-c", "size", "50", "-l", "400", "-s", "2000"]),
                                        (3, []), (4, ["-t", "3"]), (5, ["-i", "2"]), (6, ["--distance"]),
                                        (7, ["--distance", "-c", "gap", "30"])])
def test_cli_outputs_identical(tmp_path, seed, extra):
    q, t, fa = vcfgen.generate(str(tmp_path / "in"), seed=seed, contig_len=80_000 if not extra else 150_000)
    a = str(tmp_path / "ref"); b = str(tmp_path / "gpu"); c = str(tmp_path / "refB")
    run(REF, q, t, fa, a, extra)
    run(CLI, q, t, fa, b, extra)
    run(REFB, q, t, fa, c, extra)
    # order-independent files must match the unmodified reference unconditionally
    for f in ("superclusters.tsv", "phase-blocks.tsv", "phasing-summary.tsv", "switchflips.tsv"):
        assert filecmp.cmp(os.path.join(a, f), os.path.join(b, f), shallow=False), f
    # everything must match the canonical-tie-break reference byte for byte
    for f in FILES:
        assert filecmp.cmp(os.path.join(c, f), os.path.join(b, f), shallow=False), f
    assert vcf_body(os.path.join(c, "summary.vcf")) == vcf_body(os.path.join(b, "summary.vcf"))
    if "--distance" in extra:            # the --distance pass (edits_wrapper -> vd_swg_align_batch): its two files
        for f in ("distance.tsv", "distance-summary.tsv", "edits.tsv"):
            assert filecmp.cmp(os.path.join(c, f), os.path.join(b, f), shallow=False), f


def test_cli_sv_bearing_input_default_clustering(tmp_path, monkeypatch):
    """Structural variants to 600 bp with the default (biwfa) clustering and --distance: the cluster-growing stage and the
    alignment pass meet problems hundreds to thousands of diagonals wide - with the form thresholds lowered here every form
    of the wavefront kernels (warp, 256 threads, 1024 threads, cluster of eight blocks) runs inside the CLI - and the
    precision/recall stage its long path.  Same files as the reference, byte for byte."""
    q, t, fa = vcfgen.generate(str(tmp_path / "in"), seed=9, contig_len=40_000, n_contigs=2, sv_rate=0.03, sv_max=600)
    c = str(tmp_path / "refB"); b = str(tmp_path / "gpu")
    run(REFB, q, t, fa, c, ["--distance"])
    monkeypatch.setenv("VD_WF_BLOCK_MIN", "64")
    monkeypatch.setenv("VD_WF_CLUSTER_MIN", "700")
    run(CLI, q, t, fa, b, ["--distance"])
    for f in FILES + ["distance.tsv", "distance-summary.tsv", "edits.tsv"]:
        assert filecmp.cmp(os.path.join(c, f), os.path.join(b, f), shallow=False), f
    assert vcf_body(os.path.join(c, "summary.vcf")) == vcf_body(os.path.join(b, "summary.vcf"))


@pytest.mark.parametrize("name", ["demo", "adv_11", "sv_21"])
def test_dropin_at_the_function_seam(name):
    """precision_recall_threads_wrapper of the drop-in (parallel packer, vd_run_packed, host float step, parallel
    scatter) called on the reference's own superclusterData by the harness (libvdseam.so) against the recorded
    results of the reference with the canonical tie-break, twice (the handle and its page-locked arena persist)."""
    from conftest import FINAL_KEYS, load_golden, mismatches
    from oracle import checkers
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libvdseam.so")):
        pytest.skip("libvdseam.so not built")
    b, _, refB = load_golden(name)
    for threads in (1, 8):
        got, sec = checkers.reference_run(b, threads=threads, seam=True)
        assert mismatches(got, refB, FINAL_KEYS) == {}
        assert sec > 0
