"""Pins the oracle against the reference's own object code (oracle/_ref/libvdref[B].so, built
by oracle/Makefile from the unmodified sources) on fresh seeded batches.  CPU only; skipped
when the reference build is not present."""
import numpy as np
import pytest

from conftest import FINAL_KEYS, mismatches, non_tie_var_mask
from vcfdist_b200 import capi
from workloads import synth
from oracle import checkers

pytestmark = pytest.mark.skipif(not (checkers.reference_available(False) and checkers.reference_available(True)),
                                reason="oracle/_ref/libvdref*.so not built")


@pytest.mark.parametrize("seed,n,mx", [(101, 400, 24), (102, 300, 48), (103, 150, 80)])
def test_adversarial_vs_reference(seed, n, mx, capfd):
    b = synth.adversarial(seed, n, max_len=mx)
    out = checkers.oracle_run(b)
    fin = capi.finalize(b, out).trimmed()
    refB, _ = checkers.reference_run(b, canonical=True, threads=4)
    refA, _ = checkers.reference_run(b, canonical=False, threads=4)
    capfd.readouterr()          # the reference prints its data WARNs
    assert mismatches(fin, refB, FINAL_KEYS) == {}
    mask, _ = non_tie_var_mask(b, out.status)
    assert mismatches(fin, refA, FINAL_KEYS, var_mask=mask) == {}


def test_wgs_like_vs_reference(capfd):
    b = synth.wgs_like(7, 3000, sv_frac=0.01, sv_max=400)
    out = checkers.oracle_run(b)
    fin = capi.finalize(b, out).trimmed()
    refB, _ = checkers.reference_run(b, canonical=True, threads=4)
    capfd.readouterr()
    assert mismatches(fin, refB, FINAL_KEYS) == {}


def test_rplane_override(capfd):
    """VCF REF alleles of query hap 1 that differ from the FASTA change the REF-plane string
    (ref_q1, src/dist.cpp:187,195,1784-1792)."""
    from vcfdist_b200.batch import BatchBuilder, TYPE_DEL, TYPE_SUB
    bb = BatchBuilder()
    ref = b"ACGTACGTAC"
    rplane = b"ACGAACGTAC"          # q1's SUB at pos 3 claims REF 'A' instead of 'T'
    q1 = [(3, TYPE_SUB, 1, b"C", 20.0)]
    t1 = [(3, TYPE_SUB, 1, b"C", 30.0), (6, TYPE_DEL, 2, b"", 30.0)]
    bb.add(ref, [q1, [], t1, []], rplane=rplane)
    b = bb.build()
    out = checkers.oracle_run(b)
    fin = capi.finalize(b, out).trimmed()
    refB, _ = checkers.reference_run(b, canonical=True, threads=1)
    capfd.readouterr()
    assert mismatches(fin, refB, FINAL_KEYS) == {}
