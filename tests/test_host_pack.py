"""The drop-in's host packer (vcfdist_b200/host/pr_dropin.cpp: two parallel passes over the reference's
superclusterData) on CPU: oracle/_ref/vcfdist_dump is the reference CLI with that packer linked in and the REFERENCE
computing the results; the packed batch it dumps for the bundled demo must be the committed demo golden byte for
byte, on one host thread and on eight.  Needs the reference's demo files, i.e. runs in the build container only."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from workloads import synth

DEMO = "/root/reference/demo"
EXE = os.path.join(ROOT, "oracle", "_ref", "vcfdist_dump")


@pytest.mark.skipif(not (os.path.isdir(DEMO) and os.path.exists(EXE)), reason="reference demo or vcfdist_dump missing")
@pytest.mark.parametrize("threads", [1, 8])
def test_parallel_packer_reproduces_demo_golden(tmp_path, threads):
    z = np.load(os.path.join(ROOT, "tests", "golden", "demo.npz"))
    env = dict(os.environ, VD_DUMP_BATCH=str(tmp_path / "batch.vdarr"), VD_DUMP_FINAL=str(tmp_path / "final.vdarr"))
    subprocess.run([EXE, f"{DEMO}/query.vcf", f"{DEMO}/nist-v4.2.1_chr1_5Mb.vcf.gz", f"{DEMO}/GRCh38_chr1_5Mb.fa",
                    "-b", f"{DEMO}/nist-v4.2.1_chr1_5Mb.bed", "-p", str(tmp_path / "out") + "/", "-v", "0", "-t", str(threads)],
                   check=True, env=env, cwd=tmp_path, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    b = synth.batch_from_vdarr(str(tmp_path / "batch.vdarr"))
    for k in ("ref_off", "ref_seq", "var_off", "var_pos", "var_rlen", "var_type", "alt_off", "alt_seq", "var_qual"):
        want = z[k]
        assert (getattr(b, k)[: len(want)] == want).all(), k
    assert (b.rplane_seq is not None) == ("rplane_seq" in z.files)
    fin = synth.read_vdarr(str(tmp_path / "final.vdarr"))          # the reference's results, gathered back through vb[]
    for k, v in fin.items():
        assert (v.view(np.uint8) == z["refA_" + k].view(np.uint8)).all(), k
