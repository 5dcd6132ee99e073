"""Regenerates tests/golden/*.npz from the REFERENCE itself (run in the build container,
where /root/reference exists and `make -C oracle` has built oracle/_ref/).

  demo.npz        the bundled demo (demo/query.vcf vs demo/nist-v4.2.1_chr1_5Mb.vcf.gz,
                  GRCh38_chr1_5Mb.fa, BED) taken through the reference's own parser, clustering
                  and superclustering by oracle/_ref/vcfdist_dump: the packed hot-path input of
                  all 6037 superclusters and the unmodified reference's results for it
  adv_*.npz       seeded adversarial batches (workloads.synth.adversarial) with the results
                  of the unmodified reference (refA_*) and of the canonical-tie-break reference
                  (refB_*), both run through oracle/_ref/libvdref[B].so
  sv_*.npz        a few superclusters with long insertions/deletions (long-kernel shapes)
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from vcfdist_b200 import capi
from workloads import synth  # noqa: E402
from oracle import checkers  # noqa: E402
from vcfdist_b200.batch import Batch  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
DEMO = "/root/reference/demo"


def batch_arrays(b: Batch) -> dict:
    d = dict(ref_off=b.ref_off, ref_seq=b.ref_seq[: b.ref_bytes], var_off=b.var_off,
             var_pos=b.var_pos[: b.n_var], var_rlen=b.var_rlen[: b.n_var], var_type=b.var_type[: b.n_var],
             alt_off=b.alt_off, alt_seq=b.alt_seq[: b.alt_bytes], var_qual=b.var_qual[: b.n_var],
             max_qual=np.float32(b.max_qual))
    if b.rplane_seq is not None:
        d["rplane_seq"] = b.rplane_seq
    return d


def save(name, b, **refs):
    d = batch_arrays(b)
    for prefix, res in refs.items():
        for k, v in res.items():
            d[f"{prefix}_{k}"] = v
    np.savez_compressed(os.path.join(HERE, name), **d)
    print(name, "n_sc", b.n_sc, "n_var", b.n_var, "cells", int(b.cells().sum()))


def main():
    # ---- demo through the reference CLI host pipeline ----
    with tempfile.TemporaryDirectory() as tmp:
        env = dict(os.environ, VD_DUMP_BATCH=f"{tmp}/batch.vdarr", VD_DUMP_FINAL=f"{tmp}/final.vdarr")
        subprocess.run([os.path.join(ROOT, "oracle/_ref/vcfdist_dump"), f"{DEMO}/query.vcf",
                        f"{DEMO}/nist-v4.2.1_chr1_5Mb.vcf.gz", f"{DEMO}/GRCh38_chr1_5Mb.fa",
                        "-b", f"{DEMO}/nist-v4.2.1_chr1_5Mb.bed", "-p", f"{tmp}/out/", "-v", "0"],
                       check=True, env=env, cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        b = synth.batch_from_vdarr(f"{tmp}/batch.vdarr")
        refA = synth.read_vdarr(f"{tmp}/final.vdarr")
    refB, _ = checkers.reference_run(b, canonical=True, threads=8)
    save("demo.npz", b, refA=refA, refB=refB)

    # ---- adversarial ----
    for seed, n, mx in ((11, 1200, 28), (12, 600, 60)):
        b = synth.adversarial(seed, n, max_len=mx)
        rA, _ = checkers.reference_run(b, canonical=False, threads=8)
        rB, _ = checkers.reference_run(b, canonical=True, threads=8)
        save(f"adv_{seed}.npz", b, refA=rA, refB=rB)

    # ---- long shapes ----
    b = Batch.concat([synth.wgs_like(21, 60, sv_frac=0.5, sv_min=50, sv_max=600),
                      synth.sv_pairs(22, 2, 900, divergence=0.02)])
    rA, _ = checkers.reference_run(b, canonical=False, threads=8)
    rB, _ = checkers.reference_run(b, canonical=True, threads=8)
    save("sv_21.npz", b, refA=rA, refB=rB)
    make_sv_10k()


def make_sv_10k():
    """10 kb structural variants (BASELINE configs[3]) through the reference's own object code."""
    cases = [("ins", "het", 0.01), ("ins", "mixed", 0.0), ("del", "hom", 0.01), ("ins_truth_only", "het", 0.0),
             ("ins_query_only", "cross", 0.0), ("del_truth_only", "het", 0.0), ("del_query_only", "het", 0.0)]
    b = Batch.concat([synth.sv_case(500 + i, 10000, k, z, d) for i, (k, z, d) in enumerate(cases)])
    rA, sa = checkers.reference_run(b, canonical=False, threads=8)
    rB, sb = checkers.reference_run(b, canonical=True, threads=8)
    print("reference seconds", sa, sb)
    save("sv_10k.npz", b, refA=rA, refB=rB)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "sv_10k":
        make_sv_10k()
    else:
        main()
