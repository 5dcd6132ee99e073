"""Seeded synthetic FASTA + truth/query VCF files, so that the reference CLI and the CLI with the
GPU drop-in can be run end to end on the same inputs (there is no HG002 data on the GPU box).
VCF-level generation as in SURVEY.md 8d: phased GT:GQ, QUAL uniform 3-50, SNPs, small indels,
a few SVs, tandem-repeat patches; the query is the truth with drop-outs, extras, re-phased
blocks and shifted (equivalent) indel representations."""
from __future__ import annotations

import os
from typing import List, Tuple

import numpy as np

ALPH = b"ACGT"


def _seq(rng, n):
    return bytes(np.frombuffer(ALPH, np.uint8)[rng.integers(0, 4, n)])


def make_reference(rng, length: int) -> bytearray:
    ref = bytearray(_seq(rng, length))
    # tandem-repeat patches (unit 1-3, ~60 bp) every ~2 kb
    for start in range(500, length - 200, 2000):
        unit = _seq(rng, int(rng.integers(1, 4)))
        n = int(rng.integers(30, 70))
        ref[start:start + n] = (unit * (n // len(unit) + 1))[:n]
    return ref


def _variants(rng, ref: bytes, mean_gap: float, sv_rate: float, max_indel: int, sv_max: int):
    """[(pos0, REF, ALT, gt)] sorted, non-overlapping, VCF-style anchored indels."""
    out = []
    pos = int(rng.integers(100, 200))
    L = len(ref)
    while pos < L - sv_max - 200:
        r = rng.random()
        gt = ["0|1", "1|0", "1|1"][int(rng.integers(0, 3))]
        if r < 0.7:
            alt = ALPH[(ALPH.index(ref[pos]) + 1 + int(rng.integers(0, 3))) % 4]
            out.append((pos, bytes([ref[pos]]), bytes([alt]), gt))
            end = pos + 1
        else:
            n = int(rng.integers(1, max_indel + 1))
            if rng.random() < sv_rate:
                n = int(rng.integers(50, sv_max + 1))
            if r < 0.85:
                out.append((pos, bytes([ref[pos]]), bytes([ref[pos]]) + _seq(rng, n), gt))
                end = pos + 1
            else:
                out.append((pos, bytes(ref[pos:pos + n + 1]), bytes([ref[pos]]), gt))
                end = pos + n + 1
        gap = int(rng.exponential(mean_gap)) + 2
        if rng.random() < 0.08:
            gap += int(rng.integers(200, 600))      # cluster boundary
        pos = end + gap
    return out


def _derive_query(rng, ref: bytes, truth):
    out = []
    for (pos, r, a, gt) in truth:
        x = rng.random()
        if x < 0.03:
            continue                                  # FN
        if x < 0.06:
            gt = {"0|1": "1|0", "1|0": "0|1", "1|1": "1|1"}[gt]     # phase flip
        if x > 0.97 and len(r) == 1 and len(a) == 1:
            a = bytes([ALPH[(ALPH.index(a[0]) + 1) % 4]]) if a[0] in ALPH else a
            if a == r:
                continue
        out.append((pos, r, a, gt))
        if rng.random() < 0.02:                       # FP SNP nearby
            p2 = pos + len(r) + int(rng.integers(3, 20))
            if p2 < len(ref) - 10:
                alt = ALPH[(ALPH.index(ref[p2]) + 1) % 4]
                out.append((p2, bytes([ref[p2]]), bytes([alt]), "0|1"))
    out.sort(key=lambda v: v[0])
    # drop anything overlapping its predecessor
    keep, end = [], -1
    for v in out:
        if v[0] > end:
            keep.append(v)
            end = v[0] + len(v[1])
    return keep


def _write_vcf(path, contigs, records, sample, rng):
    with open(path, "w") as f:
        f.write("##fileformat=VCFv4.2\n")
        f.write('##FILTER=<ID=PASS,Description="All filters passed">\n')
        f.write('##FORMAT=<ID=GT,Number=1,Type=String,Description="Genotype">\n')
        f.write('##FORMAT=<ID=GQ,Number=1,Type=Integer,Description="Genotype quality">\n')
        for name, length in contigs:
            f.write(f"##contig=<ID={name},length={length}>\n")
        f.write(f"#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t{sample}\n")
        for name, recs in records:
            for (pos, r, a, gt) in recs:
                q = rng.uniform(3, 50)
                f.write(f"{name}\t{pos + 1}\t.\t{r.decode()}\t{a.decode()}\t{q:.2f}\tPASS\t.\tGT:GQ\t{gt}:{int(rng.integers(5, 99))}\n")


def generate(outdir: str, seed: int = 1, contig_len: int = 150_000, n_contigs: int = 2,
             mean_gap: float = 40.0, sv_rate: float = 0.01, max_indel: int = 20, sv_max: int = 300) -> Tuple[str, str, str]:
    """Writes ref.fa, truth.vcf, query.vcf into outdir; returns their paths."""
    os.makedirs(outdir, exist_ok=True)
    rng = np.random.default_rng(seed)
    contigs, trecs, qrecs = [], [], []
    fa = os.path.join(outdir, "ref.fa")
    with open(fa, "w") as f:
        for c in range(n_contigs):
            name = f"chr{c + 1}"
            ref = bytes(make_reference(rng, contig_len))
            f.write(f">{name}\n")
            for i in range(0, len(ref), 60):
                f.write(ref[i:i + 60].decode() + "\n")
            contigs.append((name, len(ref)))
            t = _variants(rng, ref, mean_gap, sv_rate, max_indel, sv_max)
            trecs.append((name, t))
            qrecs.append((name, _derive_query(rng, ref, t)))
    tv, qv = os.path.join(outdir, "truth.vcf"), os.path.join(outdir, "query.vcf")
    _write_vcf(tv, contigs, trecs, "TRUTH", rng)
    _write_vcf(qv, contigs, qrecs, "QUERY", rng)
    return qv, tv, fa
