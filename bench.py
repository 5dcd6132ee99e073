#!/usr/bin/env python
"""bench.py — DP Gcells/s and superclusters/s of the precision/recall hot path.

A "step" is one pass of the whole hot path (plan, fused short-supercluster kernel, wavefront
kernels for long superclusters, credit) over one batch of synthetic superclusters:

  wgs     (default) BASELINE.json configs[2]: HG002-WGS-like small variants — 3.6 M superclusters
          drawn with replacement (seeded) from the real HG002 chr1:1-5Mb demo batch in
          tests/golden/demo.npz, i.e. the measured WGS length distribution
  wgs_sv  configs[3]: the same plus an SV tail (0.3 % of superclusters carry one INS/DEL of
          log-uniform length 50..10 kb)
  chr20   configs[1]: 75 k superclusters

  python bench.py --gpus N --steps K --warmup W           (torchrun for N > 1, one rank per GPU)
  python bench.py --impl reference ...                    the reference's own CPU code on a
                                                          bounded sample, all host threads

value   = whole-job Gcells/s with inputs resident in HBM (vd_run_device), device-timed,
          max over ranks; cells = sum over alignments of (Lq+Lr)*Lt
e2e     = the same metric through the reference-facing C-ABI call vd_run() with pinned HOST
          buffers: H2D of the batch and D2H of the results inside the timed region
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from vcfdist_b200 import capi, shard
from workloads import synth  # noqa: E402
from oracle import checkers  # noqa: E402  (CPU-baseline legs only: the reference arm and cpu_baseline)
from vcfdist_b200.batch import Batch, Out, PackedOut, vd_batch_in, vd_batch_out, vd_packed_out  # noqa: E402

WORKLOADS = {
    "wgs": dict(n_sc=3_600_000, sv_frac=0.0, config="HG002 WGS SNP+INDEL vs GIAB v4.2.1, 1xB200 (BASELINE configs[2])"),
    "wgs_sv": dict(n_sc=3_600_000, sv_frac=0.003, config="HG002 WGS SNP+INDEL+SV to 10 Kb (BASELINE configs[3])"),
    "chr20": dict(n_sc=75_000, sv_frac=0.0, config="HG002 chr20 SNP+INDEL (BASELINE configs[1])"),
}


def load_demo() -> Batch:
    z = np.load(os.path.join(ROOT, "tests", "golden", "demo.npz"))
    return Batch(ref_off=z["ref_off"], ref_seq=z["ref_seq"], var_off=z["var_off"], var_pos=z["var_pos"],
                 var_rlen=z["var_rlen"], var_type=z["var_type"], alt_off=z["alt_off"], alt_seq=z["alt_seq"],
                 var_qual=z["var_qual"], max_qual=float(z["max_qual"]))


def make_workload(name: str, n_sc: int, seed: int, rank: int, world: int, sv_max: int):
    """This rank's shard of the global batch (weak scaling: n_sc superclusters per GPU, so the
    global batch has world*n_sc), partitioned by estimated cells (LPT).  Every rank derives the
    same global index list; only its own shard is materialised."""
    w = WORKLOADS[name]
    demo = load_demo()
    rng = np.random.default_rng(seed)
    total = n_sc * world
    pick = rng.integers(0, demo.n_sc, total)
    base_cells = demo.cells()
    cells = base_cells[pick]
    n_svs = int(round(total * w["sv_frac"]))
    sv = None
    if n_svs:
        sv = synth.wgs_like(seed + 1, n_svs, sv_frac=1.0, sv_max=sv_max)
        cells = np.concatenate([cells[: total - n_svs], sv.cells()])
    parts = shard.lpt_partition(cells, world) if world > 1 else [np.arange(total)]
    mine = parts[rank]
    small = mine[mine < total - n_svs]
    big = mine[mine >= total - n_svs] - (total - n_svs)
    pieces = [demo.take(pick[small])]
    if n_svs and len(big):
        pieces.append(sv.take(big))
    b = Batch.concat(pieces) if len(pieces) > 1 else pieces[0]
    return b, int(cells.sum()), total


class ClockSampler:
    """SM clock and throttle reasons during the timed region (B200_PROFILING.md), read through NVML every 2 ms by a
    thread of this process (`nvidia-smi -lms` needs ~100 ms to start, longer than a short timed region)."""
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, gpu: int):
        self.gpu, self.sm, self.mask, self.max_mhz, self.stop_flag, self.t = gpu, [], 0, None, False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[gpu]) if vis and vis.split(",")[gpu].isdigit() else gpu
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _loop(self):
        while not self.stop_flag:
            try:
                self.sm.append(float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                self.mask |= int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.nv:
            self.t = threading.Thread(target=self._loop, daemon=True)
            self.t.start()

    def stop(self) -> dict:
        if not self.nv:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["NVML unavailable"], "samples": 0}
        self.stop_flag = True
        self.t.join(timeout=1.0)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": [n for n, bit in self.REASONS if self.mask & bit], "samples": len(self.sm), "source": "NVML, 2 ms period"}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def run_reference(args):
    """The reference's own CPU implementation (oracle/_ref/libvdref.so = unmodified sources) of
    the same path, all host threads, on a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    if not checkers.reference_available(False):
        # the C restatement stands in when the reference objects were not built
        kind = "port"
    else:
        kind = "reference"
    # the SAME batch the GPU arm runs (--ref-sample 0, the default); a smaller sample only on request
    sample_n = args.ref_sample or args.n_sc or WORKLOADS[args.workload]["n_sc"]
    b, cells, total = make_workload(args.workload, sample_n, args.seed, 0, 1, args.sv_max)
    times = []
    for i in range(args.warmup + args.steps):
        if kind == "reference":
            _, sec = checkers.reference_run(b, canonical=False, threads=cores)
        else:
            t0 = time.perf_counter(); checkers.oracle_run(b); sec = time.perf_counter() - t0
            cores = 1
        if i >= args.warmup:
            times.append(sec)
    sec = float(np.mean(times))
    val = cells / sec / 1e9
    line = {
        "impl": "reference", "metric": "dp_gcells_per_s", "value": val, "unit": "Gcells/s",
        "superclusters_per_s": b.n_sc / sec, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int32", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.workload]["config"], "n_superclusters_per_step": b.n_sc,
                   "cells_per_step": cells},
        "cpu_baseline": {"value": val, "unit": "Gcells/s", "cores": cores, "kind": kind,
                         "sample": f"{b.n_sc} superclusters ({cells} cells) of the {args.workload} workload per step, "
                                   f"timed inside precision_recall_threads_wrapper (-t {cores})"},
        "e2e": {"value": val, "unit": "Gcells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def tie_count(status, n_sc):
    """Superclusters in which an alignment followed an ambiguous swap edge (VD_ST_TIE): the reference's choice there
    depends on hash-set iteration order, ours is canonical."""
    return int((status[: 4 * n_sc].reshape(-1, 4) & 1).any(axis=1).sum())


def cli_timer(stderr: str, idx: int) -> float:
    """Seconds of timer [idx] in the reference CLI's own end-of-run report (src/timer.cpp:29-31)."""
    import re
    m = re.search(r"\[%d\] [a-z/ ]+:\s*([0-9.]+)s" % idx, stderr)
    return float(m.group(1)) if m else float("nan")


def run_seam(args, b):
    """The product where a vcfdist user meets it: the reference's own timer around precision_recall_threads_wrapper
    (g.timers[TIME_PR_ALN], src/main.cpp:218-221).  (1) function seam: the harness rebuilds the reference's
    superclusterData for the bench batch and calls the wrapper, once resolved to the reference's own definition
    (that is the reference arm / cpu_baseline) and once to the GPU drop-in (libvdseam.so): pack + vd_run_packed +
    float step + scatter, host threads = -t.  (2) the reference CLI and the CLI with the drop-in linked in
     on a seeded synthetic VCF pair, each reporting its own '[5] precision/recall' timer."""
    cores = os.cpu_count() or 1
    res = {}
    if checkers.seam_available():
        secs, bds = [], []
        for _ in range(3):
            _, sec = checkers.reference_run(b, threads=cores, seam=True)
            secs.append(sec); bds.append(checkers.seam_breakdown())
        best = int(np.argmin(secs[1:])) + 1
        res["function_seam"] = {"api": "precision_recall_threads_wrapper (drop-in, vcfdist_b200/host/pr_dropin.cpp) on the reference's superclusterData",
                                "n_superclusters": b.n_sc, "ms_first_call": secs[0] * 1e3, "ms_per_call": secs[best] * 1e3,
                                "breakdown_ms": bds[best], "host_threads": cores}
    ref, cli = (os.path.join(ROOT, "oracle", "_ref", n) for n in ("vcfdist_ref", "vcfdist_b200cli"))
    if os.path.exists(ref) and os.path.exists(cli) and (args.cli_contig_len > 0 or args.cli_cluster_contig_len > 0):
        import tempfile
        from workloads import vcfgen
        with tempfile.TemporaryDirectory() as tmp:
            def run_cli(exe, od, q_, t_, fa_, extra, **env):
                os.makedirs(od, exist_ok=True)
                t0 = time.perf_counter()
                r = subprocess.run([exe, q_, t_, fa_, "-p", od + "/", "-v", "1", "-t", str(cores), *extra],
                                   capture_output=True, text=True, cwd=od, env=dict(os.environ, VD_DROPIN_TIMES="1", **env))
                res_ = {"rc": r.returncode, "wall_s": time.perf_counter() - t0, "timer_3_reclustering_s": cli_timer(r.stderr, 3),
                        "timer_5_precision_recall_s": cli_timer(r.stderr, 5), "timer_6_edit_distance_s": cli_timer(r.stderr, 6),
                        "timer_9_total_s": cli_timer(r.stderr, 9)}
                bd = [ln.split("GPU precision/recall:")[1].strip() for ln in r.stderr.splitlines() if "GPU precision/recall:" in ln]
                if bd:
                    res_["breakdown"] = bd[-1]
                return res_
            if args.cli_contig_len > 0:
                q, t, fa = vcfgen.generate(os.path.join(tmp, "in"), seed=args.seed, contig_len=args.cli_contig_len, n_contigs=2)
                out = {name: run_cli(exe, os.path.join(tmp, name), q, t, fa, ["-c", "gap", "50"]) for name, exe in (("reference", ref), ("drop_in", cli))}
                def same(files):
                    return bool(all(open(os.path.join(tmp, "reference", f)).read() == open(os.path.join(tmp, "drop_in", f)).read() for f in files))
                res["cli"] = {"input": f"workloads.vcfgen seed {args.seed}, 2 contigs x {args.cli_contig_len} bp, -c gap 50, -t {cores}",
                              "reference": out["reference"], "drop_in": out["drop_in"],
                              # the tie rule can move single variants between TP and FP against the UNMODIFIED reference;
                              # tests/test_cli_dropin.py holds the byte-for-byte comparison against the canonical-tie-break build
                              "order_independent_files_identical": same(("superclusters.tsv", "phase-blocks.tsv", "phasing-summary.tsv", "switchflips.tsv")),
                              "precision_recall_summary_identical": same(("precision-recall-summary.tsv",))}
            # the default clustering (biwfa, wf_swg_cluster): the stage the cluster drop-in replaces ([3] reclustering)
            if args.cli_cluster_contig_len > 0:
                q2, t2, fa2 = vcfgen.generate(os.path.join(tmp, "in2"), seed=args.seed + 1, contig_len=args.cli_cluster_contig_len, n_contigs=2)
                out2 = {name: run_cli(exe, os.path.join(tmp, name + "_biwfa"), q2, t2, fa2, ["--distance"]) for name, exe in (("reference", ref), ("drop_in", cli))}
                out2["drop_in_clustering_on_the_cpu_until_cuda_is_up"] = run_cli(cli, os.path.join(tmp, "drop_in_biwfa_nowait"), q2, t2, fa2, ["--distance"],
                                                                                  VD_GPU_CLUSTER_NOWAIT="1")
                def same2(f):
                    return bool(open(os.path.join(tmp, "reference_biwfa", f)).read() == open(os.path.join(tmp, "drop_in_biwfa", f)).read())
                res["cli_biwfa_clustering_and_distance"] = {
                    "input": f"workloads.vcfgen seed {args.seed + 1}, 2 contigs x {args.cli_cluster_contig_len} bp, default clustering (biwfa), --distance, -t {cores}",
                    "note": "[3] reclustering = wf_swg_cluster (drop-in: vd_wf_batch; includes waiting for CUDA start-up, which this stage "
                            "meets first), [5] precision/recall, [6] edit distance = edits_wrapper (drop-in: vd_swg_align_batch)",
                    **out2, "superclusters_identical": same2("superclusters.tsv"), "distance_identical": same2("distance.tsv"),
                    "edits_identical": same2("edits.tsv")}
    return res


def run_seeds(eng, args, dev, torch, first_seed_ms):
    """The resident step on four more draws of the workload (seeds 2..5 beside the line's own seed): per-seed ms per step,
    median and spread (SURVEY.md 8d asks for five seeds)."""
    out = {str(args.seed): first_seed_ms}
    for seed in range(args.seed + 1, args.seed + 5):
        b, cells, _ = make_workload(args.workload, args.n_sc or WORKLOADS[args.workload]["n_sc"], seed, 0, 1, args.sv_max)
        t_in = {k: torch.from_numpy(getattr(b, k)).to(dev) for k in ("ref_off", "ref_seq", "var_off", "var_pos", "var_rlen", "var_type",
                                                                     "alt_off", "alt_seq", "var_qual")}
        din = vd_batch_in()
        din.n_sc = b.n_sc
        for k, t in t_in.items():
            setattr(din, k, t.data_ptr())
        din.rplane_seq = None
        din.max_qual = b.max_qual
        nv = max(2 * b.n_var, 1)
        t_out = {"aln_score": torch.zeros(4 * b.n_sc, dtype=torch.int32, device=dev), "status": torch.zeros(4 * b.n_sc, dtype=torch.int32, device=dev),
                 "aln_end_plane": torch.zeros(4 * b.n_sc, dtype=torch.uint8, device=dev), "aln_beg_plane": torch.zeros(4 * b.n_sc, dtype=torch.uint8, device=dev),
                 "assigned": torch.zeros(nv, dtype=torch.uint8, device=dev), "callq": torch.zeros(nv, dtype=torch.float32, device=dev),
                 "sync_group": torch.zeros(nv, dtype=torch.int32, device=dev), "ref_ed": torch.zeros(nv, dtype=torch.int32, device=dev),
                 "query_ed": torch.zeros(nv, dtype=torch.int32, device=dev)}
        dout = vd_batch_out()
        for k, t in t_out.items():
            setattr(dout, k, t.data_ptr())
        torch.cuda.synchronize()
        for _ in range(3):
            eng.run_device(din, dout, b.n_var, b.ref_bytes, b.alt_bytes)
        ms = []
        for _ in range(5):
            eng.run_device(din, dout, b.n_var, b.ref_bytes, b.alt_bytes)
            ms.append(float(eng.stats()["ms_total"]))
        out[str(seed)] = {"ms_per_step": float(np.mean(ms)), "gcells_per_s": cells / (float(np.mean(ms)) * 1e-3) / 1e9, "cells": cells}
        del t_in, t_out
    vals = [v["ms_per_step"] for v in out.values()]
    return {"per_seed": out, "median_ms_per_step": float(np.median(vals)), "min_ms_per_step": float(min(vals)), "max_ms_per_step": float(max(vals)),
            "timing": "vd_stats.ms_total (CUDA events inside vd_run_device), 3 warm-up + 5 timed passes per seed"}


def run_secondary(eng, args, dev, torch):
    """BASELINE configs[3] (WGS + SV tail to 10 kb) at FULL scale through vd_run with host buffers: 3.6 M
    superclusters of the demo mixture of which 0.3 % carry one 50 bp..10 kb INS/DEL, with its own CPU baseline
    (the reference's object code on a bounded sample of the same workload) and tie count."""
    b, cells, n = make_workload("wgs_sv", args.sv_n_sc, args.seed, 0, 1, args.sv_max)
    out = None
    for _ in range(2):
        out = eng.run(b, out)
    ms, dev_ms = [], []
    kern = {"long_fwd": 0.0, "long_bwd": 0.0, "long_walk": 0.0, "short": 0.0, "band_phase_wall": 0.0, "long_wall": 0.0}
    R = 3
    for _ in range(R):
        t0 = time.perf_counter()
        out = eng.run(b, out)
        ms.append((time.perf_counter() - t0) * 1e3)
        st = eng.stats()
        dev_ms.append(float(st["ms_total"]))
        kern["long_fwd"] += st["ms_long_fwd"] / R; kern["long_bwd"] += st["ms_long_bwd"] / R
        kern["long_walk"] += st["ms_long_walk"] / R; kern["short"] += st["ms_short"] / R
        kern["band_phase_wall"] += st["ms_band"] / R; kern["long_wall"] += st["ms_long_wall"] / R
    m = float(np.median(ms))
    st = eng.stats()
    peak, _ = measured_peak()
    res = {"workload": WORKLOADS["wgs_sv"]["config"], "n_superclusters": n, "cells": cells,
           "e2e_ms_per_step": m, "e2e_gcells_per_s": cells / (m * 1e-3) / 1e9,
           "e2e_superclusters_per_s": n / (m * 1e-3), "device_ms_per_step": float(np.median(dev_ms)),
           "kernel_ms_per_step": kern,
           "kernel_ms_note": "long_fwd/bwd/walk are summed over rungs and shape classes, which run concurrently on their own streams; "
                             "band_phase_wall / long_wall are wall times of the banded phase and of the whole long path",
           "long_alignments": int(st["n_long"]), "long_alignments_left_to_dense_kernels": int(st["n_dense"]),
           "tie_superclusters": tie_count(out.status, b.n_sc)}
    # the banded forward sweep against the HBM roof, on the bytes it actually visits: one flag byte per kept cell
    # written, the 16-byte row records and the band records read
    if st.get("band_cells", 0) and kern["long_fwd"] > 0:
        vis = float(st["band_cells"]) + 16.0 * float(st.get("band_rows", 0)) + 16.0 * float(st.get("band_cols", 0))
        res["roofline_band_fwd"] = {"bound": "hbm", "achieved": vis / (kern["long_fwd"] * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                    "frac": vis / (kern["long_fwd"] * 1e-3) / 1e9 / peak, "visited_cells": int(st["band_cells"]),
                                    "note": "visited bytes / summed forward-kernel durations of all rungs; the sweep is latency-bound "
                                            "(one warp per alignment, a dependent column step), not bandwidth-bound"}
    if not args.no_cpu_baseline and checkers.reference_available(False):
        cores = os.cpu_count() or 1
        sb, scells, _ = make_workload("wgs_sv", args.sv_cpu_sample, args.seed, 0, 1, args.sv_max)
        _, sec = checkers.reference_run(sb, canonical=False, threads=cores)
        res["cpu_baseline"] = {"value": scells / sec / 1e9, "unit": "Gcells/s", "cores": cores, "kind": "reference",
                               "superclusters_per_s": sb.n_sc / sec,
                               "sample": f"{sb.n_sc} superclusters ({scells} cells) of the same workload, one pass, -t {cores}"}
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="wgs", choices=sorted(WORKLOADS))
    ap.add_argument("--n-sc", type=int, default=0, help="superclusters per GPU (default: the workload's)")
    ap.add_argument("--sv-max", type=int, default=10000)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--ref-sample", type=int, default=0,
                    help="--impl reference: superclusters per step (0 = the workload's full batch, i.e. the GPU arm's config)")
    ap.add_argument("--cpu-sample", type=int, default=600_000, help="superclusters in the cpu_baseline sample of the GPU arm's line")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--exchange", default="overlap", choices=["overlap", "sync"],
                    help="N > 1: the all-gather of a step's records runs beside the next step's kernels (two records alternate), "
                         "or every step waits for its own exchange")
    ap.add_argument("--sv-n-sc", type=int, default=3_600_000, help="superclusters of the secondary (configs[3]) workload")
    ap.add_argument("--sv-cpu-sample", type=int, default=40_000, help="superclusters in the secondary workload's cpu_baseline sample")
    ap.add_argument("--secondary", action="store_true", default=True,
                    help="also time the SV-bearing workload (BASELINE configs[3]) at full scale through vd_run (default at N=1)")
    ap.add_argument("--no-secondary", dest="secondary", action="store_false")
    ap.add_argument("--no-seeds", action="store_true", help="skip the four extra draws of the workload (seeds 2..5)")
    ap.add_argument("--no-seam", action="store_true", help="skip the timing at the reference's own seam (function harness + CLI)")
    ap.add_argument("--cli-contig-len", type=int, default=6_000_000, help="seam timing through the CLI: bases per synthetic contig (0 = skip)")
    ap.add_argument("--cli-cluster-contig-len", type=int, default=2_000_000,
                    help="seam timing through the CLI with the default (biwfa) clustering: bases per synthetic contig (0 = skip)")
    args = ap.parse_args()
    if args.impl == "reference":
        args.warmup = max(args.warmup, 1)
        return run_reference(args)
    if args.warmup < 3:
        args.warmup = 3

    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    n_sc = args.n_sc or WORKLOADS[args.workload]["n_sc"]
    b, cells_total, sc_total = make_workload(args.workload, n_sc, args.seed, rank, world, args.sv_max)
    n_var = b.n_var
    eng = capi.Engine(local_rank)
    stream = torch.cuda.ExternalStream(eng.stream, device=dev)

    # ---- device-resident copies of the batch and result buffers (torch owns the HBM) ----
    def dev_t(a):
        return torch.from_numpy(a).to(dev)
    d_in = {k: dev_t(getattr(b, k)) for k in ("ref_off", "ref_seq", "var_off", "var_pos", "var_rlen", "var_type",
                                              "alt_off", "alt_seq", "var_qual")}
    din = vd_batch_in()
    din.n_sc = b.n_sc
    for k, t in d_in.items():
        setattr(din, k, t.data_ptr())
    din.rplane_seq = None
    din.max_qual = b.max_qual
    # The kernels write wide (32-bit) result arrays in HBM.  With several GPUs each rank then narrows its results into
    # a 16-bit exchange record (vd_pack_device, shard.PackedRecord) and ONE all-gather per step moves the records of
    # all ranks (north_star: "a single NCCL all-gather over NVLink of the per-cluster score/credit arrays at the end").
    # The batch-global indices of a shard never change and are exchanged once, before the timed region.  Two records
    # alternate, so that the all-gather of step k (comm stream) runs beside the kernels of step k+1; every step's
    # exchange completes inside the timed region (--exchange sync: each step waits for its own all-gather).
    wide = {"aln_score": torch.zeros(4 * b.n_sc, dtype=torch.int32, device=dev), "status": torch.zeros(4 * b.n_sc, dtype=torch.int32, device=dev),
            "aln_end_plane": torch.zeros(4 * b.n_sc, dtype=torch.uint8, device=dev), "aln_beg_plane": torch.zeros(4 * b.n_sc, dtype=torch.uint8, device=dev),
            "assigned": torch.zeros(max(2 * n_var, 1), dtype=torch.uint8, device=dev), "callq": torch.zeros(max(2 * n_var, 1), dtype=torch.float32, device=dev)}
    for k in ("sync_group", "ref_ed", "query_ed"):
        wide[k] = torch.zeros(max(2 * n_var, 1), dtype=torch.int32, device=dev)
    dout = vd_batch_out()
    for name, t in wide.items():
        setattr(dout, name, t.data_ptr())
    K = 1
    precs, dpacked, gathered, comm_done = [], [], [], []
    comm = None
    if world > 1:
        counts = torch.tensor([b.n_sc, n_var], dtype=torch.int64, device=dev)
        allc = [torch.zeros_like(counts) for _ in range(world)]
        dist.all_gather(allc, counts)
        allc = torch.stack(allc).cpu().numpy()
        cap_sc, cap_var = int(allc[:, 0].max()), max(int(allc[:, 1].max()), 1)
        off_sc, off_var = int(allc[:rank, 0].sum()), int(allc[:rank, 1].sum())
        # ranks own disjoint ranges of the global arrays; exchanged once
        all_sc, all_var = shard.gather_index(np.arange(off_sc, off_sc + b.n_sc), np.arange(off_var, off_var + n_var), cap_sc, cap_var, dist, dev)
        assert sum(len(x) for x in all_sc) == sc_total
        for _ in range(2):
            pr = shard.PackedRecord(cap_sc, cap_var, dev)
            pr.set_counts(b.n_sc, n_var)
            dp = vd_packed_out()
            for name in ("aln_score", "aln_planes", "status", "sync_group", "ref_ed", "query_ed", "callq"):
                setattr(dp, name, pr.views[name].data_ptr())
            precs.append(pr); dpacked.append(dp)
            gathered.append(torch.empty(world * pr.nbytes, dtype=torch.uint8, device=dev))
            comm_done.append(torch.cuda.Event())
        comm = torch.cuda.Stream(device=dev)
    torch.cuda.synchronize()
    step_stats = {}
    step_no = [0]

    def run_pass(engine, acc, exchange):
        """One pass of the hot path over this rank's shard (+ the exchange of its results)."""
        acc.clear()
        engine.run_device(din, dout, n_var, b.ref_bytes, b.alt_bytes)      # returns with the kernels finished
        acc.update(engine.stats())
        if exchange and world > 1:
            i = step_no[0] % 2
            step_no[0] += 1
            stream.wait_event(comm_done[i])        # the all-gather that last read this record (two steps ago) is done
            engine.pack_device(dout, b.n_sc, n_var, dpacked[i])
            ready = torch.cuda.Event()
            ready.record(stream)
            with torch.cuda.stream(comm):
                comm.wait_event(ready)
                precs[i].all_gather(dist, gathered[i])
                comm_done[i].record(comm)
            if args.exchange == "sync":
                stream.wait_event(comm_done[i])

    def step_resident():
        run_pass(eng, step_stats, True)

    def finish_exchange():
        if world > 1:
            stream.wait_stream(comm)               # the timed region ends when the last step's records have arrived

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up, then EXACTLY K timed steps; device time, max over ranks ----
    for _ in range(args.warmup):
        step_resident()
    finish_exchange()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    ms_short = ms_fwd = ms_bwd = ms_walk = ms_plan = ms_kernels = 0.0
    ms_small = [0.0, 0.0, 0.0]
    barrier()
    ev0.record(stream)
    for _ in range(args.steps):
        step_resident()
        st = dict(step_stats)
        launches += st["n_launches"] + (1 if world > 1 else 0)
        ms_short += st["ms_short"]; ms_fwd += st["ms_long_fwd"]; ms_bwd += st["ms_long_bwd"]
        ms_walk += st["ms_long_walk"]; ms_plan += st["ms_plan"]; ms_kernels += st["ms_total"]
        ms_small = [a + b_ for a, b_ in zip(ms_small, st["ms_small"])]
    finish_exchange()
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop()
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    st = dict(step_stats)
    per_rank = None
    if world > 1:        # cell load and kernel time of every rank (LPT by cells, shard.lpt_partition) and the straggler's long path
        mine = torch.tensor([float(b.cells().sum()), float(b.n_sc), ms / args.steps, ms_kernels / args.steps,
                             float(st["ms_long_wall"]), float(st["n_long"])], dtype=torch.float64, device=dev)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        allr = torch.stack(allr).cpu().numpy()
        per_rank = {"cells": [int(x) for x in allr[:, 0]], "superclusters": [int(x) for x in allr[:, 1]],
                    "ms_per_step": [round(float(x), 3) for x in allr[:, 2]], "kernel_ms_per_step": [round(float(x), 3) for x in allr[:, 3]],
                    "long_path_wall_ms": [round(float(x), 3) for x in allr[:, 4]], "long_alignments": [int(x) for x in allr[:, 5]],
                    "cell_imbalance_max_over_mean": float(allr[:, 0].max() / allr[:, 0].mean())}

    # ---- e2e: vd_run() with pinned host buffers, H2D + D2H inside the timed region ----
    def pin(a):
        t_ = torch.from_numpy(a).pin_memory()
        return t_, t_.numpy()
    keep = []
    hb = {}
    for k in ("ref_off", "ref_seq", "var_off", "var_pos", "var_rlen", "var_type", "alt_off", "alt_seq", "var_qual"):
        t_, a_ = pin(getattr(b, k)); keep.append(t_); hb[k] = a_
    bp = Batch(**hb, max_qual=b.max_qual)
    # results as 16-bit records (vd_run_packed): the copy out over PCIe bounds the end-to-end step, and the host
    # float step (vd_finalize_packed) reads them as they are; the wide records (vd_run) are timed beside it
    ho = Out(b.n_sc, n_var)
    for f in Out.FIELDS:
        t_, a_ = pin(getattr(ho, f)); keep.append(t_); setattr(ho, f, a_)
    hp = PackedOut(b.n_sc, n_var)
    for f in PackedOut.FIELDS:
        t_, a_ = pin(getattr(hp, f)); keep.append(t_); setattr(hp, f, a_)
    e2e_steps = max(3, min(args.steps, 7))

    def time_e2e(call):
        for _ in range(3):
            call()
        barrier()
        ts = []
        for _ in range(e2e_steps):             # synchronous calls: results are in host memory on return
            t0 = time.perf_counter()
            call()
            ts.append((time.perf_counter() - t0) * 1e3)
        torch.cuda.synchronize()
        return ts, eng.stats()
    # the batch in compact form (vd_compact_pack: lengths instead of 64-bit offsets, 16-bit positions), page-locked
    ci = capi.compact(bp)
    for name, _, _ in ci.OWN:
        t_, a_ = pin(getattr(ci, name)); keep.append(t_); setattr(ci, name, a_)
    ci.refresh_pointers()
    wide_times, st_w = time_e2e(lambda: eng.run(bp, ho))
    packed_times, st_p = time_e2e(lambda: eng.run_packed(bp, hp))
    e2e_times, st_e = time_e2e(lambda: eng.run_compact(ci, hp))
    e2e_ms = float(np.mean(e2e_times))
    te = torch.tensor([e2e_ms, float(np.mean(wide_times)), float(np.mean(packed_times))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_ms, e2e_wide_ms, e2e_packed_ms = float(te[0].item()), float(te[1].item()), float(te[2].item())

    if rank == 0:
        peak, peak_src = measured_peak()
        ms_step = ms_max / args.steps
        value = cells_total / (ms_step * 1e-3) / 1e9
        # Dominant kernel and its roofline.  In the timed region the launch groups of
        # the short kernels run on concurrent streams, so their event times overlap; the per-kernel
        # durations used here come from a serial pass (VD_SERIAL=1: same kernels, same inputs, one stream),
        # which is also what the committed ncu launch list shows.
        #   short regime -> small_kernel<k> / wsc_kernel<S> move the compact batch of their superclusters in
        #                   and the result records out (vd_stats.io_small[k]);
        #   long regime  -> wave_fwd writes 1 B/cell of flags
        os.environ["VD_SERIAL"] = "1"
        try:
            eng_s = capi.Engine(local_rank)
        finally:
            del os.environ["VD_SERIAL"]
        ser = {"small_kernel<0>": [0.0, 0.0], "small_kernel<1>": [0.0, 0.0], "wsc_kernel<S>": [0.0, 0.0], "wave_fwd_kernel": [0.0, 0.0]}
        SER_STEPS = 3
        for i in range(2 + SER_STEPS):
            ss = {}
            run_pass(eng_s, ss, False)
            if i >= 2:
                for k_, nm in enumerate(("small_kernel<0>", "small_kernel<1>", "wsc_kernel<S>")):
                    ser[nm][0] += ss["ms_small"][k_] / SER_STEPS; ser[nm][1] += float(ss["io_small"][k_]) / SER_STEPS
                ser["wave_fwd_kernel"][0] += ss["ms_long_fwd"] / SER_STEPS; ser["wave_fwd_kernel"][1] += ss["spill_bytes"] / 3.0 / SER_STEPS
                ser_total = ss["ms_total"]
        eng_s.close()
        k_short, k_fwd, k_bwd = ms_short / args.steps, ms_fwd / args.steps, ms_bwd / args.steps
        k_small = [x / args.steps for x in ms_small]
        dom = max(ser, key=lambda k_: ser[k_][0])
        dom_ms, alg_bytes = ser[dom]
        achieved = alg_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
        line = {
            "metric": "dp_gcells_per_s", "value": value, "unit": "Gcells/s",
            "superclusters_per_s": sc_total / (ms_step * 1e-3),
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32", "data": "synthetic",
            "config": {"workload": WORKLOADS[args.workload]["config"],
                       "source": "superclusters resampled (seeded) from the real HG002 chr1:1-5Mb demo batch"
                                 + (" + synthetic SV tail" if WORKLOADS[args.workload]["sv_frac"] else ""),
                       "n_superclusters_per_step": sc_total, "cells_per_step": cells_total,
                       "per_gpu_superclusters": b.n_sc, "parallelism": f"shard{world} (LPT by cells) + one all-gather of 16-bit records per step ({args.exchange})" if world > 1 else "single",
                       "l2": "inputs+outputs exceed L2 (no flush needed)" if b.io_bytes() > 200e6 else "small batch: L2-resident"},
            "clocks": clocks,
            "e2e": {"value": cells_total / (e2e_ms * 1e-3) / 1e9, "unit": "Gcells/s",
                    "superclusters_per_s": sc_total / (e2e_ms * 1e-3), "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": int(st_e["h2d_bytes"]), "d2h_bytes_per_step": int(st_e["d2h_bytes"]),
                    "api": "vd_run_compact (C-ABI, pinned host buffers: compact batch in, 16-bit result records out)", "steps": e2e_steps,
                    "ms_min": float(min(e2e_times)), "ms_max": float(max(e2e_times)),
                    "other_entry_points": {"vd_run_packed (vd_batch_in in, 16-bit records out)":
                                           {"ms_per_step": e2e_packed_ms, "h2d_bytes_per_step": int(st_p["h2d_bytes"]), "d2h_bytes_per_step": int(st_p["d2h_bytes"])},
                                           "vd_run (vd_batch_in in, 32-bit records out)":
                                           {"ms_per_step": e2e_wide_ms, "h2d_bytes_per_step": int(st_w["h2d_bytes"]), "d2h_bytes_per_step": int(st_w["d2h_bytes"])}}},
            "gpu_launches": int(launches),
            "per_rank": per_rank,
            "exchange": ({"collective": "all_gather_into_tensor (NCCL) of shard.PackedRecord", "bytes_per_rank": precs[0].nbytes,
                          "received_bytes_per_rank_per_step": (world - 1) * precs[0].nbytes, "mode": args.exchange,
                          "static_index_exchanged_once_bytes_per_rank": 8 * (2 + cap_sc + cap_var)} if world > 1 else None),
            "tie_superclusters": {"count": tie_count(ho.status, b.n_sc) if world == 1 else None, "of": b.n_sc,
                                  "note": "superclusters with VD_ST_TIE on this rank: an ambiguous swap edge on an optimal path, where the "
                                          "reference's pick depends on unordered_set iteration order (SURVEY.md 8a); parity there is against oracle-B"},
            "kernel_ms_per_step": {"plan": ms_plan / args.steps, "short_region": k_short, "small_kernel<0>": k_small[0],
                                   "small_kernel<1>": k_small[1], "wsc_kernel<S>(sum, overlapping)": k_small[2],
                                   "wave_fwd": k_fwd, "wave_bwd": k_bwd,
                                   "wave_walk": ms_walk / args.steps, "all_kernels": ms_kernels / args.steps},
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": dom_ms,
                         "timing": "CUDA events around the kernel in a serial pass of the same step (VD_SERIAL=1); "
                                   "wsc_kernel<S> = the warp path of the mid-size superclusters: wsc_expand_kernel + all wsc_sweep_* launches "
                                   "(one per register slots x shared-memory bin x homozygous) + wsc_walk_kernel",
                         "serial_pass_ms": {k_: v_[0] for k_, v_ in ser.items()}, "serial_step_ms": ser_total,
                         "superclusters_per_class": [int(x) for x in st["n_small"]] + [int(st["n_long"]) // 4]},
        }
        # measured DRAM traffic of the dominant kernel (ncu --set full capture of this round, per launch
        # at this workload size; profiles/README.md), null when no capture matches the workload
        tr = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tr):
            t_ = json.load(open(tr)).get(f"{args.workload}:{dom}")
            if t_ and t_.get("n_sc_per_gpu") == b.n_sc:
                line["roofline"]["traffic"] = t_["dram_bytes_per_launch"]
                # the kernel is issue-bound, not HBM-bound: what the same capture says about the issue slots
                line["roofline"]["ncu"] = {k_: t_[k_] for k_ in ("issue_slots_busy_pct", "warps_active_pct",
                                                                  "active_lanes_per_instruction") if k_ in t_}
        if not args.no_seeds and world == 1:
            line["seeds"] = run_seeds(eng, args, dev, torch, {"ms_per_step": ms_kernels / args.steps, "gcells_per_s": cells_total / (ms_kernels / args.steps * 1e-3) / 1e9,
                                                               "cells": cells_total})
        if args.secondary and world == 1 and args.workload == "wgs":
            line["secondary"] = run_secondary(eng, args, dev, torch)
        if not args.no_seam and world == 1 and args.workload == "wgs":
            line["seam"] = run_seam(args, b)
        if not args.no_cpu_baseline and world == 1:
            cores = os.cpu_count() or 1
            sb, scells, _ = make_workload(args.workload, args.cpu_sample, args.seed, 0, 1, args.sv_max)
            if checkers.reference_available(False):
                secs = [checkers.reference_run(sb, canonical=False, threads=cores)[1] for _ in range(4)]
                sec = float(np.median(secs[1:]))          # one warm-up pass, median of three
                kind = "reference"
            else:
                t0 = time.perf_counter(); checkers.oracle_run(sb); sec = time.perf_counter() - t0
                kind, cores = "port", 1
            line["cpu_baseline"] = {"value": scells / sec / 1e9, "unit": "Gcells/s", "cores": cores, "kind": kind,
                                    "superclusters_per_s": sb.n_sc / sec,
                                    "sample": f"{sb.n_sc} superclusters ({scells} cells) of the same workload, one warm-up pass then "
                                              f"the median of three, reference std::thread ladder with -t {cores}"}
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
