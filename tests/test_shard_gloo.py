"""N>1 path on CPU: LPT partition by estimated cells, per-rank compute, one all-gather of the
result records (gloo, world_size 2).  The oracle stands in for the GPU here (tests only)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import OUT_KEYS, ROOT, mismatches
from vcfdist_b200 import capi, shard
from workloads import synth
from oracle import checkers


def test_lpt_partition_balances_and_covers():
    rng = np.random.default_rng(0)
    cells = np.concatenate([rng.integers(50, 400, 20000), rng.integers(10**6, 10**8, 40)])
    parts = shard.lpt_partition(cells, 8)
    allidx = np.sort(np.concatenate(parts))
    assert (allidx == np.arange(len(cells))).all()
    loads = np.array([cells[p].sum() for p in parts], np.float64)
    assert loads.max() / loads.mean() < 1.25
    # deterministic
    again = shard.lpt_partition(cells, 8)
    assert all((a == b).all() for a, b in zip(parts, again))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b = synth.wgs_like(3, 600, sv_frac=0.01, sv_max=300)
    parts = shard.lpt_partition(b.cells(), world)
    mine = parts[rank]
    sub = b.take(mine)
    out = checkers.oracle_run(sub).trimmed()
    full = shard.gather_results(out, mine, b.var_index_of(mine), b.n_sc, b.n_var, dist)
    want = checkers.oracle_run(b).trimmed()
    got = {k: v.numpy() for k, v in full.items()}
    got["status"] = got["status"].view(np.uint32)
    bad = mismatches(got, want, OUT_KEYS)
    # the bench's path: results written into one contiguous record, ONE all-gather, no scatter
    vidx = b.var_index_of(mine)
    rec = shard.ResultRecord(max(len(p) for p in parts), b.n_var, "cpu")
    rec.set_shard(mine, vidx)
    for k in ("aln_score", "aln_end_plane", "aln_beg_plane", "assigned", "sync_group", "ref_ed", "query_ed", "callq"):
        rec.views[k][: len(out[k])] = torch.from_numpy(out[k])
    rec.views["status"][: len(out["status"])] = torch.from_numpy(out["status"].view(np.int32))
    g = rec.all_gather(dist)
    for r in range(world):
        pr = rec.parse(g, r)
        sidx = pr["sc_idx"].numpy().astype(np.int64)
        assert (np.sort(sidx) == parts[r]).all()
        aidx = (sidx[:, None] * 4 + np.arange(4)[None, :]).reshape(-1)
        if not (pr["aln_score"].numpy() == want["aln_score"][aidx]).all():
            bad["record_score"] = 1
        vi = pr["var_idx"].numpy().astype(np.int64)
        for k in ("assigned", "sync_group", "ref_ed", "query_ed"):
            if not (pr[k].numpy() == want[k].reshape(2, -1)[:, vi]).all():
                bad["record_" + k] = 1
    # the bench's path beyond two GPUs: the shard in K slices, one record and one all-gather per slice
    K = 3
    cut = [len(mine) * k // K for k in range(K + 1)]
    counts = torch.tensor([[cut[k + 1] - cut[k], int(np.diff(sub.var_off[[4 * cut[k], 4 * cut[k + 1]]])[0])] for k in range(K)])
    allc = [torch.zeros_like(counts) for _ in range(world)]
    dist.all_gather(allc, counts)
    caps = torch.stack(allc).numpy().max(axis=0)
    seen_sc = []
    for k in range(K):
        sl = np.arange(cut[k], cut[k + 1])
        part = sub.take(sl)
        o = checkers.oracle_run(part).trimmed()
        rk = shard.ResultRecord(int(caps[k, 0]), max(int(caps[k, 1]), 1), "cpu")
        rk.set_shard(mine[sl], vidx[int(sub.var_off[4 * cut[k]]): int(sub.var_off[4 * cut[k + 1]])])
        for name in ("aln_score", "aln_end_plane", "aln_beg_plane", "assigned", "sync_group", "ref_ed", "query_ed", "callq"):
            rk.views[name][: len(o[name])] = torch.from_numpy(o[name])
        rk.views["status"][: len(o["status"])] = torch.from_numpy(o["status"].view(np.int32))
        gk = rk.all_gather(dist)
        for r in range(world):
            pr = rk.parse(gk, r)
            sidx = pr["sc_idx"].numpy().astype(np.int64)
            if r == rank:
                seen_sc.append(sidx)
            aidx = (sidx[:, None] * 4 + np.arange(4)[None, :]).reshape(-1)
            if not (pr["aln_score"].numpy() == want["aln_score"][aidx]).all():
                bad["slice_score"] = 1
            vi = pr["var_idx"].numpy().astype(np.int64)
            for name in ("assigned", "sync_group", "ref_ed", "query_ed"):
                if not (pr[name].numpy() == want[name].reshape(2, -1)[:, vi]).all():
                    bad["slice_" + name] = 1
    if not (np.concatenate(seen_sc) == mine).all():
        bad["slice_cover"] = 1
    # the 16-bit exchange record (what vd_pack_device fills on the GPU): indices exchanged once, then one all-gather
    cap_sc, cap_var = max(len(p) for p in parts), b.n_var
    all_sc, all_var = shard.gather_index(mine, vidx, cap_sc, cap_var, dist, "cpu")
    prec = shard.PackedRecord(cap_sc, cap_var, "cpu")
    prec.set_counts(len(mine), len(vidx))
    a, v = 4 * len(mine), 2 * len(vidx)
    sc16 = np.where(out["aln_score"] < 0, 0xFFFF, out["aln_score"]).astype(np.uint16)
    prec.views["aln_score"][:a] = torch.from_numpy(sc16.view(np.int16))
    prec.views["status"][:a] = torch.from_numpy(out["status"].astype(np.uint16).view(np.int16))
    prec.views["aln_planes"][:a] = torch.from_numpy((out["aln_end_plane"] | (out["aln_beg_plane"] << 1)).astype(np.uint8))
    prec.views["sync_group"][:v] = torch.from_numpy(((out["assigned"].astype(np.uint16) << 14) | out["sync_group"].astype(np.uint16)).view(np.int16))
    prec.views["ref_ed"][:v] = torch.from_numpy(out["ref_ed"].astype(np.uint16).view(np.int16))
    prec.views["query_ed"][:v] = torch.from_numpy(out["query_ed"].astype(np.uint16).view(np.int16))
    prec.views["callq"][:v] = torch.from_numpy(out["callq"])
    gp = prec.all_gather(dist)
    for r in range(world):
        if not (all_sc[r] == parts[r]).all():
            bad["packed_index"] = 1
        pr = prec.parse(gp, r)
        aidx = (all_sc[r][:, None] * 4 + np.arange(4)[None, :]).reshape(-1)
        for name in ("aln_score", "aln_end_plane", "aln_beg_plane"):
            if not (pr[name].numpy() == want[name][aidx]).all():
                bad["packed_" + name] = 1
        if not (pr["status"].numpy().astype(np.uint32) == want["status"][aidx]).all():
            bad["packed_status"] = 1
        for name in ("assigned", "sync_group", "ref_ed", "query_ed", "callq"):
            if not (pr[name].numpy() == want[name].reshape(2, -1)[:, all_var[r]]).all():
                bad["packed_" + name] = 1
    q.put((rank, bad))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_and_gather_matches_single_pass():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, bad in res:
        assert bad == {}, (rank, bad)
