"""Multi-GPU plumbing: superclusters are independent units (src/dist.cpp:1738-1903 touches only
sc_idx-local state), so a batch is partitioned across ranks by estimated DP-cell count with no
data-path collective, and the fixed-width result records are exchanged once at the end with a
single all-gather (NCCL over NVLink on the GPU box, gloo in the CPU tests).

The reference's analogue of the partition is the RAM/thread ladder (src/cluster.cpp:94-117,
src/dist.cpp:1670-1721)."""
from __future__ import annotations

from typing import List, Sequence

import numpy as np


def lpt_partition(cells: np.ndarray, n_parts: int, exact_top: int = 4096) -> List[np.ndarray]:
    """Longest-processing-time partition of item weights `cells` into `n_parts` index lists.
    The `exact_top` heaviest items are placed greedily on the least-loaded part (classic LPT);
    the long tail of small items is dealt round-robin in descending order, which is LPT up to
    one item per round.  Deterministic; every rank computes the same answer."""
    cells = np.asarray(cells, np.int64)
    n = len(cells)
    order = np.argsort(-cells, kind="stable")
    parts: List[List[np.ndarray]] = [[] for _ in range(n_parts)]
    load = np.zeros(n_parts, np.int64)
    top = order[: min(exact_top, n)]
    assign = np.empty(len(top), np.int64)
    for k, i in enumerate(top):
        p = int(np.argmin(load))
        assign[k] = p
        load[p] += cells[i]
    for p in range(n_parts):
        parts[p].append(top[assign == p])
    rest = order[len(top):]
    if len(rest):
        # continue on the currently least-loaded parts first
        start = np.argsort(load, kind="stable")
        for j, p in enumerate(start):
            parts[int(p)].append(rest[j::n_parts])
    return [np.sort(np.concatenate(p)) if p else np.zeros(0, np.int64) for p in parts]


def gather_results(local: dict, sc_idx: np.ndarray, var_idx: np.ndarray, n_sc: int, n_var: int,
                   dist, device=None) -> dict:
    """All-gather the per-alignment and per-variant result arrays of every rank's shard and
    scatter them back to batch order.  `local` holds this rank's vd_batch_out arrays (torch
    tensors on `device`, or numpy for the CPU tests); `sc_idx` / `var_idx` are the batch-global
    indices of the shard's superclusters / variants.  One collective per dtype width."""
    import torch
    world = dist.get_world_size()
    dev = device if device is not None else "cpu"

    def as_i32(x):
        if isinstance(x, torch.Tensor):
            t = x.to(dev)
            return t if t.dtype == torch.int32 else t.view(torch.int32)
        return torch.from_numpy(np.ascontiguousarray(x).view(np.int32)).to(dev)

    def as_u8(x):
        if isinstance(x, torch.Tensor):
            return x.to(dev)
        return torch.from_numpy(np.ascontiguousarray(x, np.uint8)).to(dev)

    n_loc_sc, n_loc_var = len(sc_idx), len(var_idx)
    counts = torch.tensor([n_loc_sc, n_loc_var], dtype=torch.int64, device=dev)
    all_counts = [torch.zeros(2, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(all_counts, counts)
    all_counts = torch.stack(all_counts).cpu().numpy()
    max_sc, max_var = int(all_counts[:, 0].max()), int(all_counts[:, 1].max())

    # pack everything of this rank into one int32 record buffer: [indices | per-alignment | per-variant]
    def pad(t, n):
        out = torch.zeros(n, dtype=t.dtype, device=dev)
        out[: t.numel()] = t
        return out

    i32 = torch.int32
    pieces = [
        pad(as_i32(sc_idx.astype(np.int32)), max_sc),
        pad(as_i32(var_idx.astype(np.int32)), max_var),
        pad(as_i32(local["aln_score"])[: 4 * n_loc_sc], 4 * max_sc),
        pad(as_i32(local["status"])[: 4 * n_loc_sc], 4 * max_sc),
    ]
    for k in ("sync_group", "ref_ed", "query_ed", "callq"):
        t = as_i32(local[k])
        for slot in range(2):
            pieces.append(pad(t[slot * n_loc_var: (slot + 1) * n_loc_var], max_var))
    u8 = torch.uint8
    pieces8 = [pad(as_u8(local["aln_end_plane"])[: 4 * n_loc_sc], 4 * max_sc),
               pad(as_u8(local["aln_beg_plane"])[: 4 * n_loc_sc], 4 * max_sc)]
    a8 = as_u8(local["assigned"])
    for slot in range(2):
        pieces8.append(pad(a8[slot * n_loc_var: (slot + 1) * n_loc_var], max_var))
    buf32 = torch.cat(pieces)
    buf8 = torch.cat(pieces8)
    # one record buffer per rank -> a single all-gather
    pad_n = (-(4 * buf32.numel() + buf8.numel())) % 16         # keep every rank's record 16-byte aligned
    rec = torch.cat([buf32.view(u8), buf8, torch.zeros(pad_n, dtype=u8, device=dev)])
    gathered = torch.empty(world * rec.numel(), dtype=u8, device=dev)
    dist.all_gather_into_tensor(gathered, rec)
    gathered = gathered.view(world, -1)

    out = {
        "aln_score": torch.zeros(4 * n_sc, dtype=i32, device=dev),
        "status": torch.zeros(4 * n_sc, dtype=i32, device=dev),
        "aln_end_plane": torch.zeros(4 * n_sc, dtype=u8, device=dev),
        "aln_beg_plane": torch.zeros(4 * n_sc, dtype=u8, device=dev),
        "assigned": torch.zeros(2 * n_var, dtype=u8, device=dev),
    }
    for k in ("sync_group", "ref_ed", "query_ed", "callq"):
        out[k] = torch.zeros(2 * n_var, dtype=i32, device=dev)
    n32 = buf32.numel()
    four = torch.arange(4, device=dev)
    for r in range(world):
        c_sc, c_var = int(all_counts[r, 0]), int(all_counts[r, 1])
        b32 = gathered[r, : 4 * n32].view(i32)
        b8 = gathered[r, 4 * n32: 4 * n32 + buf8.numel()]
        o = 0
        sidx = b32[o: o + c_sc].long(); o += max_sc
        vidx = b32[o: o + c_var].long(); o += max_var
        aidx = (sidx[:, None] * 4 + four[None, :]).reshape(-1)
        out["aln_score"][aidx] = b32[o: o + 4 * c_sc]; o += 4 * max_sc
        out["status"][aidx] = b32[o: o + 4 * c_sc]; o += 4 * max_sc
        for k in ("sync_group", "ref_ed", "query_ed", "callq"):
            for slot in range(2):
                out[k][slot * n_var + vidx] = b32[o: o + c_var]; o += max_var
        o = 0
        out["aln_end_plane"][aidx] = b8[o: o + 4 * c_sc]; o += 4 * max_sc
        out["aln_beg_plane"][aidx] = b8[o: o + 4 * c_sc]; o += 4 * max_sc
        for slot in range(2):
            out["assigned"][slot * n_var + vidx] = b8[o: o + c_var]; o += max_var
    out["callq"] = out["callq"].view(torch.float32)
    return out


class ResultRecord:
    """One contiguous, 16-byte-aligned device buffer holding a rank's whole result set, with the
    `vd_batch_out` arrays as views into it, so that vd_run_device writes straight into the record
    and the end-of-step exchange is ONE all-gather of `buf` with no packing or scatter:

        [ n_sc, n_var (int64) | sc_idx i32 | var_idx i32 | aln_score i32 | status i32 |
          sync_group i32 x2 | ref_ed i32 x2 | query_ed i32 x2 | callq f32 x2 |
          aln_end_plane u8 | aln_beg_plane u8 | assigned u8 x2 ]

    Every rank sizes the record for the largest shard (`cap_sc`, `cap_var`) so that the gathered
    tensor is a plain [world, rec_bytes] array; `parse()` returns per-rank views."""

    FIELDS32 = ("sc_idx", "var_idx", "aln_score", "status", "sync_group", "ref_ed", "query_ed", "callq")
    FIELDS8 = ("aln_end_plane", "aln_beg_plane", "assigned")

    def __init__(self, cap_sc: int, cap_var: int, device):
        import torch
        self.cap_sc, self.cap_var = cap_sc, cap_var
        n32 = {"sc_idx": cap_sc, "var_idx": cap_var, "aln_score": 4 * cap_sc, "status": 4 * cap_sc,
               "sync_group": 2 * cap_var, "ref_ed": 2 * cap_var, "query_ed": 2 * cap_var, "callq": 2 * cap_var}
        n8 = {"aln_end_plane": 4 * cap_sc, "aln_beg_plane": 4 * cap_sc, "assigned": 2 * cap_var}
        off, self.layout = 16, {}
        for k in self.FIELDS32:
            self.layout[k] = (off, n32[k], 4); off = (off + 4 * n32[k] + 15) & ~15
        for k in self.FIELDS8:
            self.layout[k] = (off, n8[k], 1); off = (off + n8[k] + 15) & ~15
        self.nbytes = off
        self.buf = torch.zeros(self.nbytes, dtype=torch.uint8, device=device)
        self.views = self._views(self.buf)

    def _views(self, buf):
        import torch
        v = {"counts": buf[:16].view(torch.int64)}
        for k, (o, n, w) in self.layout.items():
            t = buf[o: o + n * w]
            v[k] = t.view(torch.float32) if k == "callq" else (t.view(torch.int32) if w == 4 else t)
        return v

    def set_shard(self, sc_idx: np.ndarray, var_idx: np.ndarray):
        import torch
        self.views["counts"][0] = len(sc_idx)
        self.views["counts"][1] = len(var_idx)
        self.views["sc_idx"][: len(sc_idx)] = torch.from_numpy(np.asarray(sc_idx, np.int32)).to(self.buf.device)
        self.views["var_idx"][: len(var_idx)] = torch.from_numpy(np.asarray(var_idx, np.int32)).to(self.buf.device)

    def all_gather(self, dist, out=None):
        """The single collective of the path.  Returns the [world, nbytes] gathered tensor."""
        import torch
        world = dist.get_world_size()
        if out is None:
            out = torch.empty(world * self.nbytes, dtype=torch.uint8, device=self.buf.device)
        dist.all_gather_into_tensor(out, self.buf)
        return out.view(world, self.nbytes)

    def parse(self, gathered, rank: int) -> dict:
        """Views of rank `rank`'s arrays inside the gathered tensor, trimmed to its counts.
        Per-variant arrays come back as [2, n_var_of_rank] (slot-major with the shard's own stride)."""
        v = self._views(gathered[rank])
        n_sc, n_var = int(v["counts"][0]), int(v["counts"][1])
        out = {"sc_idx": v["sc_idx"][:n_sc], "var_idx": v["var_idx"][:n_var]}
        for k in ("aln_score", "status", "aln_end_plane", "aln_beg_plane"):
            out[k] = v[k][: 4 * n_sc]
        for k in ("sync_group", "ref_ed", "query_ed", "callq", "assigned"):
            out[k] = v[k][: 2 * n_var].view(2, n_var)
        return out


class PackedRecord:
    """The exchange record in 16-bit form (include/vcfdist_b200.h: vd_packed_out): what vd_pack_device writes behind a
    rank's kernels and ONE all-gather moves - half the bytes of ResultRecord.  The batch-global indices of the shard's
    superclusters and variants do not change from step to step and are exchanged once (`gather_index`), not with
    every record.

        [ n_sc, n_var (int64) | aln_score u16 | status u16 | sync_group u16 x2 (assigned << 14 | group) | ref_ed u16 x2 |
          query_ed u16 x2 | callq f32 x2 | aln_planes u8 ]"""

    FIELDS = (("aln_score", 2, "sc"), ("status", 2, "sc"), ("sync_group", 2, "var"), ("ref_ed", 2, "var"), ("query_ed", 2, "var"),
              ("callq", 4, "var"), ("aln_planes", 1, "sc"))

    def __init__(self, cap_sc: int, cap_var: int, device):
        import torch
        self.cap_sc, self.cap_var = cap_sc, cap_var
        off, self.layout = 16, {}
        for k, w, per in self.FIELDS:
            n = 4 * cap_sc if per == "sc" else 2 * cap_var
            self.layout[k] = (off, n, w); off = (off + w * n + 15) & ~15
        self.nbytes = off
        self.buf = torch.zeros(self.nbytes, dtype=torch.uint8, device=device)
        self.views = self._views(self.buf)

    def _views(self, buf):
        import torch
        v = {"counts": buf[:16].view(torch.int64)}
        for k, (o, n, w) in self.layout.items():
            t = buf[o: o + n * w]
            v[k] = t.view(torch.float32) if k == "callq" else (t.view(torch.int16) if w == 2 else t)
        return v

    def set_counts(self, n_sc: int, n_var: int):
        self.views["counts"][0] = n_sc
        self.views["counts"][1] = n_var

    def all_gather(self, dist, out=None):
        import torch
        world = dist.get_world_size()
        if out is None:
            out = torch.empty(world * self.nbytes, dtype=torch.uint8, device=self.buf.device)
        dist.all_gather_into_tensor(out, self.buf)
        return out.view(world, self.nbytes)

    def parse(self, gathered, rank: int) -> dict:
        """Rank `rank`'s arrays inside the gathered tensor, widened to the field meaning of vd_batch_out
        (per-variant arrays as [2, n_var_of_rank])."""
        import torch
        v = self._views(gathered[rank])
        n_sc, n_var = int(v["counts"][0]), int(v["counts"][1])
        u16 = lambda t: t.to(torch.int32) & 0xFFFF
        sc = u16(v["aln_score"][: 4 * n_sc])
        sg = u16(v["sync_group"][: 2 * n_var]).view(2, n_var)
        pl = v["aln_planes"][: 4 * n_sc]
        return {"aln_score": torch.where(sc == 0xFFFF, torch.full_like(sc, -1), sc), "status": u16(v["status"][: 4 * n_sc]),
                "aln_end_plane": pl & 1, "aln_beg_plane": (pl >> 1) & 1, "assigned": (sg >> 14).to(torch.uint8),
                "sync_group": sg & 0x3FFF, "ref_ed": u16(v["ref_ed"][: 2 * n_var]).view(2, n_var),
                "query_ed": u16(v["query_ed"][: 2 * n_var]).view(2, n_var), "callq": v["callq"][: 2 * n_var].view(2, n_var)}


def gather_index(sc_idx: np.ndarray, var_idx: np.ndarray, cap_sc: int, cap_var: int, dist, device):
    """One-off exchange of every rank's batch-global supercluster / variant indices -> ([world][n_sc_r], [world][n_var_r])
    as int64 numpy arrays.  The partition is static, so this is not part of a step."""
    import torch
    world = dist.get_world_size()
    mine = torch.full((2 + cap_sc + cap_var,), -1, dtype=torch.int64, device=device)
    mine[0], mine[1] = len(sc_idx), len(var_idx)
    mine[2: 2 + len(sc_idx)] = torch.from_numpy(np.asarray(sc_idx, np.int64)).to(device)
    mine[2 + cap_sc: 2 + cap_sc + len(var_idx)] = torch.from_numpy(np.asarray(var_idx, np.int64)).to(device)
    allv = torch.empty(world * mine.numel(), dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(allv, mine)
    allv = allv.view(world, -1).cpu().numpy()
    return ([allv[r, 2: 2 + allv[r, 0]] for r in range(world)],
            [allv[r, 2 + cap_sc: 2 + cap_sc + allv[r, 1]] for r in range(world)])
