"""Cluster growing on the GPU: the reference's wf_swg_cluster (src/cluster.cpp:954-1263) with
its two kernels - wf_swg_max_reach (src/dist.cpp:2150-2333) and the score of wf_swg_align (:1510-1652) - run as
BATCHES of the hand-written wavefront kernel (vd_wf_batch, csrc/vd_reach.cuh): one launch per round over every
active cluster of a haplotype instead of one call per cluster and direction on one host thread (the reference's
wall-clock bottleneck on SV input).

Every cluster starts as one variant.  Per iteration (:979-1160):
  score   of every active cluster against its reference window                    -> one batch   (:1037-1046)
  reaches leftwards and rightwards with iterative window doubling (:1049-1158)    -> one batch per doubling round
          over the (cluster, direction) searches that still hit the far end of their window
then clusters whose reaches come within reach_min_gap are merged, rightwards then leftwards (:1171-1238).
The merge passes and the doubling control stay on the host: they are O(#clusters) integer work.

var[i] = (pos, rlen, type, alt bytes) sorted by position, one haplotype of one contig.
Product code: needs the CUDA library and a GPU (capi.Engine); there is no CPU fallback.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

from .capi import Engine

TYPE_SUB, TYPE_INS, TYPE_DEL = 1, 2, 3
INT_MAX = 2 ** 31 - 1
WF_REACH, WF_SCORE = 0, 1


def apply_variants(fasta: bytes, var: Sequence[tuple], n_all: int, beg_idx: int, end_idx: int, beg_pos: int, end_pos: int) -> bytes:
    """generate_str (src/dist.cpp:81-136, min_qual 0): the window [beg_pos, end_pos) of the contig with variants
    beg_idx .. end_idx-1 applied."""
    out = bytearray()
    vi = beg_idx
    while vi < n_all and var[vi][0] < beg_pos:
        vi += 1
    ref_pos = beg_pos
    while ref_pos < end_pos:
        if vi < end_idx and ref_pos == var[vi][0]:
            _pos, rlen, ty, alt = var[vi]
            if ty == TYPE_INS:
                out += alt
            elif ty == TYPE_DEL:
                ref_pos += rlen
            elif ty == TYPE_SUB:
                out += alt
                ref_pos += 1
            else:                                    # complex: ALT replaces REF
                out += alt
                ref_pos += rlen
            vi += 1
        else:
            ref_end = min(end_pos, var[vi][0]) if vi < end_idx else end_pos
            if ref_end < ref_pos:
                raise ValueError("variants overlap or are unsorted")
            out += fasta[ref_pos:ref_end]
            ref_pos = ref_end
    return bytes(out)


class _Search:
    """One reach search of one cluster (left or right) across its doubling rounds."""
    __slots__ = ("c", "left", "first", "last", "beg_pos", "end_pos", "ref_len", "main_diag", "main_diag_start", "score", "reach", "done")


def wf_swg_cluster(eng: Engine, fasta: bytes, var: Sequence[tuple], sub: int, open_: int, extend: int,
                   max_iters: int = 4, reach_min_gap: int = 10) -> Tuple[List[int], List[int], List[int]]:
    """-> (clusters, left_reaches, right_reaches) exactly as the reference leaves them in ctgVariants."""
    n = len(var)
    if not n:
        return [], [], []
    L = len(fasta)
    prev_clusters = list(range(n + 1))
    prev_active = [True] * (n + 1)
    left_reach, right_reach = [0] * (n + 1), [0] * (n + 1)
    it = 0
    while any(prev_active):                                                    # :979-981
        it += 1
        if it > max_iters:
            break
        clusters, nc = prev_clusters, len(prev_clusters)
        left_reach[nc - 1] = right_reach[nc - 1] = INT_MAX                     # sentinels, :995-996
        act = [c for c in range(nc - 1) if prev_active[c]]                     # :1002-1010
        # ---- alignment score of every active cluster against the reference (:1037-1046): one batch ----
        qs, ts, spans = [], [], []
        for c in act:
            first, last = clusters[c], clusters[c + 1] - 1
            beg = max(0, var[first][0] - 1)
            end = min(L, var[last][0] + var[last][1] + 1)
            qs.append(apply_variants(fasta, var, n, first, last + 1, beg, end))
            ts.append(fasta[beg:end])
            spans.append((first, last))
        scores = eng.wf_batch(WF_SCORE, qs, ts, sub, open_, extend) if act else []
        # ---- reaches: every (cluster, direction) search, one batch per doubling round (:1049-1158) ----
        searches: List[_Search] = []
        for c, (first, last), score in zip(act, spans, scores):
            main_diag = sum(v[1] - len(v[3]) for v in var[first:last + 1])     # :1062-1064
            for left in (True, False):
                s = _Search()
                s.c, s.left, s.first, s.last, s.score, s.main_diag = c, left, first, last, int(score), main_diag
                s.beg_pos = var[first][0] - 1
                s.end_pos = var[last][0] + var[last][1] + 1
                s.main_diag_start = s.end_pos - var[first][0] if left else var[last][0] + var[last][1] - s.beg_pos
                s.ref_len = s.end_pos - s.beg_pos
                s.reach = s.ref_len - 1
                s.done = False
                searches.append(s)
        pending = searches
        while pending:
            qs, ts, md, mds, ms, rv = [], [], [], [], [], []
            for s in pending:                                                  # window doubling, :1066-1077 / :1127-1137
                s.ref_len *= 2
                slack = abs(s.main_diag) + s.score // extend + 3
                if s.left:
                    s.beg_pos = max(0, s.end_pos - s.ref_len - slack)
                    q = apply_variants(fasta, var, n, s.first, s.last + 1, s.beg_pos, s.end_pos)
                    start = max(0, s.end_pos - s.ref_len)
                    r = fasta[start:start + s.ref_len]
                    q, r = q[::-1], r[::-1]
                else:
                    s.end_pos = min(L, s.beg_pos + s.ref_len + slack)
                    q = apply_variants(fasta, var, n, s.first, s.last + 1, s.beg_pos, s.end_pos)
                    r = fasta[s.beg_pos:s.beg_pos + min(s.ref_len, s.end_pos - s.beg_pos)]
                qs.append(q); ts.append(r)
                md.append(s.main_diag); mds.append(s.main_diag_start); ms.append(s.score); rv.append(1 if s.left else 0)
            reaches = eng.wf_batch(WF_REACH, qs, ts, sub, open_, extend, md, mds, ms, rv)
            nxt = []
            for s, reach in zip(pending, reaches):
                s.reach = int(reach)
                hit_edge = s.beg_pos == 0 if s.left else s.end_pos == L        # :1094, :1155
                if s.reach == s.ref_len - 1 and not hit_edge:
                    nxt.append(s)
            pending = nxt
        for s in searches:
            if s.left:
                left_reach[s.c] = s.end_pos - s.reach                          # :1097
            else:
                right_reach[s.c] = s.beg_pos + s.reach + 1                     # :1158
        # ---- merge rightwards, :1171-1192 ----
        tmp_clusters, tmp_active, tmp_left, tmp_right = [], [], [], []
        c = 0
        while c < nc:
            size = 1
            max_r, min_l = right_reach[c], left_reach[c]
            while c + size < nc and max_r + reach_min_gap >= left_reach[c + size]:
                max_r = max(max_r, right_reach[c + size])
                min_l = min(min_l, left_reach[c + size])
                size += 1
            tmp_right.append(max_r); tmp_left.append(min_l)
            tmp_clusters.append(prev_clusters[c]); tmp_active.append(size > 1)
            c += size
        # ---- merge leftwards, :1208-1231 ----
        next_clusters, next_active, left_reach, right_reach = [], [], [], []
        c = len(tmp_clusters) - 1
        while c >= 0:
            min_l, max_r, active = tmp_left[c], tmp_right[c], tmp_active[c]
            while c > 0 and min_l <= tmp_right[c - 1] + reach_min_gap:
                min_l = min(min_l, tmp_left[c - 1])
                max_r = max(max_r, tmp_right[c - 1])
                active = True
                c -= 1
            left_reach.append(min_l); right_reach.append(max_r)
            next_clusters.append(tmp_clusters[c]); next_active.append(active)
            c -= 1
        next_clusters.reverse(); next_active.reverse(); left_reach.reverse(); right_reach.reverse()
        prev_clusters, prev_active = next_clusters, next_active
    return prev_clusters, left_reach, right_reach
