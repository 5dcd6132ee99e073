"""TEST INFRASTRUCTURE - restatement of the reference's wf_swg_cluster
(/root/reference/src/cluster.cpp:954-1263): per-haplotype cluster growth by affine-gap wavefront reach
(the step immediately before the precision/recall path and the reference's wall-clock
bottleneck on SV input).  Host logic in plain Python over the two pinned C kernels of oracle/vd_reach.c
(wf_swg_max_reach, score of wf_swg_align); small inputs only.  Pinned against the reference's object
code by tests/test_reach_oracle.py.  No product code imports this.

Every cluster starts as one variant.  Per iteration and ACTIVE cluster (:1003-1160):
  score   = affine score of (reference window with the cluster's variants applied) vs (reference window),
            window = [first pos - 1, last end + 1)                                     (:1037-1046)
  left    = how far to the left a path of that score can wander: reversed strings, window doubled until
            the reach no longer hits its far end or the contig start                   (:1049-1097)
  right   = the same to the right                                                      (:1112-1158)
then clusters whose reaches come within reach_min_gap of each other are merged, rightwards then
leftwards (:1171-1238); merged clusters stay active.  At most max_iters iterations (:982)."""
from __future__ import annotations

from typing import List, Sequence, Tuple

from . import checkers

TYPE_SUB, TYPE_INS, TYPE_DEL = 1, 2, 3
INT_MAX = 2 ** 31 - 1


def generate_str(fasta: bytes, var: Sequence[tuple], n_all: int, beg_idx: int, end_idx: int, beg_pos: int, end_pos: int) -> bytes:
    """src/dist.cpp:81-136 (min_qual = 0): the window [beg_pos, end_pos) with variants beg_idx..end_idx-1
    applied.  var[i] = (pos, rlen, type, alt)."""
    out = bytearray()
    vi = beg_idx
    while vi < n_all and var[vi][0] < beg_pos:
        vi += 1
    ref_pos = beg_pos
    while ref_pos < end_pos:
        if vi < end_idx and ref_pos == var[vi][0]:
            pos, rlen, ty, alt = var[vi]
            if ty == TYPE_INS:
                out += alt
            elif ty == TYPE_DEL:
                ref_pos += rlen
            elif ty == TYPE_SUB:
                out += alt
                ref_pos += 1
            else:                       # TYPE_CPX
                out += alt
                ref_pos += rlen
            vi += 1
        else:
            ref_end = min(end_pos, var[vi][0]) if vi < end_idx else end_pos
            assert ref_end >= ref_pos, "No variant, but ref_end < ref_pos (generate_str)"
            assert ref_pos <= len(fasta), "position out of range (generate_str)"
            out += fasta[ref_pos:ref_end]
            ref_pos = ref_end
    return bytes(out)


def wf_swg_cluster(fasta: bytes, var: Sequence[tuple], sub: int, open_: int, extend: int,
                   max_iters: int = 4, reach_min_gap: int = 10) -> Tuple[List[int], List[int], List[int]]:
    """-> (clusters, left_reaches, right_reaches) as the reference leaves them in ctgVariants."""
    n = len(var)
    if not n:
        return [], [], []
    L = len(fasta)
    prev_clusters = list(range(n + 1))
    prev_active = [True] * (n + 1)
    left_reach, right_reach = [0] * (n + 1), [0] * (n + 1)
    it = 0
    while any(prev_active):                                            # :979-981
        it += 1
        if it > max_iters:
            break
        clusters = prev_clusters
        nc = len(prev_clusters)
        left_reach[nc - 1] = INT_MAX                                   # sentinels, :995-996
        right_reach[nc - 1] = INT_MAX
        for c in range(nc):
            compute = prev_active[c] and c != nc - 1                   # :1002-1010
            if not compute:
                continue
            first, last = clusters[c], clusters[c + 1] - 1
            main_diag = sum(v[1] - len(v[3]) for v in var[first:last + 1])      # :1062-1064
            # alignment score of the cluster against the reference, :1037-1046
            beg = max(0, var[first][0] - 1)
            end = min(L, var[last][0] + var[last][1] + 1)
            score = checkers.swg_score_oracle(generate_str(fasta, var, n, first, last + 1, beg, end), fasta[beg:end],
                                              sub, open_, extend)
            # left reach, :1049-1097
            beg_pos = var[first][0] - 1
            end_pos = var[last][0] + var[last][1] + 1
            main_diag_start = end_pos - var[first][0]
            ref_len = end_pos - beg_pos
            reach = ref_len - 1
            while reach == ref_len - 1:                                # iterative doubling
                ref_len *= 2
                beg_pos = max(0, end_pos - ref_len - abs(main_diag) - score // extend - 3)
                q = generate_str(fasta, var, n, first, last + 1, beg_pos, end_pos)
                start = max(0, end_pos - ref_len)
                r = fasta[start:start + ref_len]                       # std::string::substr(start, ref_len)
                reach = checkers.reach_oracle(q[::-1], r[::-1], main_diag, main_diag_start, score, sub, open_, extend, True)
                if beg_pos == 0:
                    break
            left_reach[c] = end_pos - reach
            # right reach, :1112-1158
            beg_pos = var[first][0] - 1
            end_pos = var[last][0] + var[last][1] + 1
            main_diag_start = var[last][0] + var[last][1] - beg_pos
            ref_len = end_pos - beg_pos
            reach = ref_len - 1
            while reach == ref_len - 1:
                ref_len *= 2
                end_pos = min(L, beg_pos + ref_len + abs(main_diag) + score // extend + 3)
                q = generate_str(fasta, var, n, first, last + 1, beg_pos, end_pos)
                r = fasta[beg_pos:beg_pos + min(ref_len, end_pos - beg_pos)]
                reach = checkers.reach_oracle(q, r, main_diag, main_diag_start, score, sub, open_, extend, False)
                if end_pos == L:
                    break
            right_reach[c] = beg_pos + reach + 1
        # merge rightwards, :1171-1192
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
        # merge leftwards, :1208-1231
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
