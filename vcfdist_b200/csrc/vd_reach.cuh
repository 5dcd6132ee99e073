// Affine-gap wavefront kernels of the cluster-growing stage: wf_swg_max_reach
// (src/dist.cpp:2150-2333) and the score of wf_swg_align (:1510-1652), the two functions wf_swg_cluster
// (src/cluster.cpp:954-1263) spends its time in - the reference runs them one (cluster, direction) after the
// other on one thread, which made clustering the wall-clock bottleneck on SV input.
//
// Diagonals across the threads of a problem: a warp for the many cluster-sized problems, a 256- or 1024-thread block or a
// thread-block cluster of eight blocks for the few whose wavefront grows wide (see wf_kernel).  Diagonal index d in [0, nd),
// nd = |query| + |truth| - 1, stands for k = d + 1 - |query| (truth index minus query index); three wavefront kinds (M, I, D)
// hold per diagonal the furthest QUERY index reached, NONE when the diagonal is not reached; only the last max(x, o+e) + 1
// scores are kept (a ring in HBM scratch, L1/L2-resident).  A score step reads other diagonals only from EARLIER scores, so
// all diagonals of a step are independent and a step needs one barrier; the free extension along matches is a per-thread
// byte-compare loop; the exit of the reference (the first diagonal, in ascending order, that reaches the end of a string) is
// the minimum over the threads that saw one.
#pragma once
#include "vd_common.cuh"

namespace vd {

constexpr int WF_NONE = -2;
enum { WF_M = 0, WF_I = 1, WF_D = 2, WF_NW = 3 };
enum { WF_MODE_REACH = 0, WF_MODE_SCORE = 1 };

struct WfBatch {
    int n;
    const int64_t *q_off, *t_off;         // [n+1] byte offsets
    const u8 *q_seq, *t_seq;
    const int32_t *main_diag, *main_diag_start, *max_score;   // reach only
    const u8 *reverse;                                        // reach only
    const int64_t *scratch_off;           // [n+1] int offsets into scratch
    int32_t *scratch;
    int32_t *result;                      // reach: furthest truth index; score: the alignment score
    int x, o, e, mode;
};

__host__ __device__ inline int wf_ring(int x, int o, int e) { return (x > o + e ? x : o + e) + 1; }
__host__ __device__ inline int64_t wf_ring_ints(int qlen, int tlen, int x, int o, int e) {
    return (int64_t)WF_NW * wf_ring(x, o, e) * (qlen + tlen - 1);
}
// the ring, then (8-byte aligned) four words the blocks of a cluster meet in: the first exit as a 64-bit key, the best of the last scan
__host__ __device__ inline int64_t wf_scratch_ints(int qlen, int tlen, int x, int o, int e) {
    return ((wf_ring_ints(qlen, tlen, x, o, e) + 1) & ~(int64_t)1) + 4;
}

// NT threads per problem: 32 (a warp; four problems per block) for the many cluster-sized problems, WF_BLOCK (a whole
// block) for the few whose wavefront grows to hundreds of diagonals - a structural variant's cluster, where one warp would
// walk the reached range in dozens of chunks per score.  sel lists the problems of this launch.  One barrier per score:
// a thread computes the new wavefronts of ITS diagonals from earlier scores, leaves the gaps, extends along matches and
// tests the exits without anybody else's values of the current score.
constexpr int WF_BLOCK = 256, WF_WIDE = 1024;

// The widest problems (a several-kb variant: ten thousand scores over twenty thousand diagonals) outgrow one SM's path to L2 -
// the ring is megabytes, every step streams it: CL blocks of one thread-block cluster share the diagonals of such a problem.
// They meet at the hardware cluster barrier once per score (release / acquire at cluster scope); ring values cross SMs, so
// they are read past L1 (ld.global.cg), and the exit is agreed on through a 64-bit atomicMin (diagonal, result) in scratch.
constexpr int WF_CLUSTER = 8;
#ifdef VD_EMU
__device__ inline void wf_cluster_sync() {}
__device__ inline unsigned wf_cluster_rank() { return 0; }
__device__ inline int wf_ld_l2(const int *p) { return *p; }
__device__ inline unsigned long long wf_ld_l2(const unsigned long long *p) { return *p; }
#else
__device__ __forceinline__ void wf_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ unsigned wf_cluster_rank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ int wf_ld_l2(const int *p) { return __ldcg(p); }
__device__ __forceinline__ unsigned long long wf_ld_l2(const unsigned long long *p) { return __ldcg(p); }
#endif

template <int NT, int CL = 1>
__global__ void __launch_bounds__(NT == 32 ? 128 : NT) wf_kernel(WfBatch B, const int *sel, int nsel) {
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int NO_HIT = 0x7fffffff;
    __shared__ int s_first, s_best;                                                   // block form only
    constexpr int NTT = NT * CL;                                                      // threads on one problem
    const int t = NT == 32 ? (threadIdx.x & 31) : CL > 1 ? (int)(wf_cluster_rank() * NT + threadIdx.x) : threadIdx.x;
    const int w = NT == 32 ? blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5) : blockIdx.x / CL;
    if (w >= nsel) return;
    const int p = sel ? sel[w] : w;
    const u8 *query = B.q_seq + B.q_off[p], *truth = B.t_seq + B.t_off[p];
    const int qlen = (int)(B.q_off[p + 1] - B.q_off[p]), tlen = (int)(B.t_off[p + 1] - B.t_off[p]);
    const int x = B.x, o = B.o, e = B.e;
    const bool reach = B.mode == WF_MODE_REACH;
    const bool reverse = reach && B.reverse[p];
    const int nd = qlen + tlen - 1, ring = wf_ring(x, o, e);
    int *v = B.scratch + B.scratch_off[p];
    const int64_t ring_ints = (int64_t)WF_NW * ring * nd;
    unsigned long long *x_first = (unsigned long long *)(v + ((ring_ints + 1) & ~(int64_t)1));   // cluster form: (diagonal << 32) | result of the first exit
    int *x_best = (int *)(x_first + 1);
    auto ld = [&](const int *q) -> int { return CL > 1 ? wf_ld_l2(q) : *q; };
    auto wf = [&](int kind, int slot, int d) -> int * { return v + ((int64_t)kind * ring + slot) * nd + d; };
    auto slot_of = [&](int slot, int back) { const int s = slot - back; return s < 0 ? s + ring : s; };
    auto inside = [&](int q, int k) { return q >= 0 && q < qlen && k + q >= 0 && k + q < tlen; };
    // true if any thread of the problem says so; also the barrier after which the wavefronts of this score are everybody's
    auto any_of = [&](bool f) -> bool {
        if (NT == 32) { const bool r = __any_sync(FULL, f); __syncwarp(); return r; }
        return __syncthreads_or(f) != 0;
    };
    auto barrier = [&]() { if (NT == 32) __syncwarp(); else if (CL > 1) wf_cluster_sync(); else __syncthreads(); };
    for (int64_t i = t; i < ring_ints; i += NTT) v[i] = WF_NONE;
    if (NT != 32 && threadIdx.x == 0) { s_first = NO_HIT; s_best = 0; }
    if (CL > 1 && t == 0) { *x_first = ~0ull; *x_best = 0; }
    const int main_diag = reach ? B.main_diag[p] : 0;
    const int stop_q = reach ? B.main_diag_start[p] - main_diag : 0;                // :2160
    const int max_score = reach ? B.max_score[p] : 0x7fffffff;
    int score = 0, slot = 0;
    barrier();
    if (t == 0) *wf(WF_M, slot, qlen - 1) = -1;                                       // :2166 / :1528: diagonal k = 0, before the first base
    barrier();
    // Diagonals any wavefront has reached so far: [dlo, dhi].  A score step reaches at most one more on either side (its
    // sources are the same diagonal and its two neighbours at earlier scores), everything outside still holds NONE from the
    // fill above - so a step runs over this range and not over all |query| + |truth| - 1 diagonals: a structural
    // variant's cluster has thousands of diagonals and hundreds of score steps, of which a step touches a few dozen.
    int dlo = qlen - 1, dhi = qlen - 1;
    for (;;) {
        int hit_d = NO_HIT, hit_res = -1;
        // rows of the ring this score reads and writes (index arithmetic once per score, not per diagonal)
        const int cost_open = reverse ? e : o + e;
        int *const Mc = wf(WF_M, slot, 0), *const Ic = wf(WF_I, slot, 0), *const Dc = wf(WF_D, slot, 0);
        const int *const Mx = wf(WF_M, slot_of(slot, x), 0), *const Mg = wf(WF_M, slot_of(slot, cost_open), 0);
        const int *const Io = wf(WF_I, slot_of(slot, o), 0), *const Do = wf(WF_D, slot_of(slot, o), 0);
        const int *const Ie = wf(WF_I, slot_of(slot, e), 0), *const De = wf(WF_D, slot_of(slot, e), 0);
        const bool has_x = score > 0 && score - x >= 0, has_g = score > 0 && score - cost_open >= 0;
        const bool has_o = score > 0 && reverse && score - o >= 0, has_e = score > 0 && score - e >= 0;
        for (int d = dlo + t; d <= dhi; d += NTT) {
            const int k = d + 1 - qlen;
            // the wavefronts of this score from those of earlier scores (:2225-2311, :1583-1650).  :2228-2232 clears I and D
            // only, the M wavefront of the slot is overwritten through >= tests; wf_swg_align starts every wavefront of a
            // new score empty (:1583-1587)
            int mc = (reach || score == 0) ? ld(Mc + d) : WF_NONE, ic = WF_NONE, dc = WF_NONE;
            if (has_x) {                                                              // substitution (:2239-2250)
                const int pv = x == 0 ? mc : ld(Mx + d);
                if (pv != WF_NONE && pv + 1 < qlen && k + pv + 1 < tlen && pv + 1 >= mc) mc = pv + 1;
            }
            if (has_g) {                                                              // gap opening (:2252-2275)
                if (d > 0) {
                    const int pv = ld(Mg + d - 1);
                    if (pv != WF_NONE && k + pv < tlen && pv >= dc) dc = pv;
                }
                if (d < nd - 1) {
                    const int pv = ld(Mg + d + 1);
                    if (pv != WF_NONE && pv + 1 < qlen && k + pv + 1 < tlen && k + pv + 1 >= 0 && pv + 1 >= ic) ic = pv + 1;
                }
            }
            if (has_o) {                                                              // reversed problem: leaving a gap costs o (:2277-2294)
                const int pi = o == 0 ? ic : ld(Io + d), pd = o == 0 ? dc : ld(Do + d);
                if (inside(pi, k) && pi > mc) mc = pi;
                if (inside(pd, k) && pd > mc) mc = pd;
            }
            if (has_e) {                                                              // gap extension (:2296-2311)
                if (d > 0) {
                    const int pv = ld(De + d - 1);
                    if (pv != WF_NONE && k + pv < tlen && pv >= dc) dc = pv;
                }
                if (d < nd - 1) {
                    const int pv = ld(Ie + d + 1);
                    if (pv != WF_NONE && pv + 1 < qlen && k + pv + 1 < tlen && k + pv + 1 >= 0 && pv + 1 >= ic) ic = pv + 1;
                }
            }
            if (score > 0) { Ic[d] = ic; Dc[d] = dc; }
            // gaps are left for free at the score they were reached with (:2171-2184, :1533-1547); not in the reversed problem
            int q = mc;
            if (!reverse) {
                if (inside(ic, k) && ic >= q) q = ic;
                if (inside(dc, k) && dc >= q) q = dc;
            }
            // free extension along matches, then the exits, diagonals in ascending order (:2187-2212, :1550-1568)
            while ((!reach || k != main_diag || q + 1 < stop_q) && q != WF_NONE && k + q >= -1 &&
                   q < qlen - 1 && k + q < tlen - 1 && query[q + 1] == truth[k + q + 1])
                q++;
            Mc[d] = q;
            if (reach) {
                if (q + k == tlen - 1) { hit_d = d; hit_res = tlen - 1; break; }                                   // :2205-2207
                if (q == qlen - 1 && q + k >= 0 && q + k < tlen - 1) { hit_d = d; hit_res = q + k; break; }        // :2208-2210
            } else if (q == qlen - 1 && q + k == tlen - 1) { hit_d = d; hit_res = score; break; }
        }
        if (CL > 1) {
            // blocks of a cluster: the exits meet in scratch, everybody reads the outcome after the cluster barrier
            if (hit_d != NO_HIT) atomicMin(x_first, ((unsigned long long)(unsigned)hit_d << 32) | (unsigned)hit_res);
            wf_cluster_sync();
            const unsigned long long key = wf_ld_l2(x_first);
            if (key != ~0ull) {
                if (t == 0) B.result[p] = (int)(unsigned)(key & 0xffffffffu);
                return;
            }
        } else if (any_of(hit_d != NO_HIT)) {
            // the reference leaves at the FIRST diagonal in ascending order that reached an end
            int first;
            if (NT == 32) first = __reduce_min_sync(FULL, hit_d);
            else {
                if (hit_d != NO_HIT) atomicMin(&s_first, hit_d);
                __syncthreads();
                first = s_first;
            }
            if (hit_d == first) B.result[p] = hit_res;
            return;
        }
        if (score == max_score) break;                                                // :2213
        score++;
        slot = slot + 1 == ring ? 0 : slot + 1;
        dlo = max(dlo - 1, 0); dhi = min(dhi + 1, nd - 1);
    }
    // the score budget is spent: furthest truth index over everything still in the ring (:2316-2331)
    int best = 0;
    for (int64_t i = t; i < ring_ints; i += NTT) {
        const int d = (int)(i % nd);
        const int q = ld(v + i), k = d + 1 - qlen;
        if (inside(q, k) && k + q > best) best = k + q;
    }
    if (NT == 32) best = __reduce_max_sync(FULL, best);
    else if (CL > 1) {
        if (best > 0) atomicMax(x_best, best);
        wf_cluster_sync();
        best = wf_ld_l2(x_best);
    } else {
        atomicMax(&s_best, best);
        __syncthreads();
        best = s_best;
    }
    if (t == 0) B.result[p] = best;
}


// ------------------------------------------------------------------------------------------
// `--distance` pass: wf_swg_align WITH its predecessor flags (src/dist.cpp:1510-1652) and
// wf_swg_backtrack (:2625-2757) - the affine-gap alignment edits_wrapper (:1908-2077) runs per supercluster and
// haplotype.  Two passes: the score kernel above sizes the storage (every score keeps its three wavefronts and
// one flag byte per diagonal), this kernel recomputes the wavefronts with their flags - diagonals across the
// lanes of one warp per problem - and lane 0 walks back from the last cell: on the M wavefront leaving an
// insertion is preferred over leaving a deletion over a substitution (:2662-2700), inside a gap extending over
// opening (:2716-2745).  cigar[] has |query| + |truth| entries filled from the back as the reference does: two per
// match / substitution, one per inserted / deleted base, zeros in front.
// ------------------------------------------------------------------------------------------
enum { WFC_INS = 1, WFC_DEL = 2, WFC_MAT = 4, WFC_SUB = 8 };

struct WfCigarBatch {
    int n;
    const int64_t *q_off, *t_off;
    const u8 *q_seq, *t_seq;
    const int32_t *score;                 // from the score pass
    const int64_t *store_off;             // [n+1] byte offsets into store
    u8 *store;                            // per problem: int off[score+1][3][nd], then u8 flag[score+1][3][nd]
    int32_t *cigar;                       // problem p at q_off[p] + t_off[p]
    int32_t *result;                      // the score, -1 where the walk back fails (the reference would ERROR())
    int x, o, e;
};
__host__ __device__ inline int64_t wf_cigar_bytes(int qlen, int tlen, int score) {
    const int64_t cells = (int64_t)(score + 1) * WF_NW * (qlen + tlen - 1);
    return (5 * cells + 15) / 16 * 16;
}

template <int NT>                                             // threads per problem, as wf_kernel
__global__ void __launch_bounds__(NT == 32 ? 128 : NT) wf_cigar_kernel(WfCigarBatch B, const int *sel, int nsel) {
    const int lane = NT == 32 ? (threadIdx.x & 31) : threadIdx.x;
    const int w = NT == 32 ? blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5) : blockIdx.x;
    if (w >= nsel) return;
    const int p = sel ? sel[w] : w;
    auto barrier = [&]() { if (NT == 32) __syncwarp(); else __syncthreads(); };
    const u8 *query = B.q_seq + B.q_off[p], *truth = B.t_seq + B.t_off[p];
    const int qlen = (int)(B.q_off[p + 1] - B.q_off[p]), tlen = (int)(B.t_off[p + 1] - B.t_off[p]);
    const int x = B.x, o = B.o, e = B.e, final_score = B.score[p];
    const int nd = qlen + tlen - 1;
    const int64_t cells = (int64_t)(final_score + 1) * WF_NW * nd;
    int *off = (int *)(B.store + B.store_off[p]);
    u8 *flag = (u8 *)(off + cells);
    auto OFF = [&](int s, int kind, int d) -> int & { return off[((int64_t)s * WF_NW + kind) * nd + d]; };
    auto FLG = [&](int s, int kind, int d) -> u8 & { return flag[((int64_t)s * WF_NW + kind) * nd + d]; };
    for (int64_t i = lane; i < cells; i += NT) { off[i] = WF_NONE; flag[i] = 0; }
    barrier();
    if (lane == 0) { OFF(0, WF_M, qlen - 1) = -1; FLG(0, WF_M, qlen - 1) = WFC_MAT; }                  // :1528-1529
    barrier();
    int dlo = qlen - 1, dhi = qlen - 1;                      // diagonals reached so far (see wf_kernel): the passes run over these only
    for (int score = 0;; score++) {
        for (int d = dlo + lane; d <= dhi; d += NT) {                                                     // :1533-1547
            const int k = d + 1 - qlen;
#pragma unroll
            for (int kind = WF_I; kind <= WF_D; kind++) {
                const int q = OFF(score, kind, d);
                if (q >= 0 && q < qlen && k + q >= 0 && k + q < tlen && q >= OFF(score, WF_M, d)) {
                    OFF(score, WF_M, d) = q;
                    FLG(score, WF_M, d) |= (kind == WF_I) ? WFC_INS : WFC_DEL;
                }
            }
            int q = OFF(score, WF_M, d);                                                                  // :1550-1568
            while (q != WF_NONE && k + q >= -1 && q < qlen - 1 && k + q < tlen - 1 && query[q + 1] == truth[k + q + 1]) q++;
            OFF(score, WF_M, d) = q;
        }
        barrier();
        if (score == final_score) break;                     // the score pass found the end at this score
        const int s1 = score + 1;
        dlo = max(dlo - 1, 0); dhi = min(dhi + 1, nd - 1);
        for (int d = dlo + lane; d <= dhi; d += NT) {
            const int k = d + 1 - qlen;
            if (s1 - x >= 0) {                                                                            // :1592-1600
                const int pv = OFF(s1 - x, WF_M, d);
                if (pv != WF_NONE && pv + 1 < qlen && k + pv + 1 < tlen && pv + 1 >= OFF(s1, WF_M, d)) { OFF(s1, WF_M, d) = pv + 1; FLG(s1, WF_M, d) |= WFC_SUB; }
            }
            if (s1 - (o + e) >= 0) {                                                                      // :1602-1625
                const int ps = s1 - (o + e);
                if (d > 0) {
                    const int pv = OFF(ps, WF_M, d - 1);
                    if (pv != WF_NONE && k + pv < tlen && pv >= OFF(s1, WF_D, d)) { OFF(s1, WF_D, d) = pv; FLG(s1, WF_D, d) |= WFC_SUB; }
                }
                if (d < nd - 1) {
                    const int pv = OFF(ps, WF_M, d + 1);
                    if (pv != WF_NONE && pv + 1 < qlen && k + pv + 1 < tlen && k + pv + 1 >= 0 && pv + 1 >= OFF(s1, WF_I, d)) { OFF(s1, WF_I, d) = pv + 1; FLG(s1, WF_I, d) |= WFC_SUB; }
                }
            }
            if (s1 - e >= 0) {                                                                            // :1627-1650
                const int ps = s1 - e;
                if (d > 0) {
                    const int pv = OFF(ps, WF_D, d - 1);
                    if (pv != WF_NONE && k + pv < tlen && pv >= OFF(s1, WF_D, d)) { OFF(s1, WF_D, d) = pv; FLG(s1, WF_D, d) |= WFC_DEL; }
                }
                if (d < nd - 1) {
                    const int pv = OFF(ps, WF_I, d + 1);
                    if (pv != WF_NONE && pv + 1 < qlen && k + pv + 1 < tlen && k + pv + 1 >= 0 && pv + 1 >= OFF(s1, WF_I, d)) { OFF(s1, WF_I, d) = pv + 1; FLG(s1, WF_I, d) |= WFC_INS; }
                }
            }
        }
        barrier();
    }
    // ---- walk back (:2648-2755), one lane ----
    int *cigar = B.cigar + B.q_off[p] + B.t_off[p];
    for (int i = lane; i < qlen + tlen; i += NT) cigar[i] = 0;
    barrier();
    if (lane) return;
    int cp = qlen + tlen - 1, kind = WF_M, qi = qlen - 1, ti = tlen - 1, s = final_score;
    bool failed = false;
    while ((qi >= 0 || ti >= 0) && !failed) {
        if (s < 0) { failed = true; break; }
        const int d = qlen - 1 + (ti - qi);
        if (kind == WF_M) {
            const int f = FLG(s, WF_M, d);
            if (f & (WFC_INS | WFC_DEL)) {                                    // a gap was left here for free
                const int g = (f & WFC_INS) ? WF_I : WF_D;
                const int stop = OFF(s, g, d);
                while (qi > stop) { cigar[cp--] = WFC_MAT; cigar[cp--] = WFC_MAT; qi--; ti--; if (qi < 0 || ti < 0) { failed = true; break; } }
                kind = g;
            } else if (f & WFC_SUB) {
                if (s - x < 0) { failed = true; break; }
                const int stop = OFF(s - x, WF_M, d) + 1;
                while (qi > stop) { cigar[cp--] = WFC_MAT; cigar[cp--] = WFC_MAT; qi--; ti--; if (qi < 0 || ti < 0) { failed = true; break; } }
                if (failed) break;
                cigar[cp--] = WFC_SUB; cigar[cp--] = WFC_SUB; qi--; ti--;
                s -= x;
            } else if (f & WFC_MAT) {
                while (qi >= 0 && ti >= 0) { cigar[cp--] = WFC_MAT; cigar[cp--] = WFC_MAT; qi--; ti--; }
                if (qi >= 0 || ti >= 0) failed = true;
            } else failed = true;
        } else if (kind == WF_I) {
            const int f = FLG(s, WF_I, d);
            if (f & WFC_INS) { cigar[cp--] = WFC_INS; qi--; s -= e; }
            else if (f & WFC_SUB) { cigar[cp--] = WFC_INS; qi--; kind = WF_M; s -= o + e; }
            else failed = true;
        } else {
            const int f = FLG(s, WF_D, d);
            if (f & WFC_DEL) { cigar[cp--] = WFC_DEL; ti--; s -= e; }
            else if (f & WFC_SUB) { cigar[cp--] = WFC_DEL; ti--; kind = WF_M; s -= o + e; }
            else failed = true;
        }
        if (!(qi == -1 && ti == -1) && (qi < 0 || ti < 0)) failed = true;
    }
    B.result[p] = failed ? -1 : final_score;
}

}  // namespace vd
