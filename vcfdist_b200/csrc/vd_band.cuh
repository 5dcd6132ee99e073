// Banded warp-per-alignment kernels of the long path: forward (calc_prec_recall_aln, src/dist.cpp:251-443)
// and backward (calc_prec_recall_path, :486-834) column sweeps in which ONE WARP owns an alignment and no
// block barrier is ever taken.
//
// The reference explores the two-plane graph by increasing score and stops when the end is reached
// (:309-443), so it only ever touches cells whose distance is at most the final score.  A structural
// variant that truth and query both carry gives matrices of thousands x thousands with a score of a few
// dozen: the cells that matter are a narrow band around one diagonal per plane.  These kernels keep, per
// plane, a WINDOW of 32*K consecutive rows that slides with the band:
//
//   * score bound tau (Ukkonen's cut-off) sharpened by an exact lower bound of the remaining cost
//     (wave_row_hulls in vd_wave.cuh): a cell is kept while D + bound <= tau.  Every kept cell has its exact
//     distance and its complete set of optimal predecessors, so the result is exact whenever the final score
//     is <= tau; otherwise the next rung (wider window, larger tau) tries again;
//   * rows are dealt in blocks of K consecutive rows, block b always to lane b % 32; the window of a column
//     is the 32 blocks starting at the block of the lowest candidate row;
//   * the column step of a lane is a serial chain over its K rows in registers (diagonal, deletion, swap
//     candidate, the in-column insertion chain), then ONE warp scan hands the chain across lanes
//     (min-plus prefix scan forward, link-segmented suffix max of T - S backward), then a second pass over
//     the K rows finalises values and flags;
//   * nothing on the per-column dependency chain touches HBM: the packed per-row records (FwdRow / BwdRow,
//     16 bytes a row) stream through a shared-memory ring in 512-byte chunks pulled by the TMA copy engine
//     (cp.async.bulk + mbarrier) a window ahead of the sweep; the truth bases and the band records arrive in
//     32-column chunks; swap edges read the other plane's previous column from a ring in shared memory;
//     the next columns' candidate rows come from warp reductions over the kept cells;
//   * flags are stored BANDED: F[column][plane][row mod 32K], one byte per cell of the window.
//
// Rungs (rows per lane / score bound): 1/14, 2/30, 4/60, 8/120, 16/240 (a window holds 2 tau + 2 rows).  A rung whose window holds both
// planes entirely (<= 32*K rows each) runs unbounded: the dense sweep of thin-but-wide matrices.  What no
// rung solves (score above the last bound, or kept cells further apart than a window) goes to the dense
// block-per-alignment kernels of vd_wave.cuh.
#pragma once
#include "vd_wave.cuh"

namespace vd {

#ifndef VD_BAND_WARPS
#define VD_BAND_WARPS 1
#endif
constexpr int BAND_WARPS = VD_BAND_WARPS;                      // alignments per block (one warp each).  One: a block's shared memory
// (43 KB per warp at K = 16) is held as long as its slowest warp runs, and with four warps a block of the wide rungs - typically
// one live alignment and three that are not this rung's - kept a whole SM's shared memory from the other rungs' blocks
constexpr int N_RUNG = 5;
inline int band_rung_k(int r) { const int v[N_RUNG] = {1, 2, 4, 8, 16}; return v[r]; }
__host__ __device__ inline int band_rung_tau(int r) { const int v[N_RUNG] = {14, 30, 60, 120, 240}; return v[r]; }
constexpr int BAND_DONE = 0x100;      // state bit: walked (the rungs run in two rounds, see vd_api.cu)

// per-item state shared by the rungs
constexpr int BAND_PENDING = 0;       // not solved yet
constexpr int BAND_DENSE = -1;        // given up: dense kernels
// > 0: K of the rung that solved it

struct BandCtx {
    int sc, ai, Lq, Lr, Lt;
    const u8 *tinfo;
    const FwdRow *frow[2];            // QUERY plane rows, REF plane rows
    const BwdRow *brow[2];
    const int *tab[2];                // CSR swap-source tables of destination rows (toQ, toR)
    u8 *F;                            // banded flags, [Lt][2][32*kmax]
    int4 *band;                       // [Lt] candidate rows (loQ, hiQ, loR, hiR), empty: lo > hi
    int kmax;
};

__device__ inline BandCtx band_ctx(const WaveArgs &A, int item) {
    BandCtx x;
    const int e = item >> 2;
    x.ai = item & 3;
    const int i = A.i0 + e;
    x.sc = A.list[i];
    const ScPlan p = A.plan[x.sc];
    const WaveSlab W = make_wave_slab(p);
    u8 *base = A.slab + (A.offs[i] - A.offs[A.i0]);
    const int qh = x.ai >> 1, th = x.ai & 1;
    x.Lq = p.len[qh]; x.Lr = p.lr; x.Lt = p.len[2 + th];
    const WaveAln wa = wave_aln(x.Lq, x.Lr, x.Lt);
    SlabQm M(base + W.base.qm[qh], x.Lq, x.Lr);
    WaveHapQ T(base + W.hq[qh], x.Lq, x.Lr);
    x.tinfo = base + W.ht[th];
    x.frow[0] = T.fwdQ; x.frow[1] = T.fwdR;
    x.brow[0] = T.bwdQ; x.brow[1] = T.bwdR;
    x.tab[0] = M.toQ; x.tab[1] = M.toR;
    x.F = base + W.aln[x.ai] + wa.oF;
    x.band = (int4 *)(base + W.aln[x.ai] + wa.oBand);
    x.kmax = wa.kmax;
    return x;
}

// Lower bound of the score (see wave_fwdb_kernel): every path spells the reference with a SUBSET of the query
// haplotype's variants applied, and aligning a string of length Lm to the truth costs at least |Lm - Lt|.
__device__ inline int band_score_lb(const BatchDev &in, int sc, int ai, int Lr, int Lt) {
    const int qh = ai >> 1;
    const int64_t v0 = in.var_off[4 * (int64_t)sc + qh], v1 = in.var_off[4 * (int64_t)sc + qh + 1];
    const int nv = (int)(v1 - v0);
    if (nv > 10) return 0;
    int best = INF;
    for (int s = 0; s < (1 << nv); s++) {
        int lm = Lr;
        for (int j = 0; j < nv; j++)
            if ((s >> j) & 1) lm += (int)(in.alt_off[v0 + j + 1] - in.alt_off[v0 + j]) - in.var_rlen[v0 + j];
        best = min(best, abs(lm - Lt));
    }
    return best;
}

// several swap sources for one destination row (an insertion's first base, adjacent deletions): the smallest
// value wins, the larger row on equal values, and the tie is recorded (src/dist.cpp:347, :376 keep the last
// writer, whose identity depends on hash-set order).  src: the row's source list, src[0] already looked at.
__device__ VD_NOINLINE int2 band_swap_multi(const int *src, int cnt, const int *rprev_o, int pblo_o, int K, int best) {
    int sb = 0;
    const int W = 32 * K;
    for (int k = 1; k < cnt; k++) {
        const int s = src[k];
        int v = INF;
        if ((unsigned)(s / K - pblo_o) < 32u) v = rprev_o[s & (W - 1)];
        if (v < best) { best = v; sb = k << F_K_SHIFT; }
        else if (v == best && v < INF / 2) sb = (k << F_K_SHIFT) | F_TIE;   // keep the larger row
    }
    return make_int2(best, sb);
}

// ------------------------------------------------------------------------------------------
// One plane's per-row records streaming through a shared-memory ring of NS 32-row chunks (512 bytes each):
// the chunk is the unit of the TMA bulk copy, slot = chunk % NS, one mbarrier per slot.  All members are
// warp-uniform; lane 0 issues the copies, every lane waits on the barrier of a chunk before its first use.
// DESC: the sweep moves towards lower rows (backward kernel).
// ------------------------------------------------------------------------------------------
template <int NS, bool DESC>
struct RowStream {
    const uint4 *src;                 // global array, 16 bytes a row, padded to whole chunks
    uint4 *ring;                      // shared: NS * 32 rows
    unsigned long long *bar;          // shared: NS mbarriers
    int nchunk;
    int vlo, vhi;                     // chunks issued and not overwritten since: [vlo, vhi], empty when vlo > vhi
    int wlo, whi;                     // of those, the chunks already waited for: [wlo, whi]
    unsigned phase;                   // per slot: parity of its next completion

    __device__ __forceinline__ void init(const void *s, uint4 *r, unsigned long long *b, int rows) {
        src = (const uint4 *)s; ring = r; bar = b; nchunk = (rows + 31) >> 5;
        vlo = wlo = 0; vhi = whi = -1; phase = 0;
    }
    __device__ __forceinline__ void issue(int ch, int lane) {
        if (lane == 0) VD_BULK_G2S(ring + (ch & (NS - 1)) * 32, src + (int64_t)ch * 32, 512u, bar + (ch & (NS - 1)));
    }
    __device__ __forceinline__ void wait(int ch) {
        const int slot = ch & (NS - 1);
        VD_MBAR_WAIT(bar + slot, (phase >> slot) & 1u);
        phase ^= 1u << slot;
    }
    // rows [lo, hi] must be readable on return; up to `pref` further chunks are requested ahead of the sweep
    __device__ __forceinline__ void need(int lo, int hi, int lane, int pref) {
        const int cl = lo >> 5, ch = hi >> 5;
        const bool restart = vlo > vhi || (DESC ? (ch > vhi || ch < vlo - 1) : (cl < vlo || cl > vhi + 1));
        if (restart) {
            if (!DESC) { while (whi < vhi) wait(++whi); } else { while (wlo > vlo) wait(--wlo); }   // drain: keep the parities in step
            if (!DESC) { vlo = wlo = cl; vhi = whi = cl - 1; } else { vhi = whi = ch; vlo = wlo = ch + 1; }
        }
        if (!DESC) {
            if (cl > vlo) { vlo = cl; if (wlo < vlo) wlo = vlo; }                                   // chunks below the window may be overwritten
            const int target = min(nchunk - 1, ch + pref);
            while (vhi < target && vhi + 1 - vlo < NS) issue(++vhi, lane);
            while (whi < ch) wait(++whi);
        } else {
            if (ch < vhi) { vhi = ch; if (whi > vhi) whi = vhi; }
            const int target = max(0, cl - pref);
            while (vlo > target && vhi - (vlo - 1) < NS) issue(--vlo, lane);
            while (wlo > cl) wait(--wlo);
        }
    }
};

template <int K> struct BandCfg {
    static constexpr int W = 32 * K;                               // rows of a window
    static constexpr int R = 2 * W > 128 ? 2 * W : 128;            // rows of the record ring
    static constexpr int NS = R / 32;
    static constexpr int PREF = NS - (K + 1) < K ? NS - (K + 1) : K;   // chunks requested ahead of the window
    // per warp: value ring [2 parities][2 planes][W] ints, record ring [2 planes][R] x 16 B, band chunks [2][32] x 16 B, barriers
    static constexpr int SMEM = 16 * W + 32 * R + 1024 + 16 * NS;
};

// resident blocks per SM the register allocation must allow, by rung: the narrow rungs take tens of thousands of alignments
// (bound by the number of resident warps), the wide ones a few hundred long ones (bound by one warp's column step)
#ifndef VD_BAND_MINB_NARROW
#define VD_BAND_MINB_NARROW 1
#endif
constexpr int band_minb(int K) { return K <= 2 ? VD_BAND_MINB_NARROW : 1; }

// ------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------
// hint[]: first rung worth trying for an item.  A sweep that runs out of budget at column c has seen about
// tau + 1 edits in c columns: the rung whose bound covers that rate over the whole truth haplotype is tried next
// instead of every rung in between (a guess about speed only - whatever rung solves an item solves it exactly).
// exact_hint: take only the items whose hint IS this rung (first round, all rungs side by side); otherwise
// every pending item whose hint is not above it (second round, rung after rung).
template <int K>
__global__ void __launch_bounds__(32 * BAND_WARPS, band_minb(K)) band_fwd_kernel(WaveArgs A, int n_items, int *state, const int *lbound, int *hint,
                                                                   int rung, int exact_hint, WaveItems *wi) {
    VD_DYN_SHARED(smem_raw);
    typedef BandCfg<K> C;
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int W = C::W, R = C::R, NS = C::NS;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int idx = n_items - 1 - (blockIdx.x * BAND_WARPS + warp);          // biggest shape classes first
    if (idx < 0) return;
    if (state[idx] != BAND_PENDING) return;
    if (exact_hint ? hint[idx] != rung : hint[idx] > rung) return;
    const bool last_rung = rung == N_RUNG - 1;
    const BandCtx X = band_ctx(A, A.items[idx]);
    const int len[2] = {X.Lq, X.Lr};
    const bool unbounded = (X.Lq + K - 1) / K <= 32 && (X.Lr + K - 1) / K <= 32;
    const int tau = unbounded ? INF : band_rung_tau(rung);
    if (K > X.kmax || max(X.Lq, X.Lr) >= BAND_ROW_LIMIT || (!unbounded && lbound[idx] > tau)) {   // this rung cannot take it
        if (lane == 0) { if (last_rung) state[idx] = BAND_DENSE; else hint[idx] = rung + 1; }
        return;
    }
    u8 *sm = smem_raw + warp * C::SMEM;
    int *ring = (int *)sm;                                                    // [column parity][plane][W]: D of the previous column
    uint4 *srow = (uint4 *)(sm + 16 * W);                                     // [plane][R]: FwdRow records
    unsigned long long *bars = (unsigned long long *)(sm + 16 * W + 32 * R + 1024);
    if (lane == 0) for (int i = 0; i < 2 * NS; i++) VD_MBAR_INIT(bars + i, 1);
    VD_MBAR_INIT_FENCE();
    __syncwarp();
    RowStream<NS, false> strm[2];
    strm[0].init(X.frow[0], srow, bars, X.Lq);
    strm[1].init(X.frow[1], srow + R, bars + NS, X.Lr);
    u8 *F = X.F;
    const int WS = 32 * X.kmax;                                               // row stride of the flag storage (sized for kmax)

    int blk[2] = {lane, lane};                                                // the window starts at block 0
    int Dp[2][K];
#pragma unroll
    for (int P = 0; P < 2; P++)
#pragma unroll
        for (int j = 0; j < K; j++) Dp[P][j] = INF;
    int nextLo[2] = {0, 0}, nextHi[2] = {0, 0};                               // candidate rows of the next column (from the kept cells of this one)
    int pblo[2] = {0, 0};                                                     // first block of the previous column's window
    bool pvalid[2] = {false, false};                                          // the ring holds the plane's previous column
    bool failed = false;
    int c_fail = -1;                                                          // column at which the budget ran out
    long long visited = 0;                                                    // candidate cells of the sweep (statistics)
    // truth bases in 32-column chunks: one coalesced load per chunk, a chunk ahead
    int tchunk = lane < X.Lt ? X.tinfo[lane] : 0, tchunk_next = 32 + lane < X.Lt ? X.tinfo[32 + lane] : 0;
    for (int c = 0; c < X.Lt; c++) {
        if ((c & 31) == 0 && c) { tchunk = tchunk_next; tchunk_next = c + 32 + lane < X.Lt ? X.tinfo[c + 32 + lane] : 0; }
        const int tin = __shfl_sync(FULL, tchunk, c & 31);
        const int tch = tin & 0x7f;
        const bool tok = tin & 0x80;
        // ---- candidate rows of both planes (plane-local, inclusive) ----
        int cLo[2], cHi[2];
        bool has[2];
        int blo[2];
#pragma unroll
        for (int P = 0; P < 2; P++) {
            if (c == 0 || unbounded) { cLo[P] = 0; cHi[P] = unbounded ? len[P] - 1 : min(tau, len[P] - 1); }
            else { cLo[P] = max(nextLo[P], 0); cHi[P] = min(nextHi[P], len[P] - 1); }
            has[P] = cHi[P] >= cLo[P];
            blo[P] = has[P] ? cLo[P] / K : pblo[P];
            if (has[P] && cHi[P] / K - blo[P] > 31) failed = true;            // kept cells further apart than the window
        }
        if (!unbounded && c > 0 && !has[0] && !has[1]) { failed = true; c_fail = c; }   // nothing within the bound is left
#ifdef VD_BAND_DEBUG
        if (failed && lane == 0) printf("  fail at c=%d cand Q[%d,%d] R[%d,%d]\n", c, cLo[0], cHi[0], cLo[1], cHi[1]);
#endif
        if (failed) break;
        visited += (has[0] ? cHi[0] - cLo[0] + 1 : 0) + (has[1] ? cHi[1] - cLo[1] + 1 : 0);
        if (lane == 0) X.band[c] = make_int4(has[0] ? cLo[0] : 1, has[0] ? cHi[0] : 0, has[1] ? cLo[1] : 1, has[1] ? cHi[1] : 0);
        const int *rprev = ring + ((c + 1) & 1) * 2 * W;
        int *rcur = ring + (c & 1) * 2 * W;
        int nlo[2] = {INF, INF}, nhi[2] = {-1, -1};
#pragma unroll
        for (int P = 0; P < 2; P++) {
            if (!has[P]) {                                                    // plane without candidates: nothing kept in this column
                if (pvalid[P] || c == 0) {
#pragma unroll
                    for (int j = 0; j < K; j++) Dp[P][j] = INF;
                }
                continue;
            }
            strm[P].need(cLo[P], cHi[P], lane, C::PREF);
            const uint4 *rows = srow + P * R;
            // ---- my block in this column's window ----
            const int nb = blo[P] + ((lane - blo[P]) & 31);
            if (nb != blk[P]) {
                blk[P] = nb;
#pragma unroll
                for (int j = 0; j < K; j++) Dp[P][j] = INF;
            }
            const int li = nb - blo[P];                                       // position of my block in the window
            const int a0 = nb * K;
            const int o = 1 - P;
            const bool act = a0 <= cHi[P] && a0 + K - 1 >= cLo[P];
            // D[a0-1][c-1] from the lane holding the block below mine
            int up = __shfl_sync(FULL, Dp[P][K - 1], (lane - 1) & 31);
            if (li == 0) up = INF;
            u32 w0[K];                                                        // swap source | count << 20 | base << 24 of my candidate rows
            auto swap_eval = [&](const int j, int &best, int &sb) {
                const u32 w = w0[j];
                const int cnt = (int)((w >> 20) & 15u);
                best = INF; sb = 0;
                if (cnt && pvalid[o]) {
                    const int s0 = (int)(w & 0xfffffu);
                    if ((unsigned)(s0 / K - pblo[o]) < 32u) best = rprev[o * W + (s0 & (W - 1))];
                    if (cnt > 1) {                                            // rare: insertion / adjacent deletions (:347, :376)
                        const int2 r = band_swap_multi(X.tab[P] + len[P] + 1 + X.tab[P][a0 + j], cnt, rprev + o * W, pblo[o], K, best);
                        best = r.x; sb = r.y;
                    }
                }
            };
            // ---- pass 1: thread-local chain ----
            int Dc[K];
            if (act) {
                int run = INF, upj = up;
#pragma unroll
                for (int j = 0; j < K; j++) {
                    const int a = a0 + j;
                    const bool cand = a >= cLo[P] && a <= cHi[P];
                    w0[j] = cand ? rows[a & (R - 1)].x : 0xff000000u;
                    const bool m = (int)(w0[j] >> 24) == tch;
                    int b;
                    if (a == 0 && c == 0) b = 0;                              // both origins start at 0 (:299-305)
                    else {
                        b = upj + (m ? 0 : 1);                                // diagonal (:324-332, :415-422)
                        b = min(b, Dp[P][j] + 1);                             // deletion (:406-413)
                        if (tok && m) {                                       // swap (:334-349, :363-378)
                            int best, sb;
                            swap_eval(j, best, sb);
                            b = min(b, best);
                        }
                    }
                    if (!cand) b = INF;
                    run = min(b, run + 1);                                    // insertion chain (:397-404)
                    if (run > INF) run = INF;
                    Dc[j] = run;
                    upj = Dp[P][j];
                }
            } else {
#pragma unroll
                for (int j = 0; j < K; j++) { Dc[j] = INF; w0[j] = 0xff000000u; }
            }
            // ---- min-plus prefix scan of G = D(last row) - (last row) over the blocks of the window ----
            int incl = (act && Dc[K - 1] < INF / 2) ? Dc[K - 1] - (a0 + K - 1) : INF;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int v = __shfl_sync(FULL, incl, (lane - d) & 31);
                if (li >= d) incl = min(incl, v);
            }
            int carry = __shfl_sync(FULL, incl, (lane - 1) & 31);
            if (li == 0) carry = INF;
            // ---- pass 2: final values and flags ----
            if (act) {
                u32 fw[(K + 3) / 4];
#pragma unroll
                for (int i = 0; i < (K + 3) / 4; i++) fw[i] = 0;
                int prevD = carry >= INF / 2 ? INF : carry + (a0 - 1);        // D[a0-1][c]
                int upj = up;
#pragma unroll
                for (int j = 0; j < K; j++) {
                    const int a = a0 + j;
                    int d = Dc[j];
                    if (carry < INF / 2) d = min(d, carry + a);
                    int f = 0;
                    if (a >= cLo[P] && a <= cHi[P] && d < INF / 2) {
                        if (a == 0 && c == 0) f = F_DIAG;
                        else {
                            const bool m = (int)(w0[j] >> 24) == tch;
                            if (a > 0 && c > 0 && upj + (m ? 0 : 1) == d) f |= F_DIAG;
                            if (a > 0 && prevD + 1 == d) f |= F_INS;
                            if (c > 0 && Dp[P][j] + 1 == d) f |= F_DEL;
                            if (tok && m) {
                                int best, sb;
                                swap_eval(j, best, sb);
                                if (best == d) f |= F_SWP | sb;
                            }
                        }
                        if (!unbounded && d <= tau) {
                            // kept: within the bound AND able to finish within it (wave_row_hulls: with rc truth bases
                            // left the remaining cost is at least max(0, rc - hhi, hlo - rc))
                            const uint4 rec = rows[a & (R - 1)];
                            const int rc = X.Lt - 1 - c, slack = tau - d;
                            if (rc - (int)rec.w <= slack && (int)rec.z - rc <= slack) {
                                // successors in the next column: the same row (deletion), the row above (diagonal) and my swap
                                // destination; from each of them the insertion chain climbs while the bound allows: + slack rows
                                nlo[P] = min(nlo[P], a); nhi[P] = max(nhi[P], a + 1 + slack);
                                const int fd = (int)rec.y;
                                if (fd >= 0) { nlo[o] = min(nlo[o], fd); nhi[o] = max(nhi[o], fd + slack); }
                            }
                        }
                    } else d = INF;
                    fw[j >> 2] |= (u32)f << ((j & 3) * 8);
                    upj = Dp[P][j];
                    Dp[P][j] = d;
                    prevD = d;
                }
                store_flags<K>(F + ((int64_t)c * 2 + P) * WS + (a0 & (W - 1)), fw);
            } else {
#pragma unroll
                for (int j = 0; j < K; j++) Dp[P][j] = INF;
            }
#pragma unroll
            for (int j = 0; j < K; j++) rcur[P * W + lane * K + j] = Dp[P][j];
        }
        // ---- the next column's candidates: successors of the kept cells (same row, row + 1, swap destinations) ----
#pragma unroll
        for (int P = 0; P < 2; P++) { pvalid[P] = has[P]; pblo[P] = blo[P]; }
        if (!unbounded) {
            nextLo[0] = __reduce_min_sync(FULL, nlo[0]); nextHi[0] = __reduce_max_sync(FULL, nhi[0]);
            nextLo[1] = __reduce_min_sync(FULL, nlo[1]); nextHi[1] = __reduce_max_sync(FULL, nhi[1]);
        }
        __syncwarp();                                                         // ring of this column visible to all lanes
    }
#ifdef VD_BAND_DEBUG
    if (lane == 0) printf("band_fwd K=%d idx=%d sc=%d ai=%d Lq=%d Lr=%d Lt=%d tau=%d failed=%d next Q[%d,%d] R[%d,%d]\n", K, idx, X.sc, X.ai, X.Lq, X.Lr, X.Lt, tau,
                          (int)failed, nextLo[0], nextHi[0], nextLo[1], nextHi[1]);
#endif
    if (failed) {
        if (lane == 0) {
            int nh = rung + 1;
            if (c_fail > 0) {                                                 // extrapolate the edit rate seen so far
                const long long est = (long long)(tau + 1) * X.Lt / c_fail * 9 / 8;      // + 1/8: a guess too low costs a whole extra sweep
                while (nh < N_RUNG - 1 && band_rung_tau(nh) < est) nh++;
            }
            if (last_rung) state[idx] = BAND_DENSE; else hint[idx] = nh;
        }
        return;
    }
    // ---- score and end plane (:390-391, :436-440): the last rows of the last column, through the ring ----
    const int *rl = ring + ((X.Lt - 1) & 1) * 2 * W;
    int dq = INF, dr = INF;
    if (pvalid[0] && (unsigned)((X.Lq - 1) / K - pblo[0]) < 32u) dq = rl[(X.Lq - 1) & (W - 1)];
    if (pvalid[1] && (unsigned)((X.Lr - 1) / K - pblo[1]) < 32u) dr = rl[W + ((X.Lr - 1) & (W - 1))];
    const int score = min(dq, dr);
    if (score <= tau && score < INF / 2) {
        if (lane == 0) {
            const int64_t oi = 4 * (int64_t)X.sc + X.ai;
            A.out.aln_score[oi] = score;
            A.out.aln_end_plane[oi] = (u8)(dq == score ? 0 : 1);
            state[idx] = K;
            atomicAdd(&wi->band_cells, (unsigned long long)visited);
            atomicAdd(&wi->band_rows, (unsigned long long)(X.Lq + X.Lr));
            atomicAdd(&wi->band_cols, (unsigned long long)X.Lt);
        }
    } else if (lane == 0) { if (last_rung) state[idx] = BAND_DENSE; else hint[idx] = rung + 1; }
}

// ------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(32 * BAND_WARPS, band_minb(K)) band_bwd_kernel(WaveArgs A, int n_items, const int *state) {
    VD_DYN_SHARED(smem_raw);
    typedef BandCfg<K> C;
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int W = C::W, R = C::R, NS = C::NS;
    constexpr int PFD = 6;                                                    // columns the flag prefetch runs ahead
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int idx = n_items - 1 - (blockIdx.x * BAND_WARPS + warp);
    if (idx < 0) return;
    if (state[idx] != K) return;
    const BandCtx X = band_ctx(A, A.items[idx]);
    const int len[2] = {X.Lq, X.Lr};
    const int64_t oi = 4 * (int64_t)X.sc + X.ai;
    const int end_plane = A.out.aln_end_plane[oi];
    u8 *sm = smem_raw + warp * C::SMEM;
    int *ring = (int *)sm;                                                    // [column parity][plane][W]: T * 256 + forward flags
    uint4 *srow = (uint4 *)(sm + 16 * W);                                     // [plane][R]: BwdRow records
    int4 *bandbuf = (int4 *)(sm + 16 * W + 32 * R);                           // [chunk parity][32]: band records of 32 columns
    unsigned long long *bars = (unsigned long long *)(sm + 16 * W + 32 * R + 1024);
    if (lane == 0) for (int i = 0; i < 2 * NS; i++) VD_MBAR_INIT(bars + i, 1);
    VD_MBAR_INIT_FENCE();
    __syncwarp();
    RowStream<NS, true> strm[2];
    strm[0].init(X.brow[0], srow, bars, X.Lq);
    strm[1].init(X.brow[1], srow + R, bars + NS, X.Lr);
    u8 *F = X.F;
    const int WS = 32 * X.kmax;
    u32 status = 0;

    int blk[2] = {-1, -1};
    int Tn[2][K];                                                             // T of column c+1, my rows
    u32 Fn[2][(K + 3) / 4];                                                   // forward flags of column c+1, my rows
#pragma unroll
    for (int P = 0; P < 2; P++) {
#pragma unroll
        for (int j = 0; j < K; j++) Tn[P][j] = -1;
#pragma unroll
        for (int i = 0; i < (K + 3) / 4; i++) Fn[P][i] = 0;
    }
    int nblo[2] = {0, 0};                                                     // first block of column c+1's window
    bool nvalid[2] = {false, false};
    // band records and truth bases in 32-column chunks, a chunk ahead
    const int cc0 = (X.Lt - 1) >> 5;
    {
        const int col = cc0 * 32 + lane;
        bandbuf[(cc0 & 1) * 32 + lane] = col < X.Lt ? X.band[col] : make_int4(1, 0, 1, 0);
    }
    int4 bnext = cc0 > 0 ? X.band[(cc0 - 1) * 32 + lane] : make_int4(1, 0, 1, 0);
    int tchunk = cc0 * 32 + lane < X.Lt ? X.tinfo[cc0 * 32 + lane] : 0, tchunk_next = cc0 > 0 ? X.tinfo[(cc0 - 1) * 32 + lane] : 0;
    __syncwarp();
    int tch_next = 0;
    for (int c = X.Lt - 1; c >= 0; c--) {
        const bool last = c == X.Lt - 1;
        const int cc = c >> 5;
        if ((c & 31) == 31 && !last) {                                        // entering chunk cc: its records were requested a chunk ago
            bandbuf[(cc & 1) * 32 + lane] = bnext;
            tchunk = tchunk_next;
            if (cc > 0) { bnext = X.band[(cc - 1) * 32 + lane]; tchunk_next = X.tinfo[(cc - 1) * 32 + lane]; }
            __syncwarp();
        }
        const int4 bd = bandbuf[(cc & 1) * 32 + (c & 31)];
        const int cLo[2] = {bd.x, bd.z}, cHi[2] = {bd.y, bd.w};
        const int *rnext = ring + ((c + 1) & 1) * 2 * W;
        int *rcur = ring + (c & 1) * 2 * W;
        bool has[2];
        int blo[2];
#pragma unroll
        for (int P = 0; P < 2; P++) { has[P] = cHi[P] >= cLo[P]; blo[P] = has[P] ? cLo[P] / K : nblo[P]; }
#ifndef VD_EMU
        if (c >= PFD) {                                                       // the flag bytes of my slots PFD columns ahead (the slot does not depend on the window)
            const u8 *pf = F + ((int64_t)(c - PFD) * 2) * WS + lane * K;
            asm volatile("prefetch.global.L1 [%0];" ::"l"(pf));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(pf + WS));
        }
#endif
#pragma unroll
        for (int P = 0; P < 2; P++) {
            if (!has[P]) {                                                    // nothing of this plane is on any path through column c
#pragma unroll
                for (int j = 0; j < K; j++) Tn[P][j] = -1;
#pragma unroll
                for (int i = 0; i < (K + 3) / 4; i++) Fn[P][i] = 0;
                continue;
            }
            strm[P].need(cLo[P], cHi[P], lane, C::PREF);
            const uint4 *rows = srow + P * R;
            const int o = 1 - P;
            const int nb = blo[P] + ((lane - blo[P]) & 31);
            // (row a0+K, column c+1) lives on the next lane if it held block nb+1 in column c+1
            int Tup = __shfl_sync(FULL, Tn[P][0], (lane + 1) & 31);
            int Fup = __shfl_sync(FULL, (int)(Fn[P][0] & 0xff), (lane + 1) & 31);
            const int bup = __shfl_sync(FULL, blk[P], (lane + 1) & 31);
            if (bup != nb + 1) { Tup = -1; Fup = 0; }
            if (nb != blk[P]) {
                blk[P] = nb;
#pragma unroll
                for (int j = 0; j < K; j++) Tn[P][j] = -1;
#pragma unroll
                for (int i = 0; i < (K + 3) / 4; i++) Fn[P][i] = 0;
            }
            const int li = nb - blo[P];
            const int a0 = nb * K;
            const bool act = a0 <= cHi[P] && a0 + K - 1 >= cLo[P];
            // forward flags of my rows in column c (only rows of the band hold valid flags) and their records
            u32 Fc[(K + 3) / 4];
#pragma unroll
            for (int i = 0; i < (K + 3) / 4; i++) Fc[i] = 0;
            u32 si[K], tw[K], chn[K];
            if (act) {
                load_flags<K>(F + ((int64_t)c * 2 + P) * WS + (a0 & (W - 1)), Fc);
#pragma unroll
                for (int j = 0; j < K; j++) {
                    const int a = a0 + j;
                    if (a < cLo[P] || a > cHi[P]) { Fc[j >> 2] &= ~(0xffu << ((j & 3) * 8)); si[j] = 0; tw[j] = 0; chn[j] = 0xff; }
                    else { const uint4 rec = rows[a & (R - 1)]; si[j] = rec.x; tw[j] = rec.y; chn[j] = rec.z; }
                }
            } else {
#pragma unroll
                for (int j = 0; j < K; j++) { si[j] = 0; tw[j] = 0; chn[j] = 0xff; }
            }
            // flag of row a0+K in column c (the link into my last row): the next lane's first row
            int FcUp = __shfl_sync(FULL, (int)(Fc[0] & 0xff), (lane + 1) & 31);
            if (li == 31) FcUp = 0;
            // ---- pass 1: candidates from column c+1 ----
            int B[K];
            int swv[K];
#pragma unroll
            for (int j = K - 1; j >= 0; j--) {
                const int a = a0 + j;
                int b = -1;
                swv[j] = -1;
                if (act && a >= cLo[P] && a <= cHi[P]) {
                    if (last && P == end_plane && a == len[P] - 1) b = 0;                        // :543-545
                    if (!last) {
                        const int Tx = (j == K - 1) ? Tup : Tn[P][j == K - 1 ? j : j + 1];
                        const int Fx = (j == K - 1) ? Fup : byte_of(Fn[P], j == K - 1 ? j : j + 1);
                        const int tpn = (int)((tw[j] >> 25) & 1);
                        if (a + 1 < len[P] && Tx >= 0 && (Fx & F_DIAG)) b = max(b, Tx + tpn);      // :556-595, :692-731
                        if (Tn[P][j] >= 0 && (byte_of(Fn[P], j) & F_DEL)) b = max(b, Tn[P][j]);  // :774-804
                        const u32 s = si[j];
                        if ((s & 1) && nvalid[o]) {                                               // :598-679
                            const int d = (int)(s >> 8);
                            if ((unsigned)(d / K - nblo[o]) < 32u) {
                                const int tf = rnext[o * W + (d & (W - 1))];
                                const int T2 = tf >> 8, F2 = tf & 0xff;
                                if (T2 >= 0 && (F2 & F_SWP) && (F2 >> F_K_SHIFT) == (int)((s >> 1) & 7)) {
                                    swv[j] = T2 + (int)((s >> 4) & 1);
                                    b = max(b, swv[j]);
                                    if (F2 & F_TIE) status |= VD_ST_TIE;
                                }
                            }
                        }
                    }
                }
                B[j] = b;
            }
            // link(j): row a0+j+1 of column c carries F_INS, i.e. (a0+j) is reached from it (:734-771)
            unsigned links = 0;
#pragma unroll
            for (int j = 0; j < K; j++) {
                const int fb = (j == K - 1) ? FcUp : byte_of(Fc, j == K - 1 ? j : j + 1);
                if ((fb & F_INS) && a0 + j + 1 < len[P]) links |= 1u << j;
            }
            int head = NEG;                                                   // U = T - S of row a0 without anything arriving from above
#pragma unroll
            for (int j = K - 1; j >= 0; j--) {
                const int bu = B[j] >= 0 ? B[j] - (int)(tw[j] & 0xffffffu) : NEG;
                head = (j < K - 1 && ((links >> j) & 1)) ? max(bu, head) : bu;
            }
            // ---- link-segmented suffix max of the block heads over the window ----
            const bool allopen = links == ((1u << K) - 1u);
            const unsigned am = __ballot_sync(FULL, allopen);
            const int rot = blo[P] & 31;                                      // lane of the window's first block
            const unsigned al = rot ? ((am >> rot) | (am << (32 - rot))) : am;
            const unsigned nm = ~(al >> li);
            const int runl = nm ? __ffs(nm) - 1 : 32;
            int val = head;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int v = __shfl_sync(FULL, val, (lane + d) & 31);
                if (runl >= d && li + d < 32) val = max(val, v);
            }
            int Xin = __shfl_sync(FULL, val, (lane + 1) & 31);                // final U of row a0+K
            if (li == 31) Xin = NEG;
            // ---- pass 2: final T, path flags (in place) ----
            u32 pfw[(K + 3) / 4];
#pragma unroll
            for (int i = 0; i < (K + 3) / 4; i++) pfw[i] = 0;
            int Tc[K];
            {
                int uin = Xin;                                                // final U of row a0+j+1
#pragma unroll
                for (int j = K - 1; j >= 0; j--) {
                    const int a = a0 + j;
                    const bool lk = (links >> j) & 1;
                    const int S = (int)(tw[j] & 0xffffffu);
                    const int tpn = (int)((tw[j] >> 25) & 1);
                    const int bu = B[j] >= 0 ? B[j] - S : NEG;
                    const int u = (lk && uin > NEG / 2) ? max(bu, uin) : bu;
                    const int Tv = u > NEG / 2 ? u + S : -1;
                    const int tabove = (lk && uin > NEG / 2) ? uin + (S - tpn) : -1;      // T of row a+1: S(a+1) = S(a) - tp(a+1)
                    int pf = 0;
                    if (Tv >= 0) {
                        if (last && P == end_plane && a == len[P] - 1 && Tv == 0) pf |= PTR_MAT;   // :543
                        if (!last) {
                            const int Tx = (j == K - 1) ? Tup : Tn[P][j == K - 1 ? j : j + 1];
                            const int Fx = (j == K - 1) ? Fup : byte_of(Fn[P], j == K - 1 ? j : j + 1);
                            if (a + 1 < len[P] && Tx >= 0 && (Fx & F_DIAG) && Tx + tpn == Tv)
                                pf |= ((int)chn[j] == tch_next) ? PTR_MAT : PTR_SUB;
                            if (Tn[P][j] >= 0 && (byte_of(Fn[P], j) & F_DEL) && Tn[P][j] == Tv) pf |= PTR_DEL;
                            if (swv[j] >= 0 && swv[j] == Tv) pf |= PTR_SWP;
                        }
                        if (tabove >= 0 && tabove + tpn == Tv) pf |= PTR_INS;
                    }
                    pfw[j >> 2] |= (u32)pf << ((j & 3) * 8);
                    Tc[j] = Tv;
                    uin = u;
                }
            }
            if (act) store_flags<K>(F + ((int64_t)c * 2 + P) * WS + (a0 & (W - 1)), pfw);
            // publish column c
#pragma unroll
            for (int j = 0; j < K; j++) {
                Tn[P][j] = Tc[j];
                rcur[P * W + lane * K + j] = Tc[j] * 256 + byte_of(Fc, j);
            }
#pragma unroll
            for (int i = 0; i < (K + 3) / 4; i++) Fn[P][i] = Fc[i];
        }
#pragma unroll
        for (int P = 0; P < 2; P++) { nvalid[P] = has[P]; nblo[P] = blo[P]; }
        tch_next = __shfl_sync(FULL, tchunk, c & 31) & 0x7f;
        __syncwarp();
    }
    // origin plane (:811-814): QUERY if its origin was reached
    status = __reduce_or_sync(FULL, status);
    if (lane == 0) {
        const int *r0 = ring;                                                 // column 0
        int t00 = -1;
        if (nvalid[0] && nblo[0] == 0) t00 = r0[0] >> 8;
        A.out.aln_beg_plane[oi] = (u8)(t00 >= 0 ? 0 : 1);
        if (status) atomicOr(&A.out.status[oi], status);
    }
}

// path flags of the banded layout
struct PFBand {
    const u8 *F; int WS, WM, Lt;                                             // row stride per plane, window mask
    __device__ __forceinline__ int get(int hi, int qri, int ti) const {
        return F[((int64_t)ti * 2 + hi) * WS + (qri & WM)];
    }
    static constexpr int AHEAD = 8;
    __device__ __forceinline__ void prefetch(int hi, int qri, int ti) const {
#ifndef VD_EMU
        if (ti + AHEAD < Lt) {
            const u8 *p = F + ((int64_t)(ti + AHEAD) * 2 + hi) * WS + ((qri + AHEAD) & WM);
            asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
        }
#endif
    }
};

// walk + credit of the alignments the band kernels solved: wpw alignments per warp, on lanes 32/wpw apart (one per warp keeps
// walks of very different lengths from holding each other up but leaves 31 lanes idle: with tens of thousands of
// alignments the kernel is then bound by the number of resident warps)
__global__ void band_walk_kernel(WaveArgs A, int n_items, int *state, int only_k, int wpw) {
    const int lane = threadIdx.x & 31, lpi = 32 / wpw;
    const int g = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * wpw + lane / lpi;
    if (g >= n_items || (lane % lpi)) return;
    const int idx = n_items - 1 - g;
    const int K = state[idx];
    if (K <= 0 || K != only_k) return;                       // not solved, another rung's, or walked already (BAND_DONE)
    const int item = A.items[idx];
    const int e = item >> 2, ai = item & 3;
    const int i = A.i0 + e;
    const int sc = A.list[i];
    const ScPlan p = A.plan[sc];
    const WaveSlab W = make_wave_slab(p);
    u8 *base = A.slab + (A.offs[i] - A.offs[A.i0]);
    const int qh = ai >> 1, th = 2 + (ai & 1);
    SlabHap HQ(base + W.base.hap[qh], p.len[qh], p.lr), HT(base + W.base.hap[th], p.len[th], p.lr);
    SlabQm M(base + W.base.qm[qh], p.len[qh], p.lr);
    Hap<int> q{p.len[qh], HQ.str, HQ.flg, HQ.ptr, HQ.ins};
    Hap<int> t{p.len[th], HT.str, HT.flg, HT.ptr, HT.ins};
    QMaps<int> qm{M.rptr, M.rflg, M.toQ, M.toR};
    const u8 *rseq = A.in.rplane_seq + A.in.ref_off[sc];
    const WaveAln wa = wave_aln(q.len, p.lr, t.len);
    u8 *ab = base + W.aln[ai];
    GMem mem{ab + wa.oWalk};
    const AlnLayout<int64_t> L = wave_walk_layout(q.len, p.lr, t.len);
    PFBand pfr{ab + wa.oF, 32 * wa.kmax, 32 * K - 1, t.len};
    u32 status = A.out.status[4 * (int64_t)sc + ai];
    const int beg_plane = A.out.aln_beg_plane[4 * (int64_t)sc + ai];
    const int end_plane = A.out.aln_end_plane[4 * (int64_t)sc + ai];
    walk_credit<GMem, 4, int>(mem, L, pfr, q, qm, t, rseq, p.lr, beg_plane, end_plane, A.in, A.out, sc, ai, status);
    A.out.status[4 * (int64_t)sc + ai] = status;
    state[idx] = K | BAND_DONE;
}

template <int K> inline void band_launch(cudaStream_t st, const WaveArgs &A, int n_items, int *state, const int *lbound, int *hint, bool fwd, int rung, int exact, WaveItems *wi) {
    if (n_items <= 0) return;
    const int nb = (n_items + BAND_WARPS - 1) / BAND_WARPS, sm = BAND_WARPS * BandCfg<K>::SMEM;
    if (fwd) VD_LAUNCH(band_fwd_kernel<K>, nb, 32 * BAND_WARPS, sm, st, A, n_items, state, lbound, hint, rung, exact, wi);
    else VD_LAUNCH(band_bwd_kernel<K>, nb, 32 * BAND_WARPS, sm, st, A, n_items, (const int *)state);
}
inline void band_launch_rung(cudaStream_t st, int rung, const WaveArgs &A, int n_items, int *state, const int *lbound, int *hint, bool fwd, int exact, WaveItems *wi) {
    switch (band_rung_k(rung)) {
        case 1: band_launch<1>(st, A, n_items, state, lbound, hint, fwd, rung, exact, wi); break;
        case 2: band_launch<2>(st, A, n_items, state, lbound, hint, fwd, rung, exact, wi); break;
        case 4: band_launch<4>(st, A, n_items, state, lbound, hint, fwd, rung, exact, wi); break;
        case 8: band_launch<8>(st, A, n_items, state, lbound, hint, fwd, rung, exact, wi); break;
        case 16: band_launch<16>(st, A, n_items, state, lbound, hint, fwd, rung, exact, wi); break;
    }
}
template <int K> inline void band_configure_one() {
    cudaFuncSetAttribute(band_fwd_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, BAND_WARPS * BandCfg<K>::SMEM);
    cudaFuncSetAttribute(band_bwd_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, BAND_WARPS * BandCfg<K>::SMEM);
}
inline void band_configure() {
    band_configure_one<1>(); band_configure_one<2>(); band_configure_one<4>(); band_configure_one<8>(); band_configure_one<16>();
}

}  // namespace vd
