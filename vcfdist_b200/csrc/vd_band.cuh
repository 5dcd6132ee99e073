// Banded warp-per-alignment kernels of the long path: forward (calc_prec_recall_aln, src/dist.cpp:251-443)
// and backward (calc_prec_recall_path, :486-834) column sweeps in which ONE WARP owns an alignment and no
// block barrier is ever taken.
//
// The reference explores the two-plane graph by increasing score and stops when the end is reached
// (:309-443), so it only ever touches cells whose distance is at most the final score.  A structural
// variant that truth and query both carry gives matrices of thousands x thousands with a score of a few
// dozen: the cells that matter are a narrow band around one diagonal per plane.  These kernels keep, per
// plane, a WINDOW of 32*K consecutive rows that slides with the band (Ukkonen's cut-off at a score bound
// tau; exact whenever the final score is <= tau):
//
//   * rows are dealt in blocks of K consecutive rows, block b always to lane b % 32; the window of a column
//     is the 32 blocks starting at the block of the lowest candidate row.  A lane whose block leaves the
//     window picks up the block 32 further on and reloads the K packed row records (RowRec, 16 bytes a
//     row: swap source, swap destination, query-variant counts, bases) - once every K columns in a
//     diagonal region;
//   * the column step of a lane is a serial chain over its K rows in registers (diagonal, deletion, swap
//     candidate, the in-column insertion chain), then ONE warp scan hands the chain across lanes
//     (min-plus prefix scan forward, link-segmented suffix max of T - S backward), then a second pass over
//     the K rows finalises values and flags.  Swap edges read the other plane's previous column from a
//     ring in shared memory (32*K entries per plane and column parity);
//   * flags are stored BANDED: F[column][plane][row mod 32K], one byte per cell of the window - 64*K bytes
//     per column instead of (Lq+Lr) bytes, so an alignment of 10 k x 10 k needs 2.5-10 MB instead of 200 MB
//     and the walk's dependent flag loads stay within a few cache lines per column.
//
// Rungs: K = 4 (tau 40), K = 8 (tau 80), K = 16 (tau 160).  A rung whose planes both fit the window
// entirely (<= 32*K rows each) runs unbounded: that is the dense sweep of the thin-but-wide matrices (no
// variant on the query side against a long insertion on the truth side).  An alignment is tried rung by
// rung, skipping rungs below a lower bound of its score; what no rung solves (score > 160, or a band wider
// than the window) goes to the dense block-per-alignment kernels of vd_wave.cuh.
#pragma once
#include "vd_wave.cuh"

namespace vd {

constexpr int BAND_WARPS = 4;                                  // alignments per block (one warp each)
__host__ __device__ inline int band_tau(int K) { return 10 * K; }
// per-item state shared by the rungs
constexpr int BAND_PENDING = 0;       // not solved yet
constexpr int BAND_DENSE = -1;        // given up: dense kernels
// > 0: K of the rung that solved it

struct BandCtx {
    int sc, ai, Lq, Lr, Lt;
    const u8 *tinfo;
    const RowRec *row[2];             // QUERY plane rows, REF plane rows
    const int2 *hull[2];              // row-step hulls (wave_row_hulls)
    const int *optr[2];               // row -> row of the other plane (qptr, rptr)
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
    SlabHap HQ(base + W.base.hap[qh], x.Lq, x.Lr);
    SlabQm M(base + W.base.qm[qh], x.Lq, x.Lr);
    WaveHapQ T(base + W.hq[qh], x.Lq, x.Lr);
    x.tinfo = base + W.ht[th];
    x.row[0] = T.rowQ; x.row[1] = T.rowR;
    x.hull[0] = T.hullQ; x.hull[1] = T.hullR;
    x.optr[0] = HQ.ptr; x.optr[1] = M.rptr;
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

// static data of the K rows of block blk for the forward sweep: swap sources (RowRec::swi) and bases
// and, into the warp's shared-memory cache, their row-step hulls turned into the first / last column in which
// the row can still finish within the bound at distance 0: cmin = Lt-1-hi, cmax = Lt-1-lo
template <int K> __device__ __forceinline__ void band_load_rows(const RowRec *rows, const int2 *hull, int len, int blk, int Lt,
                                                                u32 (&swi)[K], u32 (&chw)[K], int2 *hcache) {
#pragma unroll
    for (int j = 0; j < K; j++) {
        const int a = blk * K + j;
        if (a < len) {
            const uint4 r = *(const uint4 *)(rows + a);
            swi[j] = r.x; chw[j] = r.w;
            const int2 hb = hull[a];
            hcache[j] = make_int2(hb.y >= INF / 2 ? -INF : Lt - 1 - hb.y, Lt - 1 - hb.x);
        } else { swi[j] = 0; chw[j] = 0xffffu; }
    }
}

// ------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(32 * BAND_WARPS) band_fwd_kernel(WaveArgs A, int n_items, int *state, const int *lbound) {
    VD_DYN_SHARED(smem_raw);
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int W = 32 * K;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int idx = n_items - 1 - (blockIdx.x * BAND_WARPS + warp);          // biggest shape classes first
    if (idx < 0) return;
    if (state[idx] != BAND_PENDING) return;
    const BandCtx X = band_ctx(A, A.items[idx]);
    if (K > X.kmax) { if (lane == 0) state[idx] = BAND_DENSE; return; }
    const int len[2] = {X.Lq, X.Lr};
    const bool unbounded = (X.Lq + K - 1) / K <= 32 && (X.Lr + K - 1) / K <= 32;
    const int tau = unbounded ? INF : band_tau(K);
    if (!unbounded && lbound[idx] > tau) {                                    // this rung cannot succeed
        if (K == 16 && lane == 0) state[idx] = BAND_DENSE;
        return;
    }
    int *ring = (int *)smem_raw + warp * (8 * W);                             // [column parity][plane][W]: D of the previous column
    int2 *hcache = (int2 *)(ring + 4 * W);                                    // [plane][W]: column range in which a row can still finish
    u8 *F = X.F;
    const int WS = 32 * X.kmax;                                               // row stride of the flag storage (sized for kmax)

    int blk[2] = {lane, lane};                                                // the window starts at block 0
    int Dp[2][K];
    u32 swi[2][K], chw[2][K];
#pragma unroll
    for (int P = 0; P < 2; P++) {
        band_load_rows<K>(X.row[P], X.hull[P], len[P], blk[P], X.Lt, swi[P], chw[P], hcache + P * W + lane * K);
#pragma unroll
        for (int j = 0; j < K; j++) Dp[P][j] = INF;
    }
    int liveLo[2] = {INF, INF}, liveHi[2] = {-1, -1}, dmin = 0;
    int pblo[2] = {0, 0};                                                     // first block of the previous column's window
    bool pvalid[2] = {false, false};                                          // the ring holds the plane's previous column
    bool failed = false;
    int tnext = X.tinfo[0];
    for (int c = 0; c < X.Lt; c++) {
        const int tin = tnext;
        if (c + 1 < X.Lt) tnext = X.tinfo[c + 1];
        const int tch = tin & 0x7f;
        const bool tok = tin & 0x80;
        // ---- candidate rows of both planes (plane-local, inclusive) ----
        int cLo[2], cHi[2];
        if (c == 0 || unbounded) {
            cLo[0] = cLo[1] = 0;
            cHi[0] = unbounded ? X.Lq - 1 : min(tau, X.Lq - 1);
            cHi[1] = unbounded ? X.Lr - 1 : min(tau, X.Lr - 1);
        } else {
            const bool hq = liveHi[0] >= liveLo[0], hr = liveHi[1] >= liveLo[1];
            if (!hq && !hr) { failed = true; break; }                         // nothing within tau is left
            int lo0 = INF, hi0 = -1, lo1 = INF, hi1 = -1;
            if (hq) { lo0 = liveLo[0]; hi0 = liveHi[0] + 1; lo1 = X.optr[0][liveLo[0]] + 1; hi1 = X.optr[0][liveHi[0]] + 1; }
            if (hr) {
                lo1 = min(lo1, liveLo[1]); hi1 = max(hi1, liveHi[1] + 1);
                lo0 = min(lo0, X.optr[1][liveLo[1]] + 1); hi0 = max(hi0, X.optr[1][liveHi[1]] + 1);
            }
            const int ext = tau - dmin;                                       // the in-column insertion chain reaches this far
            cLo[0] = max(lo0, 0); cHi[0] = min(hi0 + ext, X.Lq - 1);
            cLo[1] = max(lo1, 0); cHi[1] = min(hi1 + ext, X.Lr - 1);
        }
        bool has[2];
        int blo[2];
#pragma unroll
        for (int P = 0; P < 2; P++) {
            has[P] = cHi[P] >= cLo[P];
            blo[P] = has[P] ? cLo[P] / K : pblo[P];
            if (has[P] && cHi[P] / K - blo[P] > 31) failed = true;            // band wider than the window
        }
        if (failed) break;
        if (lane == 0) X.band[c] = make_int4(has[0] ? cLo[0] : 1, has[0] ? cHi[0] : 0, has[1] ? cLo[1] : 1, has[1] ? cHi[1] : 0);
        const int *rprev = ring + ((c + 1) & 1) * 2 * W, *rcur_c = ring + (c & 1) * 2 * W;
        int *rcur = ring + (c & 1) * 2 * W;
        (void)rcur_c;
        int nlo[2] = {INF, INF}, nhi[2] = {-1, -1}, ndmin = INF;
#pragma unroll
        for (int P = 0; P < 2; P++) {
            if (!has[P]) {                                                    // plane without candidates: nothing live in this column
                if (pvalid[P] || c == 0) {
#pragma unroll
                    for (int j = 0; j < K; j++) Dp[P][j] = INF;
                }
                continue;
            }
            // ---- my block in this column's window ----
            const int nb = blo[P] + ((lane - blo[P]) & 31);
            if (nb != blk[P]) {
                blk[P] = nb;
                band_load_rows<K>(X.row[P], X.hull[P], len[P], nb, X.Lt, swi[P], chw[P], hcache + P * W + lane * K);
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
            auto swap_eval = [&](const int j, int &best, int &sb) {
                const u32 w = swi[P][j];
                const int cnt = (int)(w >> 24);
                best = INF; sb = 0;
                if (cnt && pvalid[o]) {
                    const int s0 = (int)(w & 0xffffffu);
                    const unsigned bo = (unsigned)(s0 / K - pblo[o]);
                    if (bo < 32u) best = rprev[o * W + (s0 & (W - 1))];
                    if (cnt > 1) {                                            // rare: insertion / adjacent deletions (:347, :376)
                        const int *tab = X.tab[P];
                        const int k0 = tab[a0 + j];
                        for (int k = 1; k < cnt; k++) {
                            const int s = tab[len[P] + 1 + k0 + k];
                            int v = INF;
                            if ((unsigned)(s / K - pblo[o]) < 32u) v = rprev[o * W + (s & (W - 1))];
                            if (v < best) { best = v; sb = k << F_K_SHIFT; }
                            else if (v == best && v < INF / 2) sb = (k << F_K_SHIFT) | F_TIE;   // keep the larger row
                        }
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
                    const bool m = (int)(chw[P][j] & 0xff) == tch;
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
                    if (a < cLo[P] || a > cHi[P]) b = INF;
                    run = min(b, run + 1);                                    // insertion chain (:397-404)
                    if (run > INF) run = INF;
                    Dc[j] = run;
                    upj = Dp[P][j];
                }
            } else {
#pragma unroll
                for (int j = 0; j < K; j++) Dc[j] = INF;
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
            u32 fw[(K + 3) / 4];
#pragma unroll
            for (int i = 0; i < (K + 3) / 4; i++) fw[i] = 0;
            if (act) {
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
                            const bool m = (int)(chw[P][j] & 0xff) == tch;
                            if (a > 0 && c > 0 && upj + (m ? 0 : 1) == d) f |= F_DIAG;
                            if (a > 0 && prevD + 1 == d) f |= F_INS;
                            if (c > 0 && Dp[P][j] + 1 == d) f |= F_DEL;
                            if (tok && m) {
                                int best, sb;
                                swap_eval(j, best, sb);
                                if (best == d) f |= F_SWP | sb;
                            }
                        }
                        if (d <= tau) {
                            // within the bound AND able to finish within it (remaining cost >= distance of c to the
                            // row's column range, see wave_row_hulls)
                            const int2 cr = hcache[P * W + lane * K + j];
                            if (unbounded || (c >= cr.x - (tau - d) && c <= cr.y + (tau - d))) { nlo[P] = min(nlo[P], a); nhi[P] = a; ndmin = min(ndmin, d); }
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
        // ---- rows within tau: the next column's candidates come from them ----
#pragma unroll
        for (int P = 0; P < 2; P++) {
            pvalid[P] = has[P];
            pblo[P] = blo[P];
            if (!unbounded) {
                liveLo[P] = has[P] ? __reduce_min_sync(FULL, nlo[P]) : INF;
                liveHi[P] = has[P] ? __reduce_max_sync(FULL, nhi[P]) : -1;
            }
        }
        if (!unbounded) dmin = __reduce_min_sync(FULL, ndmin);
        __syncwarp();                                                         // ring of this column visible to all lanes
    }
#ifdef VD_BAND_DEBUG
    if (lane == 0) printf("band_fwd K=%d idx=%d sc=%d ai=%d Lq=%d Lr=%d Lt=%d tau=%d failed=%d live Q[%d,%d] R[%d,%d] dmin=%d\n", K, idx, X.sc, X.ai, X.Lq, X.Lr, X.Lt, tau,
                          (int)failed, liveLo[0], liveHi[0], liveLo[1], liveHi[1], dmin);
#endif
    if (failed) {
        if (K == 16 && lane == 0) state[idx] = BAND_DENSE;
        return;
    }
    // ---- score and end plane (:390-391, :436-440): the last rows of the last column, through the ring ----
    const int *rl = ring + ((X.Lt - 1) & 1) * 2 * W;
    int dq = INF, dr = INF;
    if (pvalid[0] && (unsigned)((X.Lq - 1) / K - pblo[0]) < 32u) dq = rl[(X.Lq - 1) & (W - 1)];
    if (pvalid[1] && (unsigned)((X.Lr - 1) / K - pblo[1]) < 32u) dr = rl[W + ((X.Lr - 1) & (W - 1))];
    const int score = min(dq, dr);
#ifdef VD_BAND_DEBUG
    if (lane == 0) printf("band_fwd K=%d idx=%d score=%d (dq %d dr %d)\n", K, idx, score, dq, dr);
#endif
    if (score <= tau && score < INF / 2) {
        if (lane == 0) {
            const int64_t oi = 4 * (int64_t)X.sc + X.ai;
            A.out.aln_score[oi] = score;
            A.out.aln_end_plane[oi] = (u8)(dq == score ? 0 : 1);
            state[idx] = K;
        }
    } else if (K == 16 && lane == 0) state[idx] = BAND_DENSE;
}

// ------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(32 * BAND_WARPS) band_bwd_kernel(WaveArgs A, int n_items, const int *state) {
    VD_DYN_SHARED(smem_raw);
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int W = 32 * K;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int idx = n_items - 1 - (blockIdx.x * BAND_WARPS + warp);
    if (idx < 0) return;
    if (state[idx] != K) return;
    const BandCtx X = band_ctx(A, A.items[idx]);
    const int len[2] = {X.Lq, X.Lr};
    const int64_t oi = 4 * (int64_t)X.sc + X.ai;
    const int end_plane = A.out.aln_end_plane[oi];
    int *ring = (int *)smem_raw + warp * (4 * W);                             // [column parity][plane][W]: (T << 8) | forward flags
    u8 *F = X.F;
    const int WS = 32 * X.kmax;
    u32 status = 0;

    int blk[2] = {-1, -1};
    int Tn[2][K];                                                             // T of column c+1, my rows
    u32 Fn[2][(K + 3) / 4];                                                   // forward flags of column c+1, my rows
    u32 si[2][K], tw[2][K];                                                   // my rows as swap sources; tps | tp(a) << 24 | tp(a+1) << 25
    u32 chn[2][(K + 3) / 4];                                                  // bases of rows a+1
#pragma unroll
    for (int P = 0; P < 2; P++) {
#pragma unroll
        for (int j = 0; j < K; j++) { Tn[P][j] = -1; si[P][j] = 0; tw[P][j] = 0; }
#pragma unroll
        for (int i = 0; i < (K + 3) / 4; i++) { Fn[P][i] = 0; chn[P][i] = 0xffffffffu; }
    }
    int nblo[2] = {0, 0};                                                     // first block of column c+1's window
    bool nvalid[2] = {false, false};
    int4 bd = X.band[X.Lt - 1];
    int tch_next = 0;
    for (int c = X.Lt - 1; c >= 0; c--) {
        const bool last = c == X.Lt - 1;
        const int cLo[2] = {bd.x, bd.z}, cHi[2] = {bd.y, bd.w};
        if (c > 0) bd = X.band[c - 1];
        const int *rnext = ring + ((c + 1) & 1) * 2 * W;
        int *rcur = ring + (c & 1) * 2 * W;
        bool has[2];
        int blo[2];
#pragma unroll
        for (int P = 0; P < 2; P++) { has[P] = cHi[P] >= cLo[P]; blo[P] = has[P] ? cLo[P] / K : nblo[P]; }
#pragma unroll
        for (int P = 0; P < 2; P++) {
            if (!has[P]) {                                                    // nothing of this plane is on any path through column c
#pragma unroll
                for (int j = 0; j < K; j++) Tn[P][j] = -1;
#pragma unroll
                for (int i = 0; i < (K + 3) / 4; i++) Fn[P][i] = 0;
                continue;
            }
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
                for (int j = 0; j < K; j++) {
                    const int a = nb * K + j;
                    u32 cw = 0xff;
                    if (a < len[P]) {
                        const uint4 r = *(const uint4 *)(X.row[P] + a);
                        si[P][j] = r.y; tw[P][j] = r.z; cw = (r.w >> 8) & 0xff;
                    } else { si[P][j] = 0; tw[P][j] = 0; }
                    chn[P][j >> 2] = (chn[P][j >> 2] & ~(0xffu << ((j & 3) * 8))) | (cw << ((j & 3) * 8));
                    Tn[P][j] = -1;
                }
#pragma unroll
                for (int i = 0; i < (K + 3) / 4; i++) Fn[P][i] = 0;
            }
            const int li = nb - blo[P];
            const int a0 = nb * K;
            const bool act = a0 <= cHi[P] && a0 + K - 1 >= cLo[P];
            // forward flags of my rows in column c (only rows of the band hold valid flags)
            u32 Fc[(K + 3) / 4];
#pragma unroll
            for (int i = 0; i < (K + 3) / 4; i++) Fc[i] = 0;
            if (act) {
                load_flags<K>(F + ((int64_t)c * 2 + P) * WS + (a0 & (W - 1)), Fc);
#pragma unroll
                for (int j = 0; j < K; j++)
                    if (a0 + j < cLo[P] || a0 + j > cHi[P]) Fc[j >> 2] &= ~(0xffu << ((j & 3) * 8));
            }
            // flag of row a0+K in column c (the link into my last row): the next lane's first row
            int FcUp = __shfl_sync(FULL, (int)(Fc[0] & 0xff), (lane + 1) & 31);
            if (li == 31) FcUp = 0;
            // ---- pass 1: candidates from column c+1, local chain in potential form U = T - S ----
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
                        const int tpn = (int)((tw[P][j] >> 25) & 1);
                        if (a + 1 < len[P] && Tx >= 0 && (Fx & F_DIAG)) b = max(b, Tx + tpn);      // :556-595, :692-731
                        if (Tn[P][j] >= 0 && (byte_of(Fn[P], j) & F_DEL)) b = max(b, Tn[P][j]);  // :774-804
                        const u32 s = si[P][j];
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
            int head = NEG;                                                   // U of row a0 without anything arriving from above
#pragma unroll
            for (int j = K - 1; j >= 0; j--) {
                const int bu = B[j] >= 0 ? B[j] - (int)(tw[P][j] & 0xffffffu) : NEG;
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
                    const int S = (int)(tw[P][j] & 0xffffffu);
                    const int tpn = (int)((tw[P][j] >> 25) & 1);
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
                                pf |= (byte_of(chn[P], j) == tch_next) ? PTR_MAT : PTR_SUB;
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
        tch_next = X.tinfo[c] & 0x7f;
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

// walk + credit of the alignments the band kernels solved: one alignment per warp (lane 0 walks)
__global__ void band_walk_kernel(WaveArgs A, int n_items, const int *state, int only_k) {
    const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (g >= n_items || (threadIdx.x & 31)) return;
    const int idx = n_items - 1 - g;
    const int K = state[idx];
    if (K <= 0 || K != only_k) return;
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
}

// items the band kernels gave up on (or never tried): input of the dense phase
__global__ void band_count_dense_kernel(const int *state, int n_items, int *n_dense) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool d = i < n_items && state[i] <= 0;
    const unsigned m = __ballot_sync(0xffffffffu, d);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(n_dense, __popc(m));
}

template <int K> inline void band_launch(cudaStream_t st, const WaveArgs &A, int n_items, int *state, const int *lbound, bool fwd) {
    if (n_items <= 0) return;
    const int nb = (n_items + BAND_WARPS - 1) / BAND_WARPS, sm = BAND_WARPS * 4 * 32 * K * 4;
    if (fwd) VD_LAUNCH(band_fwd_kernel<K>, nb, 32 * BAND_WARPS, 2 * sm, st, A, n_items, state, lbound);
    else VD_LAUNCH(band_bwd_kernel<K>, nb, 32 * BAND_WARPS, sm, st, A, n_items, (const int *)state);
}
inline void band_configure() {
    cudaFuncSetAttribute(band_fwd_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * BAND_WARPS * 4 * 32 * 16 * 4);
    cudaFuncSetAttribute(band_bwd_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, BAND_WARPS * 4 * 32 * 16 * 4);
}

}  // namespace vd
