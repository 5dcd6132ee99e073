// Warp-per-supercluster kernel for mid-size superclusters: too big for the thread-per-alignment
// classes of small_kernel, small enough that all four alignments' flag matrices fit in shared
// memory (up to 128 rows over both planes).  Everything stays on chip, as in small_kernel; HBM traffic
// is the compact batch in and the result records out.
//
//   phase 1  lanes 0..3 expand the four haplotypes (generate_ptrs_strs, src/dist.cpp:145-242)
//            and build the swap-source tables; the other lanes stage the REF-plane string
//   phase 2  the four alignments one after the other, lanes across rows (row r lives on lane
//            r % 32, register slot r / 32; both planes back to back, QUERY rows first):
//            forward column sweep (calc_prec_recall_aln, :251-443) with the previous column in
//            registers — neighbour and swap-source reads are warp shuffles, the insertion chain
//            is a plane-segmented min-plus prefix scan; one flag byte per cell to shared memory;
//            backward sweep (calc_prec_recall_path, :486-834) in gather form, the in-column
//            insertion chain as a link-segmented suffix max of T - S (S = suffix count of
//            query-variant rows, which turns the weighted chain into a plain max); the path
//            flags overwrite the forward flags in place
//   phase 3  lanes 0..3 walk their alignment and assign credit (walk_credit of vd_scalar.cuh)
#pragma once
#include "vd_scalar.cuh"

namespace vd {

constexpr u32 ST_BAD = 0x0800u;        // VD_ST_ERR_BADINPUT
constexpr int WSC_TPB = 128;           // 4 warps = 4 superclusters per block
constexpr int WSC_MAXSLOT = 4;         // register slots of 32 rows: both planes of one alignment have <= 128 rows
constexpr int WSC_MAXLEN = 120;        // haplotype / window length (int8 pointers and CSR offsets)
constexpr int N_WBIN = 5;              // shared-memory bins: bytes per supercluster
__host__ __device__ inline int wsc_bin_cap(int b) { const int v[N_WBIN] = {2560, 5120, 10240, 24576, 57344}; return v[b]; }

__host__ __device__ inline int wa4(int x) { return (x + 3) & ~3; }

struct WscLayout {
    int hap[4];      // str L | flg L | ptr L (int8) | ins Lr
    int qm[2];       // rptr Lr | rflg Lr | toQ (Lq+1+Lr) | toR (Lr+1+Lq)
    int rseq;
    int tinf[2];     // per truth hap and column: base | tok << 8 (u16)
    int F[4];        // flag matrix [Lt][N]
    int walk[4];     // path (int16 q, int16 t, u8 flags) + Levenshtein row
    int total;
};
// hom: homozygous supercluster (replicate_hom) - only query hap 0, truth hap 2 and alignment 0 are laid out
__host__ __device__ inline WscLayout wsc_layout(const ScPlan &p, bool hom = false) {
    WscLayout m;
    int o = 0;
    const int Lr = p.lr;
    for (int h = 0; h < 4; h++) { m.hap[h] = o; if (!hom || !(h & 1)) o += wa4(3 * p.len[h] + Lr); }
    for (int k = 0; k < 2; k++) { m.qm[k] = o; if (!hom || k == 0) o += wa4(2 * Lr + 2 * (p.len[k] + Lr + 1)); }
    m.rseq = o; o += wa4(Lr);
    for (int k = 0; k < 2; k++) { m.tinf[k] = o; if (!hom || k == 0) o += wa4(2 * p.len[2 + k]); }
    for (int ai = 0; ai < 4; ai++) {
        m.F[ai] = m.walk[ai] = o;
        if (hom && ai > 0) continue;
        const int N = p.len[ai >> 1] + Lr, Lt = p.len[2 + (ai & 1)];
        m.F[ai] = o; o += wa4(N * Lt);
        m.walk[ai] = o;
        const int np = N + Lt + 4, mn = (Lr < Lt ? Lr : Lt) + 1;
        o += 2 * wa4(2 * np) + wa4(np) + wa4(2 * mn);
    }
    m.total = o;
    return m;
}
// register slots (32 rows each) the supercluster needs, 0 if it does not fit this kernel
__host__ __device__ inline int wsc_slots(const ScPlan &p) {
    int maxlen = p.lr;
    for (int h = 0; h < 4; h++) maxlen = p.len[h] > maxlen ? p.len[h] : maxlen;
    if (maxlen > WSC_MAXLEN) return 0;
    const int N = (p.len[0] > p.len[1] ? p.len[0] : p.len[1]) + p.lr;
    if (N > 32 * WSC_MAXSLOT) return 0;
    return (N + 31) / 32;
}
__host__ __device__ inline int wsc_bin(int need) { for (int b = 0; b < N_WBIN; b++) if (need <= wsc_bin_cap(b)) return b; return -1; }

struct PFWarp {      // path flags, [column][row], QUERY rows first
    const u8 *F; int N, Lq;
    __device__ __forceinline__ int get(int hi, int qri, int ti) const { return F[ti * N + (hi ? Lq + qri : qri)]; }
    __device__ __forceinline__ void prefetch(int, int, int) const {}
};

// resident blocks per SM the register allocation must allow, by register-slot count (build-time knobs)
// (S=1: 8 blocks = 64 registers, measured best; more slots need more registers per thread)
#ifndef VD_WSC_MINB_S1
#define VD_WSC_MINB_S1 8
#endif
#ifndef VD_WSC_MINB_S2
#define VD_WSC_MINB_S2 6
#endif
#ifndef VD_WSC_MINB_S3
#define VD_WSC_MINB_S3 4
#endif
constexpr int wsc_minb(int S) { return S == 1 ? VD_WSC_MINB_S1 : (S == 2 ? VD_WSC_MINB_S2 : VD_WSC_MINB_S3); }

// one alignment as the sweeps see it (all pointers into the supercluster's shared-memory region)
struct WscAln {
    const u8 *qstr, *qflg, *rseq, *rflg;
    const int8_t *qptr, *rptr, *toQ, *toR;
    const u16 *tinf;
    u8 *F;
    int Lq, Lr, Lt;
};

// Forward and backward sweep of one alignment by one warp (phase 2 of the header comment); on return F
// holds the path flags and every lane the score, the end plane, the origin plane and the status bits.
template <int S>
__device__ __forceinline__ void wsc_sweep(const int lane, const WscAln &X, int &score, int &end_plane, int &beg_plane, u32 &status) {
    constexpr unsigned FULL = 0xffffffffu;
    const u8 *qstr = X.qstr, *qflg = X.qflg, *rseq = X.rseq, *rflg = X.rflg;
    const int8_t *qptr = X.qptr, *rptr = X.rptr, *toQ = X.toQ, *toR = X.toR;
    const u16 *tinf = X.tinf;
    u8 *F = X.F;
    const int Lq = X.Lq, Lr = X.Lr, Lt = X.Lt, N = Lq + Lr;
    // values of an arbitrary row g / the row above / the row below, from per-slot register arrays
    auto row_get = [&](const int (&X)[S], int g) -> int {
        int v = __shfl_sync(FULL, X[0], g & 31);
#pragma unroll
        for (int s = 1; s < S; s++) { const int v1 = __shfl_sync(FULL, X[s], g & 31); if ((g >> 5) == s) v = v1; }
        return v;
    };
    auto above = [&](const int (&X)[S], int s) -> int {
        int u = __shfl_up_sync(FULL, X[s], 1);
        if (S > 1 && s > 0) { const int w = __shfl_sync(FULL, X[s - 1], 31); if (lane == 0) u = w; }
        return u;
    };
    auto below = [&](const int (&X)[S], int s) -> int {
        int u = __shfl_down_sync(FULL, X[s], 1);
        if (S > 1 && s < S - 1) { const int w = __shfl_sync(FULL, X[s + 1], 0); if (lane == 31) u = w; }
        return u;
    };

    // static per-row data
    bool inrow[S], P[S], below_ok[S];
    int a[S], ch[S], chn[S], swi[S], si[S], tpn[S], Ssum[S];
    int maxcnt = 0;
    int tpself[S];
#pragma unroll
    for (int s = 0; s < S; s++) {
        const int r = 32 * s + lane;
        inrow[s] = r < N;
        P[s] = r >= Lq;
        a[s] = P[s] ? r - Lq : r;
        const int len = P[s] ? Lr : Lq;
        const u8 *seq = P[s] ? rseq : qstr;
        ch[s] = inrow[s] ? seq[a[s]] : 0x100;
        below_ok[s] = inrow[s] && a[s] + 1 < len;
        chn[s] = below_ok[s] ? seq[a[s] + 1] : 0x100;
        swi[s] = 0; si[s] = 0; tpself[s] = 0;
        if (inrow[s]) {
            // my row as a swap DESTINATION: first source (row index over both planes) | count << 16
            const int8_t *tab = P[s] ? toR : toQ;
            const int k0 = tab[a[s]], k1 = tab[a[s] + 1];
            if (k1 > k0) swi[s] = ((P[s] ? 0 : Lq) + (int)tab[len + 1 + k0]) | ((k1 - k0) << 16);
            // my row as a swap SOURCE: valid | k << 1 | tp(dest) << 4 | dest row << 8   (:598-679)
            const int f = P[s] ? rflg[a[s]] : qflg[a[s]];
            const int d = (int)(P[s] ? rptr[a[s]] : qptr[a[s]]) + 1;      // plane-local row on the other plane
            const int ndst = P[s] ? Lq : Lr;
            if ((!(f & P_VARIANT) || (f & P_VAR_END)) && d > 0 && d < ndst) {
                const int df = P[s] ? qflg[d] : rflg[d];
                if (!(df & P_VARIANT) || (df & P_VAR_BEG)) {
                    const int8_t *dtab = P[s] ? toQ : toR;               // CSR of the destination plane
                    const int8_t *dsrc = dtab + ndst + 1;
                    int k = 0;
                    for (int j = dtab[d]; j < dtab[d + 1]; j++) if ((int)dsrc[j] == a[s]) k = j - dtab[d];
                    int tp = 0;
                    if (P[s]) tp = ((int)qptr[d] != (int)qptr[d - 1] + 1) || (df & P_VAR_BEG);      // dest on QUERY (:656-658)
                    si[s] = 1 | (k << 1) | (tp << 4) | (((P[s] ? 0 : Lq) + d) << 8);
                }
            }
            if (!P[s] && a[s] > 0) tpself[s] = ((int)qptr[a[s]] != (int)qptr[a[s] - 1] + 1) || (qflg[a[s]] & P_VAR_BEG);   // :572-574
        }
        maxcnt = max(maxcnt, swi[s] >> 16);
    }
    maxcnt = __reduce_max_sync(FULL, maxcnt);
    {   // tp of the row below, and S = number of tp rows below mine in my plane
        unsigned tm[S];
#pragma unroll
        for (int s = 0; s < S; s++) tm[s] = __ballot_sync(FULL, tpself[s] != 0);
#pragma unroll
        for (int s = 0; s < S; s++) {
            const int tb = below(tpself, s);
            tpn[s] = below_ok[s] ? tb : 0;
            int cnt = lane < 31 ? __popc(tm[s] >> (lane + 1)) : 0;
#pragma unroll
            for (int s2 = s + 1; s2 < S; s2++) cnt += __popc(tm[s2]);
            Ssum[s] = cnt;           // tp is zero on REF rows, so this is already plane-local
        }
    }
    const int erow_q = Lq - 1, erow_r = N - 1;

    // ---------------- forward ----------------
    int Dp[S];
#pragma unroll
    for (int s = 0; s < S; s++) Dp[s] = INF;
    for (int c = 0; c < Lt; c++) {
        const int tic = tinf[c];
        const int tch = tic & 0xff;
        const bool tok = tic >> 8;                                                               // :338-339
        int diag[S], del[S], swp[S], sbv[S], x[S];
#pragma unroll
        for (int s = 0; s < S; s++) {
            const int r = 32 * s + lane;
            const int upv = above(Dp, s);                                   // D[r-1][c-1]
            const int cnt = swi[s] >> 16;
            int best = row_get(Dp, swi[s] & 0xffff), sb = 0;
            if (!cnt) best = INF;
            if (maxcnt > 1) {                                               // rare: several sources (:347, :376)
                const int8_t *tab = P[s] ? toR : toQ;
                const int len = P[s] ? Lr : Lq;
                for (int k = 1; k < maxcnt; k++) {
                    const bool has = inrow[s] && k < cnt;
                    const int g = has ? (P[s] ? 0 : Lq) + (int)tab[len + 1 + (int)tab[a[s]] + k] : 0;
                    const int v = row_get(Dp, g);
                    if (has) {
                        if (v < best) { best = v; sb = k << F_K_SHIFT; }
                        else if (v == best) sb = (k << F_K_SHIFT) | F_TIE;   // keep the larger row
                    }
                }
            }
            const bool m = ch[s] == tch;
            diag[s] = (a[s] > 0 && c > 0) ? upv + (m ? 0 : 1) : INF;         // :324-332, :415-422
            del[s] = c > 0 ? Dp[s] + 1 : INF;                               // :406-413
            swp[s] = (tok && m) ? best : INF;                               // :334-349, :363-378
            sbv[s] = sb;
            int b = min(min(diag[s], del[s]), swp[s]);
            if (a[s] == 0 && c == 0) b = 0;                                 // both origins start at 0 (:299-305)
            if (!inrow[s]) b = INF;
            x[s] = b < INF / 2 ? b - r : INF;
        }
        // insertion chain D[r] = min(b[r], D[r-1] + 1) within a plane: D[r] = r + min_{j<=r} (b[j] - j)
        int Dn[S];
        int carryQ = INF, carryR = INF;
#pragma unroll
        for (int s = 0; s < S; s++) {
            const int r = 32 * s + lane;
            const int seg = P[s] ? Lq : 0;
            int incl = x[s];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int o = __shfl_up_sync(FULL, incl, d);
                if (lane >= d && r - d >= seg) incl = min(incl, o);
            }
            incl = min(incl, P[s] ? carryR : carryQ);
            if constexpr (S > 1) {
                if (s < S - 1) {                                            // totals so far, per plane
                    const int lastq = min(31, Lq - 1 - 32 * s), lastr = min(31, N - 1 - 32 * s);
                    const int tq = __shfl_sync(FULL, incl, max(lastq, 0)), tr = __shfl_sync(FULL, incl, max(lastr, 0));
                    if (lastq >= 0) carryQ = tq;
                    if (lastr >= 0 && 32 * s + lastr >= Lq) carryR = tr;
                }
            }
            Dn[s] = (inrow[s] && incl < INF / 2) ? incl + r : INF;
        }
#pragma unroll
        for (int s = 0; s < S; s++) {
            const int r = 32 * s + lane;
            const int dabove = above(Dn, s);                                // D[r-1][c]
            if (inrow[s]) {
                const int d = Dn[s];
                int f = 0;
                if (a[s] == 0 && c == 0) f = F_DIAG;
                else {
                    if (diag[s] == d) f |= F_DIAG;
                    if (a[s] > 0 && dabove + 1 == d) f |= F_INS;            // :397-404
                    if (del[s] == d) f |= F_DEL;
                    if (swp[s] == d) f |= F_SWP | sbv[s];
                }
                F[c * N + r] = (u8)f;
            }
            Dp[s] = Dn[s];
        }
    }
    const int dq = row_get(Dp, erow_q), dr = row_get(Dp, erow_r);          // :390-391
    score = min(dq, dr);
    end_plane = (dq == score) ? 0 : 1;                            // :436-440
    __syncwarp();

    // ---------------- backward ----------------
    status = 0;
    const int erow = end_plane ? erow_r : erow_q;
    int TFn[S];                                  // (T << 8) | forward flags of column c+1
#pragma unroll
    for (int s = 0; s < S; s++) TFn[s] = (-1) << 8;
    for (int c = Lt - 1; c >= 0; c--) {
        const bool last = c == Lt - 1;
        const int tch_next = last ? 0x200 : (tinf[c + 1] & 0xff);
        int Fc[S], B[S], U[S], T[S], swv[S], tfdv[S];
        bool link[S];
#pragma unroll
        for (int s = 0; s < S; s++) Fc[s] = inrow[s] ? F[c * N + 32 * s + lane] : 0;
#pragma unroll
        for (int s = 0; s < S; s++) {
            const int r = 32 * s + lane;
            const int tfd = below(TFn, s);                                   // (r+1, c+1)
            tfdv[s] = tfd;
            const int tf2 = row_get(TFn, si[s] >> 8);                        // my swap destination at c+1
            int b = -1;
            swv[s] = -1;
            if (inrow[s]) {
                if (last && r == erow) b = 0;                                // :543-545
                if (!last) {
                    const int Td = tfd >> 8, Fd = tfd & 0xff, Tn = TFn[s] >> 8, Fn = TFn[s] & 0xff;
                    if (below_ok[s] && Td >= 0 && (Fd & F_DIAG)) b = max(b, Td + tpn[s]);      // :556-595, :692-731
                    if (Tn >= 0 && (Fn & F_DEL)) b = max(b, Tn);                               // :774-804
                    if (si[s] & 1) {                                                           // :598-679
                        const int T2 = tf2 >> 8, F2 = tf2 & 0xff;
                        if (T2 >= 0 && (F2 & F_SWP) && (F2 >> F_K_SHIFT) == ((si[s] >> 1) & 7)) {
                            swv[s] = T2 + ((si[s] >> 4) & 1);
                            b = max(b, swv[s]);
                            if (F2 & F_TIE) status |= VD_ST_TIE;
                        }
                    }
                }
            }
            B[s] = b;
            const int fbelow = below(Fc, s);                                 // forward flags of (r+1, c)
            link[s] = below_ok[s] && (fbelow & F_INS);                       // :734-771
        }
        // in-column chain T[r] = max(B[r], T[r+1] + tp(r+1)) over unbroken links: suffix max of T - S
        int carryU = NEG;
#pragma unroll
        for (int s = S - 1; s >= 0; s--) {
            const unsigned lm = __ballot_sync(FULL, link[s]);
            const unsigned nm = ~(lm >> lane);
            const int run = nm ? __ffs(nm) - 1 : 32;
            int val = B[s] >= 0 ? B[s] - Ssum[s] : NEG;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int o = __shfl_down_sync(FULL, val, d);
                if (run >= d && lane + d < 32) val = max(val, o);
            }
            if (S > 1 && s < S - 1 && run == 32 - lane) val = max(val, carryU);
            U[s] = val;
            if (S > 1 && s > 0) carryU = __shfl_sync(FULL, val, 0);
            T[s] = (inrow[s] && val > NEG / 2) ? val + Ssum[s] : -1;
        }
#pragma unroll
        for (int s = 0; s < S; s++) {
            const int r = 32 * s + lane;
            const int tbelow = below(T, s);                                  // T[r+1][c]
            const int tfd = tfdv[s];
            int pf = 0;
            const int Tv = T[s];
            if (inrow[s] && Tv >= 0) {
                if (last && r == erow && Tv == 0) pf |= PTR_MAT;             // :543
                if (!last) {
                    const int Td = tfd >> 8, Fd = tfd & 0xff, Tn = TFn[s] >> 8, Fn = TFn[s] & 0xff;
                    if (below_ok[s] && Td >= 0 && (Fd & F_DIAG) && Td + tpn[s] == Tv)
                        pf |= (chn[s] == tch_next) ? PTR_MAT : PTR_SUB;
                    if (Tn >= 0 && (Fn & F_DEL) && Tn == Tv) pf |= PTR_DEL;
                    if (swv[s] >= 0 && swv[s] == Tv) pf |= PTR_SWP;
                }
                if (link[s] && tbelow >= 0 && tbelow + tpn[s] == Tv) pf |= PTR_INS;
            }
            if (inrow[s]) F[c * N + r] = (u8)pf;
        }
#pragma unroll
        for (int s = 0; s < S; s++) TFn[s] = (T[s] << 8) | (Fc[s] & 0xff);
    }
    const int t00 = __shfl_sync(FULL, TFn[0], 0) >> 8;
    beg_plane = t00 >= 0 ? 0 : 1;                                  // :811-814
    status = __reduce_or_sync(FULL, status);
}

// PAR = false: one warp per supercluster, its four alignments one after the other (small shared-memory
//               bins: occupancy is register-limited and there are plenty of superclusters);
// PAR = true:  one block per supercluster, warp w runs alignment w (big bins: few superclusters, occupancy is
//               shared-memory-limited, so the four warps share one supercluster's footprint and its latency
//               drops fourfold).
// HOM: homozygous superclusters (always !PAR): alignment Q1T1 only, then replicate_hom.
template <int S, bool PAR, bool HOM>
__global__ void __launch_bounds__(WSC_TPB, wsc_minb(S))
wsc_kernel(BatchDev in, OutDev out, const ScPlan *__restrict__ plan, const int *__restrict__ order, int count, int warp_bytes) {
    VD_DYN_SHARED(smem);
    constexpr unsigned FULL = 0xffffffffu;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slot = PAR ? blockIdx.x : blockIdx.x * (WSC_TPB / 32) + warp;
    if (slot >= count) return;                      // !PAR: warps are independent, no block-wide barrier anywhere
    auto sync = [&]() { if constexpr (PAR) __syncthreads(); else __syncwarp(); };
    const bool lead = !PAR || warp == 0;            // the warp that runs the single-lane phases of the supercluster
    const int sc = order[slot];
    const ScPlan p = plan[sc];
    static_assert(!(PAR && HOM), "homozygous superclusters run one warp per supercluster");
    const WscLayout M = wsc_layout(p, HOM);
    u8 *base = PAR ? smem : smem + warp * warp_bytes;
    const int Lr = p.lr;
    auto hstr = [&](int h) { return base + M.hap[h]; };
    auto hflg = [&](int h) { return base + M.hap[h] + p.len[h]; };
    auto hptr = [&](int h) { return (int8_t *)(base + M.hap[h] + 2 * p.len[h]); };
    auto hins = [&](int h) { return base + M.hap[h] + 3 * p.len[h]; };
    auto qrptr = [&](int k) { return (int8_t *)(base + M.qm[k]); };
    auto qrflg = [&](int k) { return base + M.qm[k] + Lr; };
    auto qtoQ = [&](int k) { return (int8_t *)(base + M.qm[k] + 2 * Lr); };
    auto qtoR = [&](int k) { return (int8_t *)(base + M.qm[k] + 2 * Lr + (p.len[k] + 1 + Lr)); };
    u8 *rseq = base + M.rseq;

    // ---- phase 1: expansion ----
    bool ok = true;
    if (!lead) {
        // PAR: the other warps wait for the expansion
    } else if (lane < 4) {
        const int h = lane;
        const bool isq = h < 2;
        if (HOM && (h & 1)) { /* same as haplotype h-1 */ } else {
        const int len = expand_hap<int8_t>(in, sc, h, hstr(h), hflg(h), hptr(h), isq ? qrptr(h) : nullptr,
                                           isq ? qrflg(h) : nullptr, hins(h), p.len[h]);
        ok = len == p.len[h];
        if (ok && isq) ok = build_swsrc<int8_t>(hptr(h), hflg(h), len, qtoR(h), Lr) &&
                            build_swsrc<int8_t>(qrptr(h), qrflg(h), Lr, qtoQ(h), len);
        }
    } else {
        const u8 *rs = in.rplane_seq + in.ref_off[sc];
        for (int k = lane - 4; k < Lr; k += 28) rseq[k] = rs[k];
    }
    if constexpr (PAR) {
        __shared__ int s_ok;
        if (threadIdx.x == 0) s_ok = 1;
        __syncthreads();
        if (!ok) s_ok = 0;
        __syncthreads();
        ok = s_ok != 0;
    }
    if (__ballot_sync(FULL, ok) != FULL) {
        if (lead && lane < 4) { out.status[4 * (int64_t)sc + lane] = ST_BAD; out.aln_score[4 * (int64_t)sc + lane] = -1; }
        return;
    }
    __syncwarp();
    for (int k = 0; k < (HOM ? 1 : 2); k++) {        // truth columns: base | tok << 8   (:338-339, :367-368)
        const u8 *ts = hstr(2 + k), *tf = hflg(2 + k);
        u16 *ti = (u16 *)(base + M.tinf[k]);
        for (int c = PAR ? (int)threadIdx.x : lane; c < p.len[2 + k]; c += PAR ? WSC_TPB : 32) {
            const bool tok = c > 0 && (!(tf[c - 1] & P_VARIANT) || (tf[c - 1] & P_VAR_END));
            ti[c] = (u16)(ts[c] | (tok ? 0x100 : 0));
        }
    }
    sync();

    int my_score = 0, my_end = 0, my_beg = 0;
    u32 my_status = 0;

    // An alignment whose query and truth haplotypes both carry no variant compares the window with
    // itself: score 0, both path ends on the QUERY plane, no status bit, no variant to credit (the
    // Q1T1 alignment of every heterozygous site).  trivial: bit ai.  Not used when the REF plane has its
    // own string (rplane_seq): its sections could then differ from the truth and raise WARN status bits.
    unsigned trivial = 0;
    if (in.rplane_seq == in.ref_seq) {
        const int64_t *vo = in.var_off + 4 * (int64_t)sc;
        const int64_t o0 = vo[0], o1 = vo[1], o2 = vo[2], o3 = vo[3], o4 = vo[4];
        const bool eq0 = o1 == o0, eq1 = o2 == o1, et0 = o3 == o2, et1 = o4 == o3;
        trivial = (eq0 && et0 ? 1u : 0u) | (eq0 && et1 ? 2u : 0u) | (eq1 && et0 ? 4u : 0u) | (eq1 && et1 ? 8u : 0u);
    }

    // ---- phase 2: the four alignments ----
    for (int ai = PAR ? warp : 0; ai < (PAR ? warp + 1 : (HOM ? 1 : 4)); ai++) {
        if ((trivial >> ai) & 1) continue;          // my_score = my_end = my_beg = my_status = 0
        const int qh = ai >> 1, th = 2 + (ai & 1);
        const int Lq = p.len[qh], Lt = p.len[th];
        const u8 *qstr = hstr(qh), *qflg = hflg(qh);
        const u16 *tinf = (const u16 *)(base + M.tinf[ai & 1]);
        const int8_t *qptr = hptr(qh), *rptr = qrptr(qh), *toQ = qtoQ(qh), *toR = qtoR(qh);
        const u8 *rflg = qrflg(qh);
        u8 *F = base + M.F[ai];

        WscAln X{qstr, qflg, rseq, rflg, qptr, rptr, toQ, toR, tinf, F, Lq, Lr, Lt};
        int score, end_plane, beg_plane;
        u32 status;
        wsc_sweep<S>(lane, X, score, end_plane, beg_plane, status);
        if (lane == (PAR ? 0 : ai)) { my_score = score; my_end = end_plane; my_beg = beg_plane; my_status = status; }
        __syncwarp();
    }

    // ---- phase 3: walk + credit, one lane per alignment ----
    if (PAR ? lane == 0 : lane < (HOM ? 1 : 4)) {
        const int ai = PAR ? warp : lane, qh = ai >> 1, th = 2 + (ai & 1);
        const int Lq = p.len[qh], Lt = p.len[th], N = Lq + Lr;
        Hap<int8_t> q{Lq, hstr(qh), hflg(qh), hptr(qh), hins(qh)};
        Hap<int8_t> t{Lt, hstr(th), hflg(th), hptr(th), hins(th)};
        QMaps<int8_t> qm{qrptr(qh), qrflg(qh), qtoQ(qh), qtoR(qh)};
        SMemIL mem{base + M.walk[ai]};
        AlnLayout<int> L;
        const int np = N + Lt + 4;
        L.oPF = L.oF = L.oD0 = L.oD1 = L.oT0 = L.oT1 = 0;
        L.oPQ = 0; L.oPT = wa4(2 * np); L.oPS = 2 * wa4(2 * np); L.oLev = L.oPS + wa4(np); L.total = 0;
        PFWarp pfr{base + M.F[ai], N, Lq};
        u32 status = my_status;
        if (!((trivial >> ai) & 1))
            walk_credit<SMemIL, 2, int8_t>(mem, L, pfr, q, qm, t, rseq, Lr, my_beg, my_end, in, out, sc, ai, status);
        const int64_t oi = 4 * (int64_t)sc + ai;
        for (int k = 0; k < (HOM ? 4 : 1); k++) {
            out.aln_score[oi + k] = my_score;
            out.aln_end_plane[oi + k] = (u8)my_end;
            out.aln_beg_plane[oi + k] = (u8)my_beg;
            out.status[oi + k] = status;
        }
        if constexpr (HOM) replicate_hom(in, out, sc);
    }
}


// ------------------------------------------------------------------------------------------
// Block-shared variant of the warp-per-supercluster kernel (small shared-memory bins): the four
// warps of a block still run the sweeps of their own supercluster, but the single-lane phases
// (expansion + swap tables before, walk + credit after) of all four superclusters are done by
// warp 0, one lane per haplotype / alignment: 16 busy lanes in one warp instead of 4 busy lanes in
// each of four warps, i.e. a quarter of the issue slots for the part that cannot use more lanes.
// ------------------------------------------------------------------------------------------
#ifndef VD_WSC_P3_PER_WARP
#define VD_WSC_P3_PER_WARP 0
#endif
#ifdef VD_PHASE_PROF
// build-time profiling hook (never in the product build): cycles of the block's phases, summed per kernel flavour
__device__ unsigned long long g_wsc_phase[8][6];
#define VD_PH_DECL long long ph_t = clock64(); const int ph_k = (S - 1) * 2 + (HOM ? 1 : 0)
#define VD_PH_MARK(i) do { if (threadIdx.x == 0) { const long long n_ = clock64(); atomicAdd(&g_wsc_phase[ph_k][i], (unsigned long long)(n_ - ph_t)); ph_t = n_; } } while (0)
#else
#define VD_PH_DECL
#define VD_PH_MARK(i)
#endif

struct WscDesc {
    ScPlan p; WscLayout M;
    int sc;            // -1: no supercluster for this warp (tail of the launch)
    int ok;            // expansion succeeded
    unsigned trivial;  // alignments without variants on either side
    int staged;        // the variant records and ALT bytes were staged in shared memory (else: read from the batch)
    int vb[5];         // first staged variant of each haplotype (and the end)
    int res[4][4];     // per alignment: score, end plane, origin plane, status
};

// staged copy of a supercluster's variant records, in the (not yet used) flag-matrix area: pos | rlen | alt_off low words
// (nv + 1) | type | ALT bytes
struct WscStage {
    int *pos, *rlen, *aoff; u8 *type, *alt;
    __device__ WscStage(u8 *base, const WscLayout &M, int nv) {
        pos = (int *)(base + ((M.F[0] + 3) & ~3)); rlen = pos + nv; aoff = rlen + nv;
        type = (u8 *)(aoff + nv + 1); alt = type + nv;
    }
    static __device__ int bytes(int nv, int alt_bytes) { return 13 * nv + 8 + alt_bytes; }
};

template <int S, bool HOM, int NW>
__global__ void __launch_bounds__(32 * NW, wsc_minb(S))
wsc_block_kernel(BatchDev in, OutDev out, const ScPlan *__restrict__ plan, const int *__restrict__ order, int count, int warp_bytes) {
    VD_DYN_SHARED(smem);
    __shared__ WscDesc desc[NW];
    VD_PH_DECL;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slot = blockIdx.x * (NW) + warp;
    const bool live = slot < count;
    constexpr unsigned FULLM = 0xffffffffu;
    // ---- phase 0: every warp describes its own supercluster and stages its input: the variant records, the ALT bytes
    //      and the window, read with all lanes at once (four rounds of loads instead of one per field and variant) ----
    if (lane == 0) {
        WscDesc &d = desc[warp];
        d.sc = live ? order[slot] : -1;
        d.ok = 1;
        d.trivial = 0;
        d.staged = 0;
        if (live) {
            d.p = plan[d.sc];
            d.M = wsc_layout(d.p, HOM);
        }
    }
    __syncwarp();
    if (live) {
        WscDesc &d = desc[warp];
        const int sc = d.sc;
        u8 *base = smem + warp * warp_bytes;
        const int64_t myvo = lane < 5 ? in.var_off[4 * (int64_t)sc + lane] : 0;
        const int64_t r0 = in.ref_off[sc];
        const int64_t v0 = __shfl_sync(FULLM, myvo, 0);
        const int nv = (int)(__shfl_sync(FULLM, myvo, 4) - v0);
        if (lane < 5) d.vb[lane] = (int)(myvo - v0);
        const int64_t nxvo = __shfl_down_sync(FULLM, myvo, 1);
        const unsigned e = __ballot_sync(FULLM, lane < 4 && nxvo == myvo);       // bit h: haplotype h carries no variant
        if (lane == 0 && in.rplane_seq == in.ref_seq)                            // see wsc_kernel
            d.trivial = ((e & 5u) == 5u ? 1u : 0u) | ((e & 9u) == 9u ? 2u : 0u) | ((e & 6u) == 6u ? 4u : 0u) | ((e & 10u) == 10u ? 8u : 0u);
        const int room = warp_bytes - ((d.M.F[0] + 3) & ~3);
        bool staged = in.rplane_seq == in.ref_seq && WscStage::bytes(nv, 0) <= room;
        WscStage st(base, d.M, nv);
        int64_t a0 = 0;
        if (staged) {
            for (int v = lane; v <= nv; v += 32) {
                const int64_t ao = in.alt_off[v0 + v];
                if (v == 0) a0 = ao;
                st.aoff[v] = (int)ao;
                if (v < nv) { st.pos[v] = in.var_pos[v0 + v]; st.rlen[v] = in.var_rlen[v0 + v]; st.type[v] = in.var_type[v0 + v]; }
            }
        }
        const u8 *rs = in.rplane_seq + r0;
        u8 *rseq = base + d.M.rseq;
        for (int k = lane; k < d.p.lr; k += 32) rseq[k] = rs[k];
        __syncwarp();
        if (staged) {
            a0 = __shfl_sync(FULLM, a0, 0);
            const int ab = st.aoff[nv] - st.aoff[0];
            staged = ab >= 0 && WscStage::bytes(nv, ab) <= room;
            if (staged) for (int k = lane; k < ab; k += 32) st.alt[k] = in.alt_seq[a0 + k];
        }
        if (lane == 0) d.staged = staged ? 1 : 0;
    }
    __syncthreads();
    VD_PH_MARK(0);
    // shared-memory views of supercluster w
    auto hbase = [&](int w, int h) { return smem + w * warp_bytes + desc[w].M.hap[h]; };
    auto qbase = [&](int w, int k) { return smem + w * warp_bytes + desc[w].M.qm[k]; };

    // ---- phase 1: warp 0, lane 4w+h expands haplotype h of supercluster w ----
    if (warp == 0 && lane < 4 * (NW)) {
        const int w = lane >> 2, h = lane & 3;
        const WscDesc &d = desc[w];
        if (d.sc >= 0 && !(HOM && (h & 1))) {
            const int L = d.p.len[h], Lr = d.p.lr;
            u8 *str = hbase(w, h), *flg = str + L, *ins = str + 3 * L;
            int8_t *ptr = (int8_t *)(str + 2 * L);
            const bool isq = h < 2;
            int8_t *rptr = isq ? (int8_t *)qbase(w, h) : nullptr;
            u8 *rflg = isq ? qbase(w, h) + Lr : nullptr;
            int len;
            if (d.staged) {
                const WscStage st(smem + w * warp_bytes, d.M, d.vb[4]);
                const HapSrcStaged src{smem + w * warp_bytes + d.M.rseq, Lr, d.vb[h], d.vb[h + 1] - d.vb[h], st.pos, st.rlen, st.aoff, st.type, st.alt};
                len = expand_hap_src<int8_t>(src, str, flg, ptr, rptr, rflg, ins, L);
            } else {
                len = expand_hap<int8_t>(in, d.sc, h, str, flg, ptr, rptr, rflg, ins, L);
            }
            bool ok = len == L;
            if (ok && isq) {
                int8_t *toQ = (int8_t *)(qbase(w, h) + 2 * Lr), *toR = toQ + (L + 1 + Lr);
                ok = build_swsrc<int8_t>(ptr, flg, len, toR, Lr) && build_swsrc<int8_t>(rptr, rflg, Lr, toQ, len);
            }
            if (!ok) desc[w].ok = 0;
        }
    }
    __syncthreads();
    VD_PH_MARK(1);

    // ---- phase 2: every warp sweeps the alignments of its own supercluster ----
    if (live && desc[warp].ok) {
        const WscDesc &d = desc[warp];
        u8 *base = smem + warp * warp_bytes;
        const int Lr = d.p.lr;
        for (int k = 0; k < (HOM ? 1 : 2); k++) {    // truth columns: base | tok << 8   (:338-339, :367-368)
            const int Ltk = d.p.len[2 + k];
            const u8 *ts = hbase(warp, 2 + k), *tf = ts + Ltk;
            u16 *ti = (u16 *)(base + d.M.tinf[k]);
            for (int c = lane; c < Ltk; c += 32) {
                const bool tok = c > 0 && (!(tf[c - 1] & P_VARIANT) || (tf[c - 1] & P_VAR_END));
                ti[c] = (u16)(ts[c] | (tok ? 0x100 : 0));
            }
        }
        __syncwarp();
        for (int ai = 0; ai < (HOM ? 1 : 4); ai++) {
            int score = 0, end_plane = 0, beg_plane = 0;
            u32 status = 0;
            if (!((d.trivial >> ai) & 1)) {
                const int qh = ai >> 1, th = 2 + (ai & 1);
                const int Lq = d.p.len[qh], Lt = d.p.len[th];
                const u8 *qs = hbase(warp, qh);
                const u8 *qb = qbase(warp, qh);
                WscAln X{qs, qs + Lq, base + d.M.rseq, qb + Lr, (const int8_t *)(qs + 2 * Lq), (const int8_t *)qb,
                         (const int8_t *)(qb + 2 * Lr), (const int8_t *)(qb + 2 * Lr + (Lq + 1 + Lr)),
                         (const u16 *)(base + d.M.tinf[ai & 1]), base + d.M.F[ai], Lq, Lr, Lt};
                wsc_sweep<S>(lane, X, score, end_plane, beg_plane, status);
            }
            if (lane == 0) { int *r = desc[warp].res[ai]; r[0] = score; r[1] = end_plane; r[2] = beg_plane; r[3] = (int)status; }
            __syncwarp();
        }
    }
#if VD_WSC_P3_PER_WARP
    __syncwarp();
#else
    __syncthreads();
#endif
    VD_PH_MARK(2);

    // ---- phase 3: warp 0, lane 4w+ai walks alignment ai of supercluster w (VD_WSC_P3_PER_WARP: every warp its own
    //      supercluster on lanes 0-3 - four times the issue slots for a quarter of the latency) ----
#ifdef VD_PHASE_PROF
    if (VD_WSC_P3_PER_WARP || warp == 0) {
        __syncwarp();
        if (lane < (VD_WSC_P3_PER_WARP ? 4 : 4 * (NW))) [&]() {
#else
    if (VD_WSC_P3_PER_WARP ? lane < 4 : (warp == 0 && lane < 4 * (NW))) {
#endif
        const int w = VD_WSC_P3_PER_WARP ? warp : lane >> 2, ai = lane & 3;
        const WscDesc &d = desc[w];
        if (d.sc < 0 || (HOM && ai)) return;
        const int sc = d.sc;
        const int64_t oi = 4 * (int64_t)sc + ai;
        if (!d.ok) {
            for (int k = 0; k < (HOM ? 4 : 1); k++) { out.status[oi + k] = ST_BAD; out.aln_score[oi + k] = -1; }
            return;
        }
        const int qh = ai >> 1, th = 2 + (ai & 1);
        const int Lr = d.p.lr, Lq = d.p.len[qh], Lt = d.p.len[th], N = Lq + Lr;
        u8 *base = smem + w * warp_bytes;
        const u8 *qs = hbase(w, qh), *ts = hbase(w, th), *qb = qbase(w, qh);
        Hap<int8_t> q{Lq, qs, qs + Lq, (const int8_t *)(qs + 2 * Lq), qs + 3 * Lq};
        Hap<int8_t> t{Lt, ts, ts + Lt, (const int8_t *)(ts + 2 * Lt), ts + 3 * Lt};
        QMaps<int8_t> qm{(const int8_t *)qb, qb + Lr, (const int8_t *)(qb + 2 * Lr), (const int8_t *)(qb + 2 * Lr + (Lq + 1 + Lr))};
        SMemIL mem{base + d.M.walk[ai]};
        AlnLayout<int> L;
        const int np = N + Lt + 4;
        L.oPF = L.oF = L.oD0 = L.oD1 = L.oT0 = L.oT1 = 0;
        L.oPQ = 0; L.oPT = wa4(2 * np); L.oPS = 2 * wa4(2 * np); L.oLev = L.oPS + wa4(np); L.total = 0;
        PFWarp pfr{base + d.M.F[ai], N, Lq};
        const int *r = d.res[ai];
        u32 status = (u32)r[3];
        if (!((d.trivial >> ai) & 1))
            walk_credit<SMemIL, 2, int8_t>(mem, L, pfr, q, qm, t, base + d.M.rseq, Lr, r[2], r[1], in, out, sc, ai, status);
        for (int k = 0; k < (HOM ? 4 : 1); k++) {
            out.aln_score[oi + k] = r[0];
            out.aln_end_plane[oi + k] = (u8)r[1];
            out.aln_beg_plane[oi + k] = (u8)r[2];
            out.status[oi + k] = status;
        }
        if constexpr (HOM) replicate_hom(in, out, sc);
#ifdef VD_PHASE_PROF
        }();
        __syncwarp();
        VD_PH_MARK(3);
        if (threadIdx.x == 0) atomicAdd(&g_wsc_phase[ph_k][4], 1ull);
    }
#else
    }
#endif
}

// ------------------------------------------------------------------------------------------
// Split form of the same computation (the default): three launches per group instead of one fused kernel.
// In the fused kernels the single-lane phases (expansion + swap tables before the sweeps, walk + credit after) run as ONE
// warp per block while the block's other warps wait at a barrier and its registers and shared memory stay allocated:
// measured, they are 50 % of a block's lifetime for 10 % of its instructions (profiles/, r2_b9).  Here they are kernels
// of their own with one THREAD per haplotype / alignment over the whole group - hundreds of thousands of independent
// threads instead of one warp per block - and the warp sweeps keep the SMs to themselves.  The hand-off goes through a
// slab in HBM (one slot of the group's shared-memory bin size per supercluster, laid out like the shared memory of the
// fused kernels: WscLayout): about 3 KB per supercluster written once and read once, against 20 us of a block's time.
//
//   wsc_expand_kernel   thread per (supercluster, haplotype): expand_hap + swap tables + truth-column info -> slab
//   wsc_sweep_kernel<S> block per supercluster, warp per alignment: the expanded part of the slot comes in by one TMA
//                       bulk copy (cp.async.bulk + mbarrier), forward and backward sweep in shared memory as before,
//                       path flags back to the slot
//   wsc_walk_kernel     thread per alignment: walk + credit over the slot
// ------------------------------------------------------------------------------------------
struct WscHdr { int ok; unsigned trivial; };
// alignments the sweeps have finished and the walk kernel has to do, (global slot << 2 | alignment), in completion order:
// the walk kernel then runs on dense warps instead of one thread per (slot, alignment) of which two thirds have nothing to do
struct WscWork { int *count; int *item; int slot_base; };

// the wsc launch groups of one chunk in slot order (kernel parameter of the two thread-per-item kernels, which run once
// over all groups: a launch of a few thousand threads would be all latency)
constexpr int WSC_MAXG = 2 * WSC_MAXSLOT * N_WBIN;
struct WscGroups {
    int n;
    int first[WSC_MAXG + 1];         // first slot of group k, first[n] = number of slots
    int order_first[WSC_MAXG];       // where the group's superclusters start in order[]
    int stride[WSC_MAXG];            // slab bytes per slot
    int hom[WSC_MAXG];
    long long base[WSC_MAXG];        // slab offset of the group's first slot
};
struct WscSlot { int sc, hom, stride; u8 *base; };
__device__ __forceinline__ WscSlot wsc_slot(const WscGroups &G, int slot, const int *__restrict__ order, u8 *slab) {
    int k = 0;
    while (k + 1 < G.n && slot >= G.first[k + 1]) k++;
    const int j = slot - G.first[k];
    return WscSlot{order[G.order_first[k] + j], G.hom[k], G.stride[k], slab + G.base[k] + (int64_t)j * G.stride[k]};
}

__global__ void __launch_bounds__(128)
wsc_expand_kernel(BatchDev in, const ScPlan *__restrict__ plan, const int *__restrict__ order, const __grid_constant__ WscGroups G, u8 *slab, WscHdr *hdr) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    const int slot = g >> 2, h = g & 3, lane = threadIdx.x & 31;
    const bool live = slot < G.first[G.n];
    bool ok = true;
    unsigned trivial = 0;
    if (live) {
        const WscSlot ws = wsc_slot(G, slot, order, slab);
        const int sc = ws.sc;
        const bool HOM = ws.hom;
        const ScPlan p = plan[sc];
        const WscLayout M = wsc_layout(p, HOM);
        u8 *base = ws.base;
        const int L = p.len[h], Lr = p.lr;
        if (!(HOM && (h & 1))) {
            // built in thread-local memory (L1-cached: the swap tables re-read what the expansion wrote), then copied to
            // the slot region by region, which have the same internal layout
            __align__(4) u8 lh[(4 * WSC_MAXLEN + 3) & ~3], lq[(2 * WSC_MAXLEN + 2 * (2 * WSC_MAXLEN + 1) + 3) & ~3];
            u8 *str = lh, *flg = str + L, *ins = str + 3 * L;
            int8_t *ptr = (int8_t *)(str + 2 * L);
            const bool isq = h < 2;
            int8_t *rptr = isq ? (int8_t *)lq : nullptr;
            u8 *rflg = isq ? lq + Lr : nullptr;
            const int len = expand_hap<int8_t>(in, sc, h, str, flg, ptr, rptr, rflg, ins, L);
            ok = len == L;
            if (ok && isq) {
                int8_t *toQ = (int8_t *)(lq + 2 * Lr), *toR = toQ + (L + 1 + Lr);
                ok = build_swsrc<int8_t>(ptr, flg, len, toR, Lr) && build_swsrc<int8_t>(rptr, rflg, Lr, toQ, len);
            }
            if (ok) {
                u32 *dst = (u32 *)(base + M.hap[h]);
                const u32 *src = (const u32 *)lh;
                for (int k = 0; k < (3 * L + Lr + 3) / 4; k++) dst[k] = src[k];
                if (isq) {
                    dst = (u32 *)(base + M.qm[h]); src = (const u32 *)lq;
                    for (int k = 0; k < (2 * Lr + 2 * (L + Lr + 1) + 3) / 4; k++) dst[k] = src[k];
                }
            }
            if (ok && !isq) {                        // truth columns: base | tok << 8   (:338-339, :367-368)
                u16 *ti = (u16 *)(base + M.tinf[h - 2]);
                for (int c = 0; c < L; c++) {
                    const bool tok = c > 0 && (!(flg[c - 1] & P_VARIANT) || (flg[c - 1] & P_VAR_END));
                    ti[c] = (u16)(str[c] | (tok ? 0x100 : 0));
                }
            }
        }
        if (h == 1) {                                // the REF-plane string
            const u8 *rs = in.rplane_seq + in.ref_off[sc];
            u8 *rseq = base + M.rseq;
            for (int k = 0; k < Lr; k++) rseq[k] = rs[k];
        }
        if (h == 3 && in.rplane_seq == in.ref_seq) { // see wsc_kernel
            const int64_t *vo = in.var_off + 4 * (int64_t)sc;
            const int64_t o0 = vo[0], o1 = vo[1], o2 = vo[2], o3 = vo[3], o4 = vo[4];
            const bool eq0 = o1 == o0, eq1 = o2 == o1, et0 = o3 == o2, et1 = o4 == o3;
            trivial = (eq0 && et0 ? 1u : 0u) | (eq0 && et1 ? 2u : 0u) | (eq1 && et0 ? 4u : 0u) | (eq1 && et1 ? 8u : 0u);
        }
    }
    // the four threads of a supercluster are neighbouring lanes
    const unsigned okm = __ballot_sync(0xffffffffu, ok);
    trivial = __shfl_sync(0xffffffffu, trivial, lane | 3);
    if (live && h == 0) hdr[slot] = WscHdr{((okm >> (lane & ~3)) & 15u) == 15u ? 1 : 0, trivial};
}

template <int S, bool HOM>
__global__ void __launch_bounds__(HOM ? 32 : 128, HOM ? 16 : wsc_minb(S))
wsc_sweep_kernel(BatchDev in, OutDev out, const ScPlan *__restrict__ plan, const int *__restrict__ order, int count, u8 *slab, int stride,
                 const WscHdr *__restrict__ hdr, WscWork work) {
    VD_DYN_SHARED(smem);
    __shared__ __align__(8) unsigned long long s_bar;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slot = blockIdx.x;
    const int sc = order[slot];
    const ScPlan p = plan[sc];
    const WscLayout M = wsc_layout(p, HOM);
    const WscHdr hd = hdr[slot];
    u8 *gbase = slab + (int64_t)slot * stride;
    const int64_t oi = 4 * (int64_t)sc + (HOM ? 0 : warp);
    if (!hd.ok) {
        if (lane == 0) for (int k = 0; k < (HOM ? 4 : 1); k++) { out.status[oi + k] = ST_BAD; out.aln_score[oi + k] = -1; }
        return;
    }
    // the expanded haplotypes, maps and tables of the slot: one TMA bulk copy, every thread waits on its barrier
    if (threadIdx.x == 0) { VD_MBAR_INIT(&s_bar, 1); VD_MBAR_INIT_FENCE(); }
    __syncthreads();
    if (threadIdx.x == 0) VD_BULK_G2S(smem, gbase, (unsigned)((M.F[0] + 15) & ~15), &s_bar);
    VD_MBAR_WAIT(&s_bar, 0u);
    __syncthreads();
    const int ai = HOM ? 0 : warp;
    int score = 0, end_plane = 0, beg_plane = 0;
    u32 status = 0;
    if (!((hd.trivial >> ai) & 1)) {
        const int qh = ai >> 1, th = 2 + (ai & 1);
        const int Lr = p.lr, Lq = p.len[qh], Lt = p.len[th], N = Lq + Lr;
        const u8 *qs = smem + M.hap[qh], *qb = smem + M.qm[qh];
        WscAln X{qs, qs + Lq, smem + M.rseq, qb + Lr, (const int8_t *)(qs + 2 * Lq), (const int8_t *)qb,
                 (const int8_t *)(qb + 2 * Lr), (const int8_t *)(qb + 2 * Lr + (Lq + 1 + Lr)),
                 (const u16 *)(smem + M.tinf[ai & 1]), smem + M.F[ai], Lq, Lr, Lt};
        wsc_sweep<S>(lane, X, score, end_plane, beg_plane, status);
        __syncwarp();
        // path flags to the slot (the regions are 4-byte aligned in both places)
        const u32 *src = (const u32 *)(smem + M.F[ai]);
        u32 *dst = (u32 *)(gbase + M.F[ai]);
        for (int k = lane; k < (N * Lt + 3) / 4; k += 32) dst[k] = src[k];
    }
    if (lane == 0) {
        for (int k = 0; k < (HOM ? 4 : 1); k++) {
            out.aln_score[oi + k] = score;
            out.aln_end_plane[oi + k] = (u8)end_plane;
            out.aln_beg_plane[oi + k] = (u8)beg_plane;
            out.status[oi + k] = status;
        }
        if (!((hd.trivial >> ai) & 1)) {
            __threadfence();                                   // records and path flags before the work item
            work.item[atomicAdd(work.count, 1)] = ((work.slot_base + slot) << 2) | ai;
        }
    }
}

// Small bins: one WARP per supercluster (its alignments one after the other, as in the fused kernel), four independent
// warps per block - no block-wide barrier; each warp pulls its slot's expanded part in with its own TMA bulk copy.
template <int S, bool HOM>
__global__ void __launch_bounds__(WSC_TPB, wsc_minb(S))
wsc_sweep_warp_kernel(BatchDev in, OutDev out, const ScPlan *__restrict__ plan, const int *__restrict__ order, int count, u8 *slab, int stride,
                      const WscHdr *__restrict__ hdr, WscWork work) {
    VD_DYN_SHARED(smem);
    __shared__ __align__(8) unsigned long long s_bar[WSC_TPB / 32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slot = blockIdx.x * (WSC_TPB / 32) + warp;
    if (slot >= count) return;
    const int sc = order[slot];
    const ScPlan p = plan[sc];
    const WscLayout M = wsc_layout(p, HOM);
    const WscHdr hd = hdr[slot];
    u8 *gbase = slab + (int64_t)slot * stride;
    u8 *base = smem + warp * stride;
    const int64_t oi = 4 * (int64_t)sc;
    if (!hd.ok) {
        if (lane < 4) { out.status[oi + lane] = ST_BAD; out.aln_score[oi + lane] = -1; }
        return;
    }
    if (lane == 0) { VD_MBAR_INIT(&s_bar[warp], 1); VD_MBAR_INIT_FENCE(); }
    __syncwarp();
    if (lane == 0) VD_BULK_G2S(base, gbase, (unsigned)((M.F[0] + 15) & ~15), &s_bar[warp]);
    VD_MBAR_WAIT(&s_bar[warp], 0u);
    __syncwarp();
    const int Lr = p.lr;
    for (int ai = 0; ai < (HOM ? 1 : 4); ai++) {
        int score = 0, end_plane = 0, beg_plane = 0;
        u32 status = 0;
        if (!((hd.trivial >> ai) & 1)) {
            const int qh = ai >> 1, th = 2 + (ai & 1);
            const int Lq = p.len[qh], Lt = p.len[th], N = Lq + Lr;
            const u8 *qs = base + M.hap[qh], *qb = base + M.qm[qh];
            u8 *F = base + M.F[0];                   // one flag matrix at a time: it leaves for the slot right after its sweeps
            WscAln X{qs, qs + Lq, base + M.rseq, qb + Lr, (const int8_t *)(qs + 2 * Lq), (const int8_t *)qb,
                     (const int8_t *)(qb + 2 * Lr), (const int8_t *)(qb + 2 * Lr + (Lq + 1 + Lr)),
                     (const u16 *)(base + M.tinf[ai & 1]), F, Lq, Lr, Lt};
            wsc_sweep<S>(lane, X, score, end_plane, beg_plane, status);
            __syncwarp();
            const u32 *src = (const u32 *)F;
            u32 *dst = (u32 *)(gbase + M.F[ai]);
            for (int k = lane; k < (N * Lt + 3) / 4; k += 32) dst[k] = src[k];
            __syncwarp();
        }
        if (lane == 0)
            for (int k = ai; k < (HOM ? 4 : ai + 1); k++) {
                out.aln_score[oi + k] = score;
                out.aln_end_plane[oi + k] = (u8)end_plane;
                out.aln_beg_plane[oi + k] = (u8)beg_plane;
                out.status[oi + k] = status;
            }
    }
    if (lane == 0) {                                           // this supercluster's alignments for the walk kernel
        const unsigned todo = (HOM ? 1u : 15u) & ~hd.trivial;
        const int n = __popc(todo);
        if (n) {
            __threadfence();
            int at = atomicAdd(work.count, n);
            for (int ai = 0; ai < 4; ai++) if ((todo >> ai) & 1) work.item[at++] = ((work.slot_base + slot) << 2) | ai;
        }
    }
}

#ifndef VD_WSC_WALK_MINB
#define VD_WSC_WALK_MINB 1
#endif
__global__ void __launch_bounds__(128, VD_WSC_WALK_MINB)
wsc_walk_kernel(BatchDev in, OutDev out, const ScPlan *__restrict__ plan, const int *__restrict__ order, const __grid_constant__ WscGroups G, u8 *slab,
                const WscHdr *__restrict__ hdr, const int *__restrict__ work_count, const int *__restrict__ work_item) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= *work_count) return;                               // everything else the sweep kernels have written already
    const int slot = work_item[g] >> 2, ai = work_item[g] & 3;
    const WscSlot ws = wsc_slot(G, slot, order, slab);
    const bool HOM = ws.hom;
    const int sc = ws.sc;
    const ScPlan p = plan[sc];
    const WscLayout M = wsc_layout(p, HOM);
    u8 *base = ws.base;
    const int64_t oi = 4 * (int64_t)sc + ai;
    const int qh = ai >> 1, th = 2 + (ai & 1);
    const int Lr = p.lr, Lq = p.len[qh], Lt = p.len[th], N = Lq + Lr;
    const u8 *qs = base + M.hap[qh], *ts = base + M.hap[th], *qb = base + M.qm[qh];
    Hap<int8_t> q{Lq, qs, qs + Lq, (const int8_t *)(qs + 2 * Lq), qs + 3 * Lq};
    Hap<int8_t> t{Lt, ts, ts + Lt, (const int8_t *)(ts + 2 * Lt), ts + 3 * Lt};
    QMaps<int8_t> qm{(const int8_t *)qb, qb + Lr, (const int8_t *)(qb + 2 * Lr), (const int8_t *)(qb + 2 * Lr + (Lq + 1 + Lr))};
    // path and Levenshtein scratch in thread-local memory (L1-cached, written and read back by this thread only): in the
    // slot they would cost an L2 round trip per entry of the credit loop
    constexpr int NPMAX = 32 * WSC_MAXSLOT + WSC_MAXLEN + 4, MNMAX = WSC_MAXLEN + 1;
    __align__(4) u8 scratch[2 * ((2 * NPMAX + 3) & ~3) + ((NPMAX + 3) & ~3) + ((2 * MNMAX + 3) & ~3)];
    GMemIL mem{scratch};
    AlnLayout<int> L;
    const int np = N + Lt + 4;
    L.oPF = L.oF = L.oD0 = L.oD1 = L.oT0 = L.oT1 = 0;
    L.oPQ = 0; L.oPT = wa4(2 * np); L.oPS = 2 * wa4(2 * np); L.oLev = L.oPS + wa4(np); L.total = 0;
    PFWarp pfr{base + M.F[ai], N, Lq};
    u32 status = out.status[oi];
    const int end_plane = out.aln_end_plane[oi], beg_plane = out.aln_beg_plane[oi];
    walk_credit<GMemIL, 2, int8_t>(mem, L, pfr, q, qm, t, base + M.rseq, Lr, beg_plane, end_plane, in, out, sc, ai, status);
    for (int k = 0; k < (HOM ? 4 : 1); k++) out.status[oi + k] = status;
    if (HOM) replicate_hom(in, out, sc);
}

#ifndef VD_WSC_PAR_MINBIN
#define VD_WSC_PAR_MINBIN 2
#endif
#ifndef VD_WSC_PAR_MINBIN_S2
#define VD_WSC_PAR_MINBIN_S2 VD_WSC_PAR_MINBIN
#endif
constexpr int WSC_PAR_MINBIN_S2 = VD_WSC_PAR_MINBIN_S2;   // the same threshold for two or more register slots
constexpr int WSC_PAR_MINBIN = VD_WSC_PAR_MINBIN;      // bins >= 10 KB per supercluster: one block per supercluster, one warp per alignment
#ifndef VD_WSC_SHARED
#define VD_WSC_SHARED 1
#endif
#ifndef VD_WSC_NW
#define VD_WSC_NW 4
#endif
constexpr int WSC_NW = VD_WSC_NW;      // superclusters (warps) per block of wsc_block_kernel
constexpr int WSC_SMEM_MAX = 227 * 1024;
template <int S> inline void wsc_configure_one() {
    const int shmax = WSC_NW * wsc_bin_cap(N_WBIN - 1) < WSC_SMEM_MAX ? WSC_NW * wsc_bin_cap(N_WBIN - 1) : WSC_SMEM_MAX;
    cudaFuncSetAttribute(wsc_block_kernel<S, false, WSC_NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, shmax);
    cudaFuncSetAttribute(wsc_block_kernel<S, true, WSC_NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, shmax);
    cudaFuncSetAttribute(wsc_kernel<S, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (WSC_TPB / 32) * wsc_bin_cap(WSC_PAR_MINBIN > 0 ? WSC_PAR_MINBIN - 1 : 0));
    cudaFuncSetAttribute(wsc_kernel<S, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, wsc_bin_cap(N_WBIN - 1));
    cudaFuncSetAttribute(wsc_kernel<S, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (WSC_TPB / 32) * wsc_bin_cap(N_WBIN - 1));
}
inline void wsc_configure() { wsc_configure_one<1>(); wsc_configure_one<2>(); wsc_configure_one<3>(); wsc_configure_one<4>(); }
template <int S> inline void wsc_launch_one(cudaStream_t st, int bin, bool hom, const BatchDev &in, const OutDev &out, const ScPlan *plan,
                                            const int *order, int count) {
    const int wb = wsc_bin_cap(bin), wpb = WSC_TPB / 32;
    const bool par = bin >= (S >= 2 ? WSC_PAR_MINBIN_S2 : WSC_PAR_MINBIN);
    if (VD_WSC_SHARED && (hom || !par) && WSC_NW * wb <= WSC_SMEM_MAX) {
        auto kb = hom ? wsc_block_kernel<S, true, WSC_NW> : wsc_block_kernel<S, false, WSC_NW>;
        VD_LAUNCH(kb, (count + WSC_NW - 1) / WSC_NW, 32 * WSC_NW, WSC_NW * wb, st, in, out, plan, order, count, wb);
    } else if (hom) {
        auto kh = wsc_kernel<S, false, true>;
        VD_LAUNCH(kh, (count + wpb - 1) / wpb, WSC_TPB, wpb * wb, st, in, out, plan, order, count, wb);
    } else if (par) {
        auto kp = wsc_kernel<S, true, false>;
        VD_LAUNCH(kp, count, WSC_TPB, wb, st, in, out, plan, order, count, wb);
    } else {
        auto kw = wsc_kernel<S, false, false>;
        VD_LAUNCH(kw, (count + wpb - 1) / wpb, WSC_TPB, wpb * wb, st, in, out, plan, order, count, wb);
    }
}
template <int S> inline void wsc_split_configure_one() {
    cudaFuncSetAttribute(wsc_sweep_kernel<S, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, wsc_bin_cap(N_WBIN - 1));
    cudaFuncSetAttribute(wsc_sweep_kernel<S, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, wsc_bin_cap(N_WBIN - 1));
    const int shmax = (WSC_TPB / 32) * wsc_bin_cap(N_WBIN - 1) < WSC_SMEM_MAX ? (WSC_TPB / 32) * wsc_bin_cap(N_WBIN - 1) : WSC_SMEM_MAX;
    cudaFuncSetAttribute(wsc_sweep_warp_kernel<S, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, shmax);
    cudaFuncSetAttribute(wsc_sweep_warp_kernel<S, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, shmax);
}
inline void wsc_split_configure() { wsc_split_configure_one<1>(); wsc_split_configure_one<2>(); wsc_split_configure_one<3>(); wsc_split_configure_one<4>(); }
// split form, sweeps of one group over its slab (count slots of wsc_bin_cap(bin) bytes); the expansion before and the
// walk after run once over all groups (wsc_expand_launch / wsc_walk_launch)
template <int S> inline void wsc_split_launch_one(cudaStream_t st, int bin, bool hom, const BatchDev &in, const OutDev &out, const ScPlan *plan,
                                                  const int *order, int count, u8 *slab, WscHdr *hdr, WscWork work) {
    const int wb = wsc_bin_cap(bin), wpb = WSC_TPB / 32;
    const bool par = !hom && bin >= (S >= 2 ? WSC_PAR_MINBIN_S2 : WSC_PAR_MINBIN);       // big bins: a block per supercluster
    const bool fits = wpb * wb <= WSC_SMEM_MAX;
    if (hom) {
        if (fits) {
            auto kw = wsc_sweep_warp_kernel<S, true>;
            VD_LAUNCH(kw, (count + wpb - 1) / wpb, WSC_TPB, wpb * wb, st, in, out, plan, order, count, slab, wb, (const WscHdr *)hdr, work);
        } else {
            auto ks = wsc_sweep_kernel<S, true>;
            VD_LAUNCH(ks, count, 32, wb, st, in, out, plan, order, count, slab, wb, (const WscHdr *)hdr, work);
        }
    } else {
        if (!par && fits) {
            auto kw = wsc_sweep_warp_kernel<S, false>;
            VD_LAUNCH(kw, (count + wpb - 1) / wpb, WSC_TPB, wpb * wb, st, in, out, plan, order, count, slab, wb, (const WscHdr *)hdr, work);
        } else {
            auto ks = wsc_sweep_kernel<S, false>;
            VD_LAUNCH(ks, count, 128, wb, st, in, out, plan, order, count, slab, wb, (const WscHdr *)hdr, work);
        }
    }
}
inline void wsc_split_launch(cudaStream_t st, int slots, int bin, bool hom, const BatchDev &in, const OutDev &out, const ScPlan *plan,
                             const int *order, int count, u8 *slab, WscHdr *hdr, WscWork work) {
    if (count <= 0) return;
    switch (slots) {
        case 1: wsc_split_launch_one<1>(st, bin, hom, in, out, plan, order, count, slab, hdr, work); break;
        case 2: wsc_split_launch_one<2>(st, bin, hom, in, out, plan, order, count, slab, hdr, work); break;
        case 3: wsc_split_launch_one<3>(st, bin, hom, in, out, plan, order, count, slab, hdr, work); break;
        case 4: wsc_split_launch_one<4>(st, bin, hom, in, out, plan, order, count, slab, hdr, work); break;
    }
}
inline void wsc_expand_launch(cudaStream_t st, const BatchDev &in, const ScPlan *plan, const int *order, const WscGroups &G, u8 *slab, WscHdr *hdr) {
    const int n = G.first[G.n];
    if (n > 0) VD_LAUNCH(wsc_expand_kernel, (4 * n + 127) / 128, 128, 0, st, in, plan, order, G, slab, hdr);
}
inline void wsc_walk_launch(cudaStream_t st, const BatchDev &in, const OutDev &out, const ScPlan *plan, const int *order, const WscGroups &G, u8 *slab,
                            const WscHdr *hdr, const int *work_count, const int *work_item) {
    const int n = G.first[G.n];          // the grid covers the worst case (every alignment to be walked); the count is on the device
    if (n > 0) VD_LAUNCH(wsc_walk_kernel, (4 * n + 127) / 128, 128, 0, st, in, out, plan, order, G, slab, hdr, work_count, work_item);
}
inline void wsc_launch(cudaStream_t st, int slots, int bin, bool hom, const BatchDev &in, const OutDev &out, const ScPlan *plan,
                       const int *order, int count) {
    if (count <= 0) return;
    switch (slots) {
        case 1: wsc_launch_one<1>(st, bin, hom, in, out, plan, order, count); break;
        case 2: wsc_launch_one<2>(st, bin, hom, in, out, plan, order, count); break;
        case 3: wsc_launch_one<3>(st, bin, hom, in, out, plan, order, count); break;
        case 4: wsc_launch_one<4>(st, bin, hom, in, out, plan, order, count); break;
    }
}

}  // namespace vd
