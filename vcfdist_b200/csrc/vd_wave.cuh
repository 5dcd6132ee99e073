// Wavefront kernels for superclusters that do not fit the fused shared-memory kernel.
//
// One thread block per alignment.  Rows of both planes (QUERY plane, padded to a multiple of
// K, then REF plane) are dealt K consecutive rows per thread; the kernel sweeps the truth
// haplotype one column at a time:
//
//   forward   base[a] = min(diag, del, swap) from column c-1 (own rows in registers, the
//             previous column of BOTH planes in shared memory for neighbour and swap reads),
//             then the insertion chain D[a] = min(base[a], D[a-1]+1) as a min-plus prefix scan
//             (thread-local, warp shuffles, one shared-memory hop across warps).  One flag
//             byte per cell is written to HBM, K bytes per thread, vectorised and coalesced.
//   backward  reverse sweep: T[y] = max over optimal successors (gather form), the in-column
//             INS chain as a max-plus suffix scan; the path-flag byte overwrites the forward
//             flag byte in place.
//   walk      one thread per alignment: path walk, sync sections, Levenshtein, integer credit
//             (walk_credit of vd_scalar.cuh) reading the path flags from HBM.
//
// HBM traffic is 3 B/cell (forward write, backward read + write); everything else is
// O(rows + columns).  Reference: src/dist.cpp:251-443 (forward), :486-834 (backward),
// :842-1401 (walk + credit).
#pragma once
#include "vd_kernels.cuh"

namespace vd {

constexpr int kBigClass = CLS_WAVE;
constexpr int FWD_TAU0 = 96;       // first score bound of the register-blocked forward sweep (then x4 per retry)

// ---- per-hap tables the wavefront kernels need, built once per supercluster --------------
// srcinfo: bit0 valid, bits1-3 k (index in the destination's source list), bit4 tp(dest),
//          bits 8.. destination row (plane-local)
template <class PT>
__device__ inline void build_srcinfo(const PT *ptr, const u8 *flg, int nsrc,       // sources
                                     const PT *tab, int ndst,                     // dest CSR
                                     const u8 *dflg, const PT *dptr,              // dest plane flags / hap->ref ptrs (tp), dptr null for REF dest
                                     int *info) {
    int n = 0;      // running position in the CSR source array (ascending b)
    for (int b = 0; b < nsrc; b++) {
        const int f = flg[b];
        int v = 0;
        const bool ok = !(f & P_VARIANT) || (f & P_VAR_END);
        const int d = (int)ptr[b] + 1;
        if (ok && d >= 0 && d < ndst) {
            // the n-th admitted source is CSR entry n (ascending b)
            const int k = n - (int)tab[d];
            n++;
            const int df = dflg[d];
            if (d > 0 && (!(df & P_VARIANT) || (df & P_VAR_BEG))) {               // :600-602, :638-640
                int tp = 0;
                if (dptr) tp = ((int)dptr[d] != (int)dptr[d - 1] + 1) || (df & P_VAR_BEG);  // :656-658
                v = 1 | (k << 1) | (tp << 4) | (d << 8);
            }
        }
        info[b] = v;
    }
}

// Packed per-row records of the banded warp kernels (vd_band.cuh), 16 bytes each so that 32 rows are one
// 512-byte bulk copy into a warp's shared-memory ring.
struct FwdRow {          // forward sweep
    u32 w0;              // my row as a swap DESTINATION: first source row of the other plane (bits 0-19) | number of
                         // sources (bits 20-23) | base of the row (bits 24-31)
    int fdest;           // my row as a swap SOURCE of the forward sweep: ptr+1 on the other plane, -1: none (:335-337, :364-366)
    int hlo, hhi;        // fewest / most row steps from the row to the end of either plane (wave_row_hulls)
};
struct BwdRow {          // backward sweep
    u32 si;              // my row as the recorded swap source: srcinfo (bit 0 valid, bits 1-3 k, bit 4 tp(dest), bits 8.. destination row)
    u32 tw;              // rows a' > a of my plane with tp(a') (bits 0-23) | tp(a) << 24 | tp(a+1) << 25
    u32 chn;             // base of row a+1 (0xff: none)
    u32 pad;
};
constexpr int BAND_ROW_LIMIT = 1 << 20;      // rows per plane the packed records can address

struct WaveHapQ {        // extra per query hap
    int *srcQ;           // [Lq]  QUERY rows as swap sources (destinations on the REF plane)
    int *srcR;           // [Lr]  REF rows as swap sources (destinations on the QUERY plane)
    u32 *swiQ;           // [Lq]  QUERY rows as swap DESTINATIONS: first source (REF row) | count << 16
    u32 *swiR;           // [Lr]  REF rows as swap DESTINATIONS: first source (QUERY row) | count << 16
    u8 *tpb;             // [Lq]  tp(a): entering QUERY row a counts a query variant (:572-574)
    u16 *tps;            // [Lq]  number of rows a' > a with tp(a') (potential of the backward insertion chain)
    FwdRow *fwdQ, *fwdR; // [Lq], [Lr] packed records of the banded warp kernels; every array is padded to whole
    BwdRow *bwdQ, *bwdR; //            32-row chunks (the unit of the bulk copies)
    __device__ WaveHapQ(u8 *base, int Lq, int Lr) {
        const int64_t pq = align_up(Lq, 32), pr = align_up(Lr, 32);
        fwdQ = (FwdRow *)base; fwdR = fwdQ + pq;
        bwdQ = (BwdRow *)(fwdR + pr); bwdR = bwdQ + pq;
        srcQ = (int *)(bwdR + pr); srcR = srcQ + Lq; swiQ = (u32 *)(srcR + Lr); swiR = swiQ + Lq; tpb = (u8 *)(swiR + Lr);
        tps = (u16 *)(tpb + align_up(Lq, 16));
    }
};
__host__ __device__ inline int64_t wave_hapq_bytes(int Lq, int Lr) { return 32 * (align_up(Lq, 32) + align_up(Lr, 32)) + 8 * ((int64_t)Lq + Lr) + align_up(Lq, 16) + align_up(2 * (int64_t)Lq, 16); }
__host__ __device__ inline int64_t wave_hapt_bytes(int Lt) { return align_up(Lt, 16); }     // tinfo: base | tok<<7

// kernel shape classes: (threads per block, rows per thread)
constexpr int N_WCLS = 8;
__host__ __device__ inline int wave_tpb(int c) { const int v[N_WCLS] = {32, 32, 32, 128, 256, 512, 1024, 1024}; return v[c]; }
__host__ __device__ inline int wave_k(int c) { const int v[N_WCLS] = {1, 2, 4, 4, 8, 16, 16, 32}; return v[c]; }
__host__ __device__ inline int wave_class(int Lq, int Lr, int Lt) {
    for (int c = 0; c < N_WCLS; c++) {
        const int K = wave_k(c);
        const int np = (Lq + K - 1) / K * K + (Lr + K - 1) / K * K;
        if (np <= wave_tpb(c) * K && np + Lt < 65000) return c;
    }
    return -1;
}
// item lists: one per shape class, plus one for the alignments no block kernel takes (more than 32768 rows over
// both planes, or a score that could overflow 16 bits): banded warp kernels first, thread-per-alignment kernel else
constexpr int N_WLIST = N_WCLS + 1;
__host__ __device__ inline int wave_list_of(int Lq, int Lr, int Lt) { const int c = wave_class(Lq, Lr, Lt); return c < 0 ? N_WCLS : c; }
// banded flag storage (vd_band.cuh): smallest rung (rows per lane) whose window of 32*K rows holds both planes
// entirely (0: none), and the widest rung an alignment can reach - its storage is sized for that one
__host__ __device__ inline int band_fit_k(int Lq, int Lr) {
    const int m = Lq > Lr ? Lq : Lr;
    return m <= 128 ? 4 : (m <= 256 ? 8 : (m <= 512 ? 16 : 0));
}
__host__ __device__ inline int band_kmax(int Lq, int Lr) { const int k = band_fit_k(Lq, Lr); return k ? k : 16; }

// scratch of one alignment in the slab
struct WaveAln {
    int64_t oF;          // BANDED flags [Lt][2][32*kmax] (vd_band.cuh); the dense matrix [Lt][NP] of the block kernels
                         // lives in a separate buffer that only the alignments the banded kernels give up on get
    int64_t oBand;       // int4[Lt]: candidate rows of every column (banded kernels)
    int64_t oWalk;       // walk scratch (path + Levenshtein row), AlnLayout offsets relative to it
    int64_t total;
    int cls, K, padQ, NP, kmax;
};
__host__ __device__ inline WaveAln wave_aln(int Lq, int Lr, int Lt) {
    WaveAln w;
    w.cls = wave_class(Lq, Lr, Lt);
    w.K = w.cls >= 0 ? wave_k(w.cls) : 1;
    w.padQ = (Lq + w.K - 1) / w.K * w.K;
    w.NP = w.padQ + (Lr + w.K - 1) / w.K * w.K;      // both planes padded to whole threads
    w.kmax = band_kmax(Lq, Lr);
    w.oF = 0;
    w.oBand = align_up((int64_t)Lt * 64 * w.kmax, 16);
    w.oWalk = w.oBand + 16 * (int64_t)Lt;
    // path (int32 q, int32 t, u8 flags) + lev row
    const int np = Lq + Lr + Lt + 4;
    const int mn = (Lr < Lt ? Lr : Lt) + 1;
    int64_t walk = 8 * (int64_t)np + align_up(np, 4) + 4 * (int64_t)mn;
    const int64_t lists = 16 * (int64_t)w.NP + 64;       // sparse backward: 2 frontier lists + 2 worklists (int32)
    if (lists > walk && w.cls >= 0) walk = lists;
    w.total = align_up(w.oWalk + walk, 16);
    return w;
}
// bytes of the dense-phase scratch of one alignment: the flag matrix of the block kernels, or everything the
// thread-per-alignment kernel needs when no block kernel takes the shape
__host__ __device__ inline int64_t wave_dense_bytes(int Lq, int Lr, int Lt) {
    const WaveAln w = wave_aln(Lq, Lr, Lt);
    if (w.cls < 0) return align_up(make_layout<int64_t, 4, false>(Lq + Lr, Lt, Lr).total, 256);
    return align_up((int64_t)w.NP * Lt, 256);
}
// walk scratch offsets in AlnLayout form (only the path / lev members are used)
__device__ inline AlnLayout<int64_t> wave_walk_layout(int Lq, int Lr, int Lt) {
    AlnLayout<int64_t> L;
    const int np = Lq + Lr + Lt + 4;
    L.oPF = L.oF = L.oD0 = L.oD1 = L.oT0 = L.oT1 = 0;
    L.oPQ = 0;
    L.oPT = 4 * (int64_t)np;
    L.oPS = 8 * (int64_t)np;
    L.oLev = L.oPS + align_up(np, 4);
    L.total = 0;
    return L;
}

// slab layout of a wave-class supercluster: the scalar layout's hap/qm regions, then the
// wave tables, then per-alignment scratch
struct WaveSlab {
    SlabLayout base;
    int64_t hq[2], ht[2], aln[4];
    int64_t total;
};
__host__ __device__ inline WaveSlab make_wave_slab(const ScPlan &p) {
    WaveSlab w;
    w.base = make_slab(p, false);
    int64_t o = w.base.total;
    for (int k = 0; k < 2; k++) { w.hq[k] = o; o = align_up(o + wave_hapq_bytes(p.len[k], p.lr), 16); }
    for (int k = 0; k < 2; k++) { w.ht[k] = o; o = align_up(o + wave_hapt_bytes(p.len[2 + k]), 16); }
    for (int ai = 0; ai < 4; ai++) {
        w.aln[ai] = o;
        o += wave_aln(p.len[ai >> 1], p.lr, p.len[2 + (ai & 1)]).total;
    }
    w.total = o;
    return w;
}

__global__ void wave_size_kernel(const ScPlan *plan, const int *list, int n, int64_t *bytes) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const ScPlan p = plan[list[i]];
    bytes[i] = p.cls == CLS_WAVE ? make_wave_slab(p).total : make_slab(p, true).total;
}

// Row-step hulls for the score-bounded sweeps.  Forget the truth for a moment and look at the ROW graph of a
// query haplotype: QUERY row a -> a+1, REF row r -> r+1, and the swap edges (source row -> ptr+1 on the other
// plane, :335-337, :364-366).  A path from a cell to the end of either plane takes k row steps for some k in
// [lo, hi] = the shortest / longest path in that graph; every row step that is not an insertion consumes a
// truth base and every deletion consumes a truth base without a row step, so with rc truth bases left the
// remaining cost is at least max(0, rc - hi, lo - rc).  The bound is consistent along every edge of the
// two-plane graph (an edge taking dr row steps and dc truth bases costs at least |dr - dc|, except diagonal
// and swap edges with dr = dc), so a sweep that only keeps cells with D + bound <= tau still gives every kept
// cell its exact distance and its complete set of optimal predecessors - and it drops, for instance, the cells
// that reach the far side of a 10 kb insertion through the REF plane at no cost but can never finish.
// Computed backwards by merging the two planes so that every swap destination is done before its source.
__device__ inline void wave_row_hulls(const int *qptr, const u8 *qflg, int Lq, const int *rptr, const u8 *rflg, int Lr,
                                      FwdRow *hullQ, FwdRow *hullR) {
    int a = Lq - 1, r = Lr - 1;
    while (a >= 0 || r >= 0) {
        bool okQ = false, okR = false, swQ = false, swR = false;
        int dQ = -1, dR = -1;
        if (a >= 0) {
            const int f = qflg[a];
            dQ = qptr[a] + 1;
            swQ = (!(f & P_VARIANT) || (f & P_VAR_END)) && dQ >= 0 && dQ < Lr;
            okQ = !swQ || dQ > r;
        }
        if (r >= 0) {
            const int f = rflg[r];
            dR = rptr[r] + 1;
            swR = (!(f & P_VARIANT) || (f & P_VAR_END)) && dR >= 0 && dR < Lq;
            okR = !swR || dR > a;
        }
        const bool takeQ = okQ || (!okR && a >= 0);
        if (takeQ) {
            int lo, hi;
            if (a == Lq - 1) lo = hi = 0; else { lo = hullQ[a + 1].hlo + 1; hi = hullQ[a + 1].hhi + 1; }
            if (swQ) {
                if (dQ > r) { lo = min(lo, hullR[dQ].hlo + 1); hi = max(hi, hullR[dQ].hhi + 1); }
                else { lo = 0; hi = INF; }                     // cannot happen in a DAG; no bound rather than a wrong one
            }
            hullQ[a].hlo = lo; hullQ[a].hhi = hi; hullQ[a].fdest = swQ ? dQ : -1;
            a--;
        } else {
            int lo, hi;
            if (r == Lr - 1) lo = hi = 0; else { lo = hullR[r + 1].hlo + 1; hi = hullR[r + 1].hhi + 1; }
            if (swR) {
                if (dR > a) { lo = min(lo, hullQ[dR].hlo + 1); hi = max(hi, hullQ[dR].hhi + 1); }
                else { lo = 0; hi = INF; }
            }
            hullR[r].hlo = lo; hullR[r].hhi = hi; hullR[r].fdest = swR ? dR : -1;
            r--;
        }
    }
}

// one thread per (entry, hap): wave tables after slab_setup_kernel's expansion
__global__ void wave_tables_kernel(BatchDev in, const ScPlan *plan, const int *list, int i0, int i1,
                                   const int64_t *offs, u8 *slab, const int *hap_ok) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = i0 + (g >> 2), h = g & 3;
    if (i >= i1) return;
    const u8 *rseq = in.rplane_seq + in.ref_off[list[i]];
    const ScPlan p = plan[list[i]];
    if (p.cls != CLS_WAVE) return;
    if (p.hom && (h & 1)) return;                       // homozygous: haplotypes 1 / 3 equal 0 / 2 and are never used
    const int *okp = hap_ok + 4 * (int64_t)(i - i0);
    if (!(okp[0] && okp[1] && okp[2] && okp[3])) return;
    const WaveSlab W = make_wave_slab(p);
    u8 *base = slab + (offs[i] - offs[i0]);
    SlabHap H(base + W.base.hap[h], p.len[h], p.lr);
    if (h < 2) {
        SlabQm M(base + W.base.qm[h], p.len[h], p.lr);
        WaveHapQ X(base + W.hq[h], p.len[h], p.lr);
        // QUERY rows as sources -> destinations on REF (CSR toR); dest flags = rflg, no tp on REF
        build_srcinfo<int>(H.ptr, H.flg, p.len[h], M.toR, p.lr, M.rflg, nullptr, X.srcQ);
        // REF rows as sources -> destinations on QUERY (CSR toQ); dest flags = hap flags, tp from hap ptrs
        build_srcinfo<int>(M.rptr, M.rflg, p.lr, M.toQ, p.len[h], H.flg, H.ptr, X.srcR);
        for (int a = 0; a < p.len[h]; a++)
            X.tpb[a] = (a > 0 && ((H.ptr[a] != H.ptr[a - 1] + 1) || (H.flg[a] & P_VAR_BEG))) ? 1 : 0;
        {
            int cnt = 0;
            for (int a = p.len[h] - 1; a >= 0; a--) { X.tps[a] = (u16)cnt; cnt += X.tpb[a]; }
        }
        for (int a = 0; a < p.len[h]; a++) {
            const int k0 = M.toQ[a], k1 = M.toQ[a + 1];
            X.swiQ[a] = k1 > k0 ? ((u32)M.toQ[p.len[h] + 1 + k0] | ((u32)(k1 - k0) << 16)) : 0u;
        }
        for (int a = 0; a < p.lr; a++) {
            const int k0 = M.toR[a], k1 = M.toR[a + 1];
            X.swiR[a] = k1 > k0 ? ((u32)M.toR[p.lr + 1 + k0] | ((u32)(k1 - k0) << 16)) : 0u;
        }
        // packed row records of the banded warp kernels
        const int Lq = p.len[h];
        wave_row_hulls(H.ptr, H.flg, Lq, M.rptr, M.rflg, p.lr, X.fwdQ, X.fwdR);
        {
            int cnt = 0;
            for (int a = Lq - 1; a >= 0; a--) {
                const int k0 = M.toQ[a], k1 = M.toQ[a + 1];
                X.fwdQ[a].w0 = (k1 > k0 ? (((u32)M.toQ[Lq + 1 + k0] & 0xfffffu) | ((u32)(k1 - k0) << 20)) : 0u) | ((u32)(H.str[a] & 0x7f) << 24);
                BwdRow r;
                r.si = (u32)X.srcQ[a];
                const u32 tpa = X.tpb[a], tpn = a + 1 < Lq ? X.tpb[a + 1] : 0;
                r.tw = (u32)cnt | (tpa << 24) | (tpn << 25);
                r.chn = a + 1 < Lq ? (u32)(H.str[a + 1] & 0x7f) : 0xffu;
                r.pad = 0;
                X.bwdQ[a] = r;
                cnt += (int)tpa;
            }
        }
        for (int a = 0; a < p.lr; a++) {
            const int k0 = M.toR[a], k1 = M.toR[a + 1];
            X.fwdR[a].w0 = (k1 > k0 ? (((u32)M.toR[p.lr + 1 + k0] & 0xfffffu) | ((u32)(k1 - k0) << 20)) : 0u) | ((u32)(rseq[a] & 0x7f) << 24);
            BwdRow r;
            r.si = (u32)X.srcR[a];
            r.tw = 0;                                    // tp is zero on the REF plane
            r.chn = a + 1 < p.lr ? (u32)(rseq[a + 1] & 0x7f) : 0xffu;
            r.pad = 0;
            X.bwdR[a] = r;
        }
    } else {
        u8 *tinfo = base + W.ht[h - 2];
        for (int c = 0; c < p.len[h]; c++) {
            const bool tok = c > 0 && (!(H.flg[c - 1] & P_VARIANT) || (H.flg[c - 1] & P_VAR_END));   // :338-339
            tinfo[c] = (u8)((H.str[c] & 0x7f) | (tok ? 0x80 : 0));
        }
    }
}

// class-sorted list of (entry, alignment) items of one chunk
struct WaveItems {
    int count[N_WLIST];
    int cursor[N_WLIST];
    int n_dense;                 // items left to the dense phase once the banded kernels are through
    int pad_;
    unsigned long long spill_cells;
    unsigned long long dense_bytes;
    unsigned long long band_cells, band_rows, band_cols;   // banded kernels: candidate cells, rows and columns of the solved alignments
};
struct ClsBase { int b[N_WLIST]; };
__global__ void wave_count_kernel(const ScPlan *plan, const int *list, int i0, int i1, const int *hap_ok,
                                  WaveItems *wi) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = i0 + (g >> 2), ai = g & 3;
    if (i >= i1) return;
    const ScPlan p = plan[list[i]];
    if (p.cls != CLS_WAVE) return;
    if (p.hom && ai) return;                             // homozygous: alignment 0 only, records replicated afterwards
    const int *okp = hap_ok + 4 * (int64_t)(i - i0);
    if (!(okp[0] && okp[1] && okp[2] && okp[3])) return;
    atomicAdd(&wi->count[wave_list_of(p.len[ai >> 1], p.lr, p.len[2 + (ai & 1)])], 1);
}
__device__ inline int band_score_lb(const BatchDev &in, int sc, int ai, int Lr, int Lt);
// state[]: 0 = pending for the banded kernels (band_on) or -1 = dense phase; lbound[]: lower bound of the score
__global__ void wave_fill_kernel(BatchDev in, const ScPlan *plan, const int *list, int i0, int i1, const int *hap_ok,
                                 WaveItems *wi, ClsBase cb, int *items, OutDev out, int *state, int *lbound, int *hint, int band_on) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = i0 + (g >> 2), ai = g & 3;
    if (i >= i1) return;
    const int sc = list[i];
    const ScPlan p = plan[sc];
    if (p.cls != CLS_WAVE) return;
    const int *okp = hap_ok + 4 * (int64_t)(i - i0);
    if (!(okp[0] && okp[1] && okp[2] && okp[3])) {
        out.status[4 * (int64_t)sc + ai] = ST_BAD;
        out.aln_score[4 * (int64_t)sc + ai] = -1;
        return;
    }
    if (p.hom && ai) return;
    const int Lq = p.len[ai >> 1], Lt = p.len[2 + (ai & 1)];
    const int pos = cb.b[wave_list_of(Lq, p.lr, Lt)] + atomicAdd(&wi->cursor[wave_list_of(Lq, p.lr, Lt)], 1);
    items[pos] = ((i - i0) << 2) | ai;
    state[pos] = band_on ? 0 : -1;
    lbound[pos] = band_score_lb(in, sc, ai, p.lr, Lt);
    hint[pos] = 0;
}

// dense phase: scratch bytes of every item the banded kernels did not solve (0 for the others)
__global__ void wave_dense_size_kernel(const ScPlan *plan, const int *list, int i0, const int *items, const int *state, int n_items,
                                       int64_t *bytes, WaveItems *wi) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_items) return;
    int64_t b = 0;
    if (state[idx] <= 0) {
        const int item = items[idx];
        const ScPlan p = plan[list[i0 + (item >> 2)]];
        const int ai = item & 3;
        b = wave_dense_bytes(p.len[ai >> 1], p.lr, p.len[2 + (ai & 1)]);
        atomicAdd(&wi->n_dense, 1);
        atomicAdd(&wi->dense_bytes, (unsigned long long)b);
        const WaveAln wa = wave_aln(p.len[ai >> 1], p.lr, p.len[2 + (ai & 1)]);
        if (wa.cls >= 0) atomicAdd(&wi->spill_cells, (unsigned long long)wa.NP * p.len[2 + (ai & 1)]);
    }
    bytes[idx] = b;
}

struct WaveArgs {
    BatchDev in;
    OutDev out;
    const ScPlan *plan;
    const int *list;
    int i0;
    const int64_t *offs;
    u8 *slab;
    const int *items;
    const int *bstate;           // per item: > 0 solved by the banded kernels (vd_band.cuh): the dense kernels skip it
    u8 *dense;                   // dense-phase scratch: flag matrices of the block kernels
    const int64_t *dense_off;    // per item, byte offset into dense
};

// everything a block needs about its alignment
template <class TT> struct WaveCtxT {      // TT: element type of the CSR swap tables (int in HBM, short in smem)
    int sc, ai, Lq, Lr, Lt, padQ, NP;
    const u8 *qstr, *rseq, *tinfo, *qflg, *rflg, *tpb;
    const u16 *tps;
    short *band;                // [Lt][4] rows visited by the banded forward sweep (walk scratch, dead until the walk)
    const TT *toQ, *toR;
    const TT *qptr, *rptr;      // query->ref and ref->query pointers (band bookkeeping of the forward sweep)
    const int *srcQ, *srcR;
    const u32 *swiQ, *swiR;     // per destination row: first swap source | count << 16 (slab path only)
    u8 *F;                      // dense flag matrix [Lt][NP] (dense-phase scratch)
    u8 *walk;                   // the alignment's walk scratch in the slab
    bool skip;                  // solved by the banded kernels
};
typedef WaveCtxT<int> WaveCtx;
__device__ inline WaveCtx wave_ctx(const WaveArgs &A, int idx) {
    WaveCtx x;
    const int item = A.items[idx];
    x.skip = A.bstate && A.bstate[idx] > 0;            // solved (and walked) by the banded kernels
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
    x.padQ = wa.padQ; x.NP = wa.NP;
    SlabHap HQ(base + W.base.hap[qh], x.Lq, x.Lr);
    SlabQm M(base + W.base.qm[qh], x.Lq, x.Lr);
    WaveHapQ X(base + W.hq[qh], x.Lq, x.Lr);
    x.qstr = HQ.str; x.qflg = HQ.flg; x.rflg = M.rflg; x.tpb = X.tpb; x.tps = X.tps;
    x.toQ = M.toQ; x.toR = M.toR; x.srcQ = X.srcQ; x.srcR = X.srcR;
    x.qptr = HQ.ptr; x.rptr = M.rptr;
    x.swiQ = X.swiQ; x.swiR = X.swiR;
    x.rseq = A.in.rplane_seq + A.in.ref_off[x.sc];
    x.tinfo = base + W.ht[th];
    x.F = (A.dense && !x.skip) ? A.dense + A.dense_off[idx] : nullptr;
    x.walk = base + W.aln[x.ai] + wa.oWalk;
    x.band = (short *)x.walk;
    return x;
}

template <int K> __device__ __forceinline__ void store_flags(u8 *dst, const u32 *w) {
    if constexpr (K == 1) dst[0] = (u8)w[0];
    else if constexpr (K == 2) *(u16 *)dst = (u16)w[0];
    else if constexpr (K == 4) *(u32 *)dst = w[0];
    else if constexpr (K == 8) *(uint2 *)dst = make_uint2(w[0], w[1]);
    else {
#pragma unroll
        for (int i = 0; i < K / 16; i++) ((uint4 *)dst)[i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
    }
}
template <int K> __device__ __forceinline__ void load_flags(const u8 *src, u32 *w) {
    if constexpr (K == 1) w[0] = src[0];
    else if constexpr (K == 2) w[0] = *(const u16 *)src;
    else if constexpr (K == 4) w[0] = *(const u32 *)src;
    else if constexpr (K == 8) { const uint2 v = *(const uint2 *)src; w[0] = v.x; w[1] = v.y; }
    else {
#pragma unroll
        for (int i = 0; i < K / 16; i++) {
            const uint4 v = ((const uint4 *)src)[i];
            w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
        }
    }
}
__device__ __forceinline__ int byte_of(const u32 *w, int j) { return (w[j >> 2] >> ((j & 3) * 8)) & 0xff; }

template <int TPB> __device__ __forceinline__ void block_sync() {
    if constexpr (TPB == 32) __syncwarp(); else __syncthreads();
}

// ------------------------------------------------------------------------------------------
// forward: calc_prec_recall_aln (:251-443) as a column sweep
// ------------------------------------------------------------------------------------------
// t = thread index within the TPB-thread group working on this alignment; smem_raw = that group's
// scratch (wave_fwd_smem bytes); sEnd = two ints of shared scratch.  Returns score / end plane to
// every thread of the group.
template <int TPB, int K, class TT>
__device__ __forceinline__ void wave_fwd_body(const WaveCtxT<TT> &X, const int t, u8 *smem_raw, int *sEnd,
                                              int &score, int &end_plane, const int tau0 = FWD_TAU0) {
    constexpr int NPMAX = TPB * K;
    u16 *sD0 = (u16 *)smem_raw, *sD1 = sD0 + NPMAX;          // previous / current column, both planes
    int *sLive = (int *)(sD1 + NPMAX);                       // [2 columns][loQ, hiQ, loR, hiR] rows with D <= tau
    int *sW = sLive + 8;                                     // [2 segments][32 warps] scan totals
    const int lane = t & 31, warp = t >> 5;
    const int tQ = X.padQ / K;                               // first thread of the REF plane
    const bool P = t >= tQ;
    const int row0 = t * K;
    const int a0 = P ? row0 - X.padQ : row0;
    const int len = P ? X.Lr : X.Lq;
    const u8 *seq = P ? X.rseq : X.qstr;
    const TT *tab = P ? X.toR : X.toQ;                       // CSR of swap sources of my plane's rows
    const TT *src = tab + len + 1;
    const int obase = P ? 0 : X.padQ;                        // padded row offset of the other plane
    const int segstart = P ? tQ : 0;

    u8 ch[K];
#pragma unroll
    for (int j = 0; j < K; j++) ch[j] = (a0 + j < len) ? (u8)(seq[a0 + j] & 0x7f) : (u8)0xff;
    // swap sources of my rows, resolved once: first source as a padded row index of the other plane
    // (bits 0-15) and the number of sources (bits 16-19).  Almost every row has exactly one, so the
    // per-column swap candidate is a single shared-memory load instead of a chain of dependent
    // CSR loads from global memory (which put ~100 cycles per row on the column's critical path).
    u32 swi[K];
#pragma unroll
    for (int j = 0; j < K; j++) {
        swi[j] = 0;
        if (a0 + j < len) {
            const int k0 = tab[a0 + j], k1 = tab[a0 + j + 1];
            if (k1 > k0) swi[j] = (u32)(obase + (int)src[k0]) | ((u32)(k1 - k0) << 16);
        }
    }
    auto swap_eval = [&](const int j, const u16 *sPrev, int &best, int &sb) {
        const u32 w = swi[j];
        const int cnt = (int)(w >> 16);
        best = INF; sb = 0;
        if (cnt) {
            const int v0 = sPrev[w & 0xffffu];
            best = v0 == 0xffff ? INF : v0;
            if (cnt > 1) {                                     // rare: insertion / adjacent deletions
                const int k0 = tab[a0 + j];
                for (int k = 1; k < cnt; k++) {
                    int v = sPrev[obase + (int)src[k0 + k]];
                    v = v == 0xffff ? INF : v;
                    if (v < best) { best = v; sb = k << F_K_SHIFT; }
                    else if (v == best) sb = (k << F_K_SHIFT) | F_TIE;     // keep the larger row
                }
            }
        }
    };

    // Score-bounded sweep (Ukkonen): cells whose distance exceeds tau cannot lie on a path of cost
    // <= tau, so threads whose rows cannot hold such a cell skip the column.  Every cell with
    // D <= tau keeps its exact value and flags (its optimal predecessors have D <= tau as well), so
    // the result is exact whenever the final score is <= tau; otherwise tau is quadrupled and the
    // sweep repeated (worst case 1/3 extra work).  The last attempt has no bound at all.
    const int unbounded = X.NP + X.Lt + 1;
    for (int tau = (unbounded <= 4 * tau0 || tau0 >= unbounded) ? unbounded : tau0;; tau = (4 * tau >= unbounded) ? unbounded : 4 * tau) {
        int Dp[K];
#pragma unroll
        for (int j = 0; j < K; j++) Dp[j] = INF;
#pragma unroll
        for (int j = 0; j < K; j++) sD0[row0 + j] = 0xffff;
        if (t == 0) {
            sLive[0] = sLive[2] = sLive[4] = sLive[6] = INF;
            sLive[1] = sLive[3] = sLive[5] = sLive[7] = -1;
        }
        block_sync<TPB>();

        int tnext = X.tinfo[0];
        bool dead = false;
        for (int c = 0; c < X.Lt; c++) {
            u16 *sPrev = (c & 1) ? sD1 : sD0, *sCur = (c & 1) ? sD0 : sD1;
            const int tinfo = tnext;
            if (c + 1 < X.Lt) tnext = X.tinfo[c + 1];
            const int tch = tinfo & 0x7f;
            const bool tok = tinfo & 0x80;
            // ---- rows of my plane that can hold D <= tau in this column (plane-local, inclusive) ----
            int candLo, candHi;
            if (c == 0) { candLo = 0; candHi = tau; }
            else {
                const int *lv = sLive + ((c - 1) & 1) * 4;
                const int loQ = lv[0], hiQ = lv[1], loR = lv[2], hiR = lv[3];
                if (hiQ < loQ && hiR < loR) { dead = true; break; }        // nothing within tau is left
                int lo = INF, hi = -1;
                if (!P) {
                    if (hiQ >= loQ) { lo = loQ; hi = hiQ + 1; }                                  // del, diag
                    if (hiR >= loR) { lo = min(lo, (int)X.rptr[loR] + 1); hi = max(hi, (int)X.rptr[hiR] + 1); }   // swap dests
                } else {
                    if (hiR >= loR) { lo = loR; hi = hiR + 1; }
                    if (hiQ >= loQ) { lo = min(lo, (int)X.qptr[loQ] + 1); hi = max(hi, (int)X.qptr[hiQ] + 1); }
                }
                candLo = lo; candHi = hi + tau;                            // + the in-column INS chain
            }
            const bool act = a0 <= candHi && a0 + K - 1 >= candLo && a0 < len;
            // ---- pass 1: thread-local chain ----
            int up = INF;                                       // D[a0-1][c-1]
            int Dc[K];
            if (act) {
                {
                    const int v = (a0 > 0 && c > 0) ? sPrev[row0 - 1] : 0xffff;
                    up = v == 0xffff ? INF : v;
                }
                int run = INF;                                  // D[a-1][c] within the thread (no carry yet)
                int upj = up;
#pragma unroll
                for (int j = 0; j < K; j++) {
                    const bool m = ch[j] == tch;
                    int b;
                    if (a0 + j == 0 && c == 0) b = (a0 + j < len) ? 0 : INF;   // both origins start at 0 (:299-305)
                    else {
                        b = upj + (m ? 0 : 1);                  // diag (INF-safe: INF+1 stays huge)
                        b = min(b, Dp[j] + 1);                  // del
                        if (tok && m) {                         // swap (:334-349, :363-378)
                            int best, sb;
                            swap_eval(j, sPrev, best, sb);
                            b = min(b, best);
                        }
                        if (a0 + j >= len) b = INF;
                    }
                    run = min(b, run + 1);
                    Dc[j] = run;
                    upj = Dp[j];
                }
            } else {
#pragma unroll
                for (int j = 0; j < K; j++) Dc[j] = INF;
            }
            // ---- cross-thread min-plus prefix scan of G = last - lastrow ----
            int G = act ? Dc[K - 1] - (row0 + K - 1) : INF;
            int incl = G;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int o = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d && t - d >= segstart) incl = min(incl, o);
            }
            int excl = __shfl_up_sync(0xffffffffu, incl, 1);
            if (lane == 0 || t - 1 < segstart) excl = INF;
            int carry = excl;
            if constexpr (TPB > 32) {
                // warp totals per segment: the last lane's inclusive value covers the lanes of its own
                // segment; a warp straddling the plane boundary also publishes its QUERY-part total
                int *sWc = sW + (c & 1) * 64;
                const int lastQ = tQ - 1;                        // last QUERY thread
                if (lane == 31) sWc[(P ? 32 : 0) + warp] = incl;
                if (t == lastQ && lane != 31) sWc[warp] = incl;
                if (lane == 31 && !P) { /* pure QUERY warp: REF total empty */ sWc[32 + warp] = INF; }
                if (lane == 31 && P && (t - 31) >= tQ) sWc[warp] = INF;      // pure REF warp: QUERY total empty
                __syncthreads();
                // totals of earlier warps in my segment.  The slot choice must be warp-uniform for the
                // butterfly: a warp that starts on the QUERY plane reads QUERY totals; its REF lanes (a
                // straddling warp) have no earlier REF warp at all.
                const bool warpQ = 32 * warp < tQ;
                int wv = (lane < warp) ? sWc[(warpQ ? 0 : 32) + lane] : INF;
#pragma unroll
                for (int d = 16; d; d >>= 1) wv = min(wv, __shfl_xor_sync(0xffffffffu, wv, d));
                if (warpQ && P) wv = INF;
                carry = min(carry, wv);
            } else {
                __syncwarp();
            }
            // everyone has read sLive of column c-1: recycle that slot for column c+1
            if (t == 0) {
                int *nx = sLive + ((c + 1) & 1) * 4;
                nx[0] = nx[2] = INF; nx[1] = nx[3] = -1;
            }
            // ---- pass 2: final values and flags ----
            if (act) {
                u32 fw[(K + 3) / 4];
#pragma unroll
                for (int i = 0; i < (K + 3) / 4; i++) fw[i] = 0;
                // carry is min(G) = D_last - lastrow over earlier threads, so the chain value at row r is carry + r
                int prevD = carry >= INF / 2 ? INF : carry + (row0 - 1);           // D[a0-1][c]
                int upj = up;
                int mylo = INF, myhi = -1;
#pragma unroll
                for (int j = 0; j < K; j++) {
                    const int r = row0 + j;
                    int d = Dc[j];
                    if (carry < INF / 2) d = min(d, carry + r);
                    int f = 0;
                    if (a0 + j < len) {
                        if (a0 + j == 0 && c == 0) f = F_DIAG;
                        else {
                            const bool m = ch[j] == tch;
                            if (a0 + j > 0 && c > 0 && upj + (m ? 0 : 1) == d) f |= F_DIAG;
                            if (a0 + j > 0 && prevD + 1 == d) f |= F_INS;
                            if (c > 0 && Dp[j] + 1 == d) f |= F_DEL;
                            if (tok && m) {
                                int best, sb;
                                swap_eval(j, sPrev, best, sb);
                                if (best == d) f |= F_SWP | sb;
                            }
                        }
                        if (d <= tau) { mylo = min(mylo, a0 + j); myhi = a0 + j; }
                    } else d = INF;
                    fw[j >> 2] |= (u32)f << ((j & 3) * 8);
                    upj = Dp[j];
                    Dp[j] = d;
                    prevD = d;
                    sCur[r] = d >= 0xffff ? (u16)0xffff : (u16)d;
                }
                store_flags<K>(X.F + (int64_t)c * X.NP + row0, fw);
                if (myhi >= 0) {
                    int *lc = sLive + (c & 1) * 4 + (P ? 2 : 0);
                    atomicMin(&lc[0], mylo);
                    atomicMax(&lc[1], myhi);
                }
            } else {
#pragma unroll
                for (int j = 0; j < K; j++) { Dp[j] = INF; sCur[row0 + j] = 0xffff; }
            }
            block_sync<TPB>();
        }
        // ---- score and end plane (:390-391, :436-440) ----
        if (t == 0) { sEnd[0] = INF; sEnd[1] = INF; }
        block_sync<TPB>();
        if (!dead) {
            const int rq = X.Lq - 1, rr = X.padQ + X.Lr - 1;
            if (rq >= row0 && rq < row0 + K) {
#pragma unroll
                for (int j = 0; j < K; j++) if (row0 + j == rq) sEnd[0] = Dp[j];
            }
            if (rr >= row0 && rr < row0 + K) {
#pragma unroll
                for (int j = 0; j < K; j++) if (row0 + j == rr) sEnd[1] = Dp[j];
            }
        }
        block_sync<TPB>();
        score = min(sEnd[0], sEnd[1]);
        end_plane = sEnd[0] == score ? 0 : 1;
        block_sync<TPB>();
#ifdef VD_DEBUG
        if (t == 0) printf("fwd sc=%d ai=%d NP=%d Lt=%d tau=%d score=%d dead=%d\n", X.sc, X.ai, X.NP, X.Lt, tau, score, (int)dead);
#endif
        if (score <= tau || tau >= unbounded) break;
    }
}


// ------------------------------------------------------------------------------------------
// Banded forward sweep for long alignments.  Same recurrence and flags as wave_fwd_body, but only
// the rows that can hold a distance <= tau are visited in each column (Ukkonen's cut-off; exact
// whenever the final score is <= tau).  The band is a few hundred rows wide while the matrix side
// is thousands, so rows are dealt ONE per thread from the band's start in every column (D of the
// previous column lives in shared memory for all rows, so the thread<->row mapping is free to
// slide with the band); bands wider than the block are processed in chunks with the INS-chain
// carry handed from chunk to chunk.  tau = 80, 160, ... 1280 (with tau = 80 the candidate rows of both planes,
// 2 x (2 tau + 1 + tau), fit one 512-thread chunk); alignments that need more are flagged
// for the dense register-blocked kernel.
// ------------------------------------------------------------------------------------------
constexpr int FWDB_TPB = 512;
constexpr int FWDB_TAU0 = 80;
constexpr int FWDB_TAU_MAX = 1280;
__host__ __device__ inline int fwdb_smem(int npmax) { return 4 * npmax + 32 + 2 * 64 * 4 + 64; }

__global__ void __launch_bounds__(FWDB_TPB) wave_fwdb_kernel(WaveArgs A, int item0, int npmax, int *need_dense) {
    VD_DYN_SHARED(smem_raw);
    const WaveCtx X = wave_ctx(A, item0 + blockIdx.x);
    if (X.skip) return;
    u16 *sD0 = (u16 *)smem_raw, *sD1 = sD0 + npmax;
    int *sLive = (int *)(sD1 + npmax);
    int *sW = sLive + 8;                                     // [2 parities][2 segments][32 warps]
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    constexpr int NW = FWDB_TPB / 32;
    const int64_t oi = 4 * (int64_t)X.sc + X.ai;
    const int NPr = X.padQ + X.Lr;                           // rows in use (REF plane starts at padQ)
    bool solved = false;

    // Lower bound of the score: every path spells the reference with a SUBSET of the query haplotype's
    // variants applied (planes are only left / entered at variant boundaries, :335-337, :364-366), and
    // aligning a string of length Lm to the truth costs at least |Lm - Lt|.  Bounds below it are skipped -
    // an SV that only one side carries goes straight to the full-matrix kernel instead of failing three
    // ever wider banded sweeps first.
    __shared__ int s_lb;
    {
        const int qh = X.ai >> 1;
        const int64_t v0 = A.in.var_off[4 * (int64_t)X.sc + qh], v1 = A.in.var_off[4 * (int64_t)X.sc + qh + 1];
        const int nv = (int)(v1 - v0);
        if (t == 0) s_lb = nv <= 9 ? INF : 0;
        __syncthreads();
        if (nv <= 9 && t < (1 << nv)) {
            int lm = X.Lr;
            for (int j = 0; j < nv; j++)
                if ((t >> j) & 1) lm += (int)(A.in.alt_off[v0 + j + 1] - A.in.alt_off[v0 + j]) - A.in.var_rlen[v0 + j];
            atomicMin(&s_lb, abs(lm - X.Lt));
        }
        __syncthreads();
    }
    const int score_lb = s_lb;

    for (int tau = FWDB_TAU0; tau <= FWDB_TAU_MAX && !solved; tau *= 2) {
        if (tau < score_lb) continue;
        for (int r = t; r < NPr; r += FWDB_TPB) { sD0[r] = 0xffff; sD1[r] = 0xffff; }
        if (t == 0) {
            sLive[0] = sLive[2] = sLive[4] = sLive[6] = INF;
            sLive[1] = sLive[3] = sLive[5] = sLive[7] = -1;
        }
        __syncthreads();
        // candidate intervals of the two previous columns (plane-local, inclusive; empty: lo > hi)
        int p1lo[2] = {1, 1}, p1hi[2] = {0, 0}, p2lo[2] = {1, 1}, p2hi[2] = {0, 0};
        int tnext = X.tinfo[0];
        bool dead = false;
        int wpar = 0;
        for (int c = 0; c < X.Lt; c++) {
            u16 *sPrev = (c & 1) ? sD1 : sD0, *sCur = (c & 1) ? sD0 : sD1;
            const int tinfo = tnext;
            if (c + 1 < X.Lt) tnext = X.tinfo[c + 1];
            const int tch = tinfo & 0x7f;
            const bool tok = tinfo & 0x80;
            // ---- candidate rows of both planes ----
            int clo[2], chi[2];
            if (c == 0) { clo[0] = clo[1] = 0; chi[0] = min(tau, X.Lq - 1); chi[1] = min(tau, X.Lr - 1); }
            else {
                const int *lv = sLive + ((c - 1) & 1) * 4;
                const int loQ = lv[0], hiQ = lv[1], loR = lv[2], hiR = lv[3];
                if (hiQ < loQ && hiR < loR) { dead = true; break; }
                int lo0 = INF, hi0 = -1, lo1 = INF, hi1 = -1;
                if (hiQ >= loQ) { lo0 = loQ; hi0 = hiQ + 1; lo1 = X.qptr[loQ] + 1; hi1 = X.qptr[hiQ] + 1; }
                if (hiR >= loR) {
                    lo1 = min(lo1, loR); hi1 = max(hi1, hiR + 1);
                    lo0 = min(lo0, X.rptr[loR] + 1); hi0 = max(hi0, X.rptr[hiR] + 1);
                }
                clo[0] = max(lo0, 0); chi[0] = min(hi0 + tau, X.Lq - 1);
                clo[1] = max(lo1, 0); chi[1] = min(hi1 + tau, X.Lr - 1);
            }
            const int nQ = max(0, chi[0] - clo[0] + 1), nR = max(0, chi[1] - clo[1] + 1);
            const int total = nQ + nR;
            if (t == 0) *(short4 *)(X.band + 4 * (int64_t)c) = make_short4((short)(nQ ? clo[0] : 1), (short)(nQ ? chi[0] : 0),
                                                                             (short)(nR ? clo[1] : 1), (short)(nR ? chi[1] : 0));
            int cg0 = INF, cg1 = INF;                        // min(D - a) over the rows already done, per plane
            for (int base = 0; base < total; base += FWDB_TPB) {
                const int v = base + t;
                const bool valid = v < total;
                const bool P = v >= nQ;
                const int a = P ? clo[1] + (v - nQ) : clo[0] + v;
                const int row = P ? X.padQ + a : a;
                int b = INF, diag = INF, del = INF, swp = INF, sb = 0;
                bool m = false;
                if (valid) {
                    const int chv = (P ? X.rseq[a] : X.qstr[a]) & 0x7f;
                    m = chv == tch;
                    if (a == 0 && c == 0) b = 0;                                     // :299-305
                    else {
                        if (c > 0) {
                            const int dv = sPrev[row];
                            if (dv != 0xffff) del = dv + 1;                              // :406-413
                            if (a > 0) {
                                const int uv = sPrev[row - 1];
                                if (uv != 0xffff) diag = uv + (m ? 0 : 1);               // :324-332, :415-422
                            }
                            if (tok && m) {                                              // :334-349, :363-378
                                const u32 w = P ? X.swiR[a] : X.swiQ[a];
                                const int cnt = (int)(w >> 16);
                                if (cnt) {
                                    const int ob = P ? 0 : X.padQ;
                                    const int v0 = sPrev[ob + (int)(w & 0xffffu)];
                                    swp = v0 == 0xffff ? INF : v0;
                                    if (cnt > 1) {
                                        const int *tab = P ? X.toR : X.toQ;
                                        const int *src = tab + (P ? X.Lr : X.Lq) + 1;
                                        const int k0 = tab[a];
                                        for (int k = 1; k < cnt; k++) {
                                            int v2 = sPrev[ob + src[k0 + k]];
                                            v2 = v2 == 0xffff ? INF : v2;
                                            if (v2 < swp) { swp = v2; sb = k << F_K_SHIFT; }
                                            else if (v2 == swp) sb = (k << F_K_SHIFT) | F_TIE;
                                        }
                                    }
                                }
                            }
                        }
                        b = min(min(diag, del), swp);
                    }
                }
                // ---- min-plus scan of G = b - a over the threads of my plane in this chunk ----
                const int segstart = P ? max(nQ - base, 0) : 0;          // first thread of my plane in this chunk
                int incl = (valid && b < INF) ? b - a : INF;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int o = __shfl_up_sync(0xffffffffu, incl, d);
                    if (lane >= d && t - d >= segstart) incl = min(incl, o);
                }
                int excl = __shfl_up_sync(0xffffffffu, incl, 1);
                if (lane == 0 || t - 1 < segstart) excl = INF;
                int *sWc = sW + wpar * 64;
                wpar ^= 1;
                // per-warp totals of each plane: lanes of plane Q are a prefix of the chunk
                {
                    const unsigned mq = __ballot_sync(0xffffffffu, !P);
                    int tq = INF, tr = INF;
                    if (mq) tq = __shfl_sync(0xffffffffu, incl, 31 - __clz(mq));   // last Q lane: inclusive over Q lanes
                    if (~mq) tr = __shfl_sync(0xffffffffu, incl, 31);              // last lane: inclusive over R lanes
                    if (lane == 0) { sWc[warp] = mq ? tq : INF; sWc[32 + warp] = (~mq) ? tr : INF; }
                }
                __syncthreads();
                // everyone has read sLive[c-1] by now: recycle its slot for column c+1
                if (t == 0 && base == 0) { int *nx = sLive + ((c + 1) & 1) * 4; nx[0] = nx[2] = INF; nx[1] = nx[3] = -1; }
                int preq = (lane < warp) ? sWc[lane] : INF, prer = (lane < warp) ? sWc[32 + lane] : INF;
                int totq = (lane < NW) ? sWc[lane] : INF, totr = (lane < NW) ? sWc[32 + lane] : INF;
#pragma unroll
                for (int d = 16; d; d >>= 1) {
                    preq = min(preq, __shfl_xor_sync(0xffffffffu, preq, d));
                    prer = min(prer, __shfl_xor_sync(0xffffffffu, prer, d));
                    totq = min(totq, __shfl_xor_sync(0xffffffffu, totq, d));
                    totr = min(totr, __shfl_xor_sync(0xffffffffu, totr, d));
                }
                int liveA = -1;                              // my row if it is within tau
                if (valid) {
                    const int carry = min(excl, min(P ? prer : preq, P ? cg1 : cg0));
                    const int ins = (a > 0 && carry < INF / 2) ? carry + a : INF;         // = D[a-1][c] + 1  (:397-404)
                    const int d = min(b, ins);
                    int f = 0;
                    if (a == 0 && c == 0) f = F_DIAG;
                    else if (d < INF / 2) {
                        if (diag == d) f |= F_DIAG;
                        if (ins == d) f |= F_INS;
                        if (del == d) f |= F_DEL;
                        if (swp == d) f |= F_SWP | sb;
                    }
                    sCur[row] = d >= 0xffff ? (u16)0xffff : (u16)d;
                    X.F[(int64_t)c * X.NP + row] = (u8)f;
                    if (d <= tau) liveA = a;
                }
                {   // one shared-memory atomic per warp and bound instead of one per thread
                    const unsigned mq2 = __ballot_sync(0xffffffffu, liveA >= 0 && !P);
                    const unsigned mr2 = __ballot_sync(0xffffffffu, liveA >= 0 && P);
                    int *lc = sLive + (c & 1) * 4;
                    // rows ascend with the lane inside a plane: lowest / highest set lane give min / max
                    if (mq2) {
                        const int lo = __shfl_sync(0xffffffffu, liveA, __ffs(mq2) - 1);
                        const int hi = __shfl_sync(0xffffffffu, liveA, 31 - __clz(mq2));
                        if (lane == 0) { atomicMin(&lc[0], lo); atomicMax(&lc[1], hi); }
                    }
                    if (mr2) {
                        const int lo = __shfl_sync(0xffffffffu, liveA, __ffs(mr2) - 1);
                        const int hi = __shfl_sync(0xffffffffu, liveA, 31 - __clz(mr2));
                        if (lane == 0) { atomicMin(&lc[2], lo); atomicMax(&lc[3], hi); }
                    }
                }
                cg0 = min(cg0, totq); cg1 = min(cg1, totr);
            }
            // ---- rows that were candidates two columns ago but not now hold stale values ----
#pragma unroll
            for (int P = 0; P < 2; P++) {
                const int rb = P ? X.padQ : 0;
                for (int a = p2lo[P] + t; a <= p2hi[P]; a += FWDB_TPB)
                    if (a < clo[P] || a > chi[P]) sCur[rb + a] = 0xffff;
                p2lo[P] = p1lo[P]; p2hi[P] = p1hi[P];
                p1lo[P] = clo[P]; p1hi[P] = chi[P];
            }
            __syncthreads();
        }
        __syncthreads();
        if (!dead) {
            const u16 *sFin = (X.Lt & 1) ? sD1 : sD0;       // column Lt-1 was written as "cur" of its parity
            const int vq = sFin[X.Lq - 1], vr = sFin[X.padQ + X.Lr - 1];
            const int dq = vq == 0xffff ? INF : vq, dr = vr == 0xffff ? INF : vr;
            const int score = min(dq, dr);
            if (score <= tau) {
                solved = true;
                if (t == 0) {
                    A.out.aln_score[oi] = score;
                    A.out.aln_end_plane[oi] = (u8)(dq == score ? 0 : 1);
                }
            }
        }
        __syncthreads();
    }
    if (t == 0) need_dense[item0 + blockIdx.x] = solved ? 0 : 1;
}

template <int TPB, int K>
__global__ void __launch_bounds__(TPB) wave_fwd_kernel(WaveArgs A, int item0, const int *need_dense) {
    VD_DYN_SHARED(smem_raw);
    __shared__ int sEnd[2];
    if (need_dense && !need_dense[item0 + blockIdx.x]) return;       // solved by the banded sweep
    const WaveCtx X = wave_ctx(A, item0 + blockIdx.x);
    if (X.skip) return;
    int score, end_plane;
    // after a failed banded sweep (tau up to FWDB_TAU_MAX) go straight to the unbounded pass
    wave_fwd_body<TPB, K, int>(X, threadIdx.x, smem_raw, sEnd, score, end_plane, need_dense ? (1 << 28) : FWD_TAU0);
    if (threadIdx.x == 0) {
        A.out.aln_score[4 * (int64_t)X.sc + X.ai] = score;
        A.out.aln_end_plane[4 * (int64_t)X.sc + X.ai] = (u8)end_plane;
    }
}

// ------------------------------------------------------------------------------------------
// backward: calc_prec_recall_path (:486-834) as a reverse column sweep, path flags in place
// ------------------------------------------------------------------------------------------
// Returns the origin plane (valid on t == 0) and this thread's status bits.
template <int TPB, int K, class TT>
__device__ __forceinline__ void wave_bwd_body(const WaveCtxT<TT> &X, const int t, u8 *smem_raw, const int end_plane,
                                              int &beg_plane, u32 &status_out) {
    constexpr int NPMAX = TPB * K;
    short *sT0 = (short *)smem_raw, *sT1 = sT0 + NPMAX;      // T of column c+1 / c, all rows
    u8 *sF0 = (u8 *)(sT1 + NPMAX), *sF1 = sF0 + NPMAX;       // forward flags of column c+1 / c
    int *sW = (int *)(sF1 + NPMAX);                          // [32 warps][2] scan summaries
    const int lane = t & 31, warp = t >> 5;
    const int tQ = X.padQ / K;
    const bool P = t >= tQ;
    const int row0 = t * K;
    const int a0 = P ? row0 - X.padQ : row0;
    const int len = P ? X.Lr : X.Lq;
    const u8 *seq = P ? X.rseq : X.qstr;
    const int *sinfo = P ? X.srcR : X.srcQ;                  // my rows as swap SOURCES
    const int dbase = P ? 0 : X.padQ;                        // padded row offset of the destination plane
    const int erow = end_plane ? X.padQ + X.Lr - 1 : X.Lq - 1;

    u8 ch[K + 1];                                            // bases of rows a0 .. a0+K (one past, for the diagonal)
#pragma unroll
    for (int j = 0; j <= K; j++) ch[j] = (a0 + j < len) ? (u8)(seq[a0 + j] & 0x7f) : (u8)0xff;
    unsigned long long tpw = 0;                              // bit j: tp(a0+j), j = 0..K
    auto tpbit = [&](int j) -> int { return (int)((tpw >> j) & 1ull); };
#pragma unroll
    for (int j = 0; j <= K; j++) if (!P && a0 + j < len && X.tpb[a0 + j]) tpw |= 1ull << j;
    int si[K];
#pragma unroll
    for (int j = 0; j < K; j++) si[j] = (a0 + j < len) ? sinfo[a0 + j] : 0;

    int Tn[K];                                               // T of column c+1, my rows
    u32 Fn[(K + 3) / 4];                                     // forward flags of column c+1, my rows
#pragma unroll
    for (int j = 0; j < K; j++) Tn[j] = -1;
#pragma unroll
    for (int i = 0; i < (K + 3) / 4; i++) Fn[i] = 0;
    u32 status = 0;
    u32 Fc[(K + 3) / 4];
#pragma unroll
    for (int i = 0; i < (K + 3) / 4; i++) Fc[i] = 0;
    if (row0 < X.NP) load_flags<K>(X.F + (int64_t)(X.Lt - 1) * X.NP + row0, Fc);
    int tch_next = 0;                                        // truth base of column c+1

    for (int c = X.Lt - 1; c >= 0; c--) {
        short *sTn = (c & 1) ? sT1 : sT0, *sTc = (c & 1) ? sT0 : sT1;      // column c+1 was written as "cur" last iteration
        u8 *sFn = (c & 1) ? sF1 : sF0, *sFc = (c & 1) ? sF0 : sF1;
        const bool last = c == X.Lt - 1;
        // prefetch next column's forward flags
        u32 Fp[(K + 3) / 4];
#pragma unroll
        for (int i = 0; i < (K + 3) / 4; i++) Fp[i] = 0;
        if (c > 0 && row0 < X.NP) load_flags<K>(X.F + (int64_t)(c - 1) * X.NP + row0, Fp);
        // publish my first row's flag of column c for the neighbour's INS link
        sFc[row0] = (u8)byte_of(Fc, 0);
        // neighbour values of column c+1 (row a0+K): T and F
        int Tup = -1, Fup = 0;
        if (!last && a0 + K < len) { Tup = sTn[row0 + K]; Fup = sFn[row0 + K]; }
        // ---- pass 1: B[j] from column c+1, then local suffix chain ----
        int B[K];
#pragma unroll
        for (int j = K - 1; j >= 0; j--) {
            int b = -1;
            if (a0 + j < len) {
                if (last && row0 + j == erow) b = 0;                                   // :543-545
                if (!last) {
                    const int Tx = (j == K - 1) ? Tup : Tn[j + 1];                     // (P, a+1, c+1)
                    const int Fx = (j == K - 1) ? Fup : byte_of(Fn, j + 1);
                    if (a0 + j + 1 < len && Tx >= 0 && (Fx & F_DIAG)) b = max(b, Tx + tpbit(j + 1));
                    if (Tn[j] >= 0 && (byte_of(Fn, j) & F_DEL)) b = max(b, Tn[j]);     // (P, a, c+1)
                    const int s = si[j];
                    if (s & 1) {                                                       // swap: I am the recorded source
                        const int xr = dbase + (s >> 8);
                        const int Tx2 = sTn[xr], Fx2 = sFn[xr];
                        if (Tx2 >= 0 && (Fx2 & F_SWP) && (Fx2 >> F_K_SHIFT) == ((s >> 1) & 7)) {
                            b = max(b, Tx2 + ((s >> 4) & 1));
                            if (Fx2 & F_TIE) status |= VD_ST_TIE;
                        }
                    }
                }
            }
            B[j] = b;
        }
        // local chain assuming no carry-in: Tl[j] = max(B[j], link(j+1) ? Tl[j+1] + w(j+1) : -1)
        // cum[j] = weight of the unbroken INS chain from row a0+K down to row a0+j (NEG if broken)
        int Tl0, cumtop;
        {
            int run = -1, cum = 0;
            // link into row j comes from row j+1 of column c: flag F_INS of row j+1 (own Fc, or the
            // neighbour's first row, published in sFc after the sync below -> handled via cum)
#pragma unroll
            for (int j = K - 1; j >= 0; j--) {
                if (j < K - 1) {
                    const bool link = (a0 + j + 1 < len) && (byte_of(Fc, j + 1) & F_INS);
                    const int w = tpbit(j + 1);
                    run = (link && run >= 0) ? run + w : -1;
                    cum = (link && cum > NEG) ? cum + w : NEG;
                }
                run = max(run, B[j]);
            }
            Tl0 = run;
            cumtop = cum;       // chain weight from my row K-1 ... down to row 0 (excl. the link into K-1)
        }
        block_sync<TPB>();                                   // sFc[row0] of every thread visible
        // link from the neighbour's first row (a0+K) into my last row
        const bool toplink = (a0 + K < len) && (sFc[row0 + K] & F_INS);
        const int wtop = tpbit(K);
        // my affine-max function of the carry-in X = T[a0+K][c]:  T[a0] = max(Tl[0], X + Aw)
        int Aw = (toplink && cumtop > NEG) ? cumtop + wtop : NEG;
        int Mw = Tl0;
        // suffix composition across lanes: S_t = f_t o f_{t+1} o ... ; (A1,M1)o(A2,M2) = (A1+A2, max(M1, M2+A1))
        int As = Aw, Ms = Mw;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int A2 = __shfl_down_sync(0xffffffffu, As, d);
            const int M2 = __shfl_down_sync(0xffffffffu, Ms, d);
            if (lane + d < 32) {
                if (M2 >= 0 && As > NEG) Ms = max(Ms, M2 + As);
                As = (As > NEG && A2 > NEG) ? As + A2 : NEG;
            }
        }
        int Xw = -1;                                         // T of the first row of the next warp
        if constexpr (TPB > 32) {
            int *sWc = sW + (c & 1) * 64;
            if (lane == 0) { sWc[2 * warp] = As; sWc[2 * warp + 1] = Ms; }
            __syncthreads();
            constexpr int NW = TPB / 32;
            int Aq = (lane < NW) ? sWc[2 * lane] : NEG, Mq = (lane < NW) ? sWc[2 * lane + 1] : -1;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int A2 = __shfl_down_sync(0xffffffffu, Aq, d);
                const int M2 = __shfl_down_sync(0xffffffffu, Mq, d);
                if (lane + d < 32) {
                    if (M2 >= 0 && Aq > NEG) Mq = max(Mq, M2 + Aq);
                    Aq = (Aq > NEG && A2 > NEG) ? Aq + A2 : NEG;
                }
            }
            Xw = __shfl_sync(0xffffffffu, Mq, (warp + 1) & 31);
            if (warp + 1 >= NW) Xw = -1;
        }
        // T of my first row with the true carry, then the carry-in of each thread = T_first of t+1
        int Tfirst = Ms;
        if (Xw >= 0 && As > NEG) Tfirst = max(Tfirst, Xw + As);
        int Xin = __shfl_down_sync(0xffffffffu, Tfirst, 1);
        if (lane == 31) Xin = Xw;
        // ---- pass 2: final T and path flags ----
        u32 pfw[(K + 3) / 4];
#pragma unroll
        for (int i = 0; i < (K + 3) / 4; i++) pfw[i] = 0;
        int Tc[K];
        {
            int above = (toplink && Xin >= 0) ? Xin + wtop : -1;        // INS candidate into row K-1
#pragma unroll
            for (int j = K - 1; j >= 0; j--) {
                const int Tv = max(B[j], above);
                Tc[j] = Tv;
                int pf = 0;
                if (Tv >= 0 && a0 + j < len) {
                    if (last && row0 + j == erow && Tv == 0) pf |= PTR_MAT;            // :543
                    if (!last) {
                        const int Tx = (j == K - 1) ? Tup : Tn[j + 1];
                        const int Fx = (j == K - 1) ? Fup : byte_of(Fn, j + 1);
                        if (a0 + j + 1 < len && Tx >= 0 && (Fx & F_DIAG) && Tx + tpbit(j + 1) == Tv)
                            pf |= (ch[j + 1] == tch_next) ? PTR_MAT : PTR_SUB;
                        if (Tn[j] >= 0 && (byte_of(Fn, j) & F_DEL) && Tn[j] == Tv) pf |= PTR_DEL;
                        const int s = si[j];
                        if (s & 1) {
                            const int xr = dbase + (s >> 8);
                            const int Tx2 = sTn[xr], Fx2 = sFn[xr];
                            if (Tx2 >= 0 && (Fx2 & F_SWP) && (Fx2 >> F_K_SHIFT) == ((s >> 1) & 7) &&
                                Tx2 + ((s >> 4) & 1) == Tv) pf |= PTR_SWP;
                        }
                    }
                    if (above >= 0 && above == Tv) pf |= PTR_INS;
                }
                pfw[j >> 2] |= (u32)pf << ((j & 3) * 8);
                // INS candidate into row j-1 comes from row j
                if (j > 0) {
                    const bool link = (a0 + j < len) && (byte_of(Fc, j) & F_INS);
                    above = (link && Tv >= 0) ? Tv + tpbit(j) : -1;
                }
            }
        }
        // in-place: the path flags replace the forward flags of column c
#ifdef VD_DEBUG
        if (row0 < X.NP)
            printf("bwd c=%d t=%d row0=%d P=%d a0=%d len=%d Fc0=%x B0=%d Tl0=%d toplink=%d Aw=%d Ms=%d Tfirst=%d Xin=%d Tc0=%d pf0=%x Tup=%d Fup=%x erow=%d\n",
                   c, t, row0, (int)P, a0, len, byte_of(Fc, 0), B[0], Tl0, (int)toplink, Aw, Ms, Tfirst, Xin, Tc[0], byte_of(pfw, 0), Tup, Fup, erow);
#endif
        if (row0 < X.NP) store_flags<K>(X.F + (int64_t)c * X.NP + row0, pfw);
        // publish column c for the next iteration
#pragma unroll
        for (int j = 0; j < K; j++) {          // sFc[row0] was published above and is being read by the neighbour
            sTc[row0 + j] = (short)Tc[j]; Tn[j] = Tc[j];
            if (j > 0) sFc[row0 + j] = (u8)byte_of(Fc, j);
        }
#pragma unroll
        for (int i = 0; i < (K + 3) / 4; i++) { Fn[i] = Fc[i]; Fc[i] = Fp[i]; }
        tch_next = X.tinfo[c] & 0x7f;
        block_sync<TPB>();
    }
    // origin plane (:811-814): QUERY if its origin was reached
    beg_plane = Tn[0] >= 0 ? 0 : 1;
    status_out = status;
}

template <int TPB, int K>
__global__ void __launch_bounds__(TPB) wave_bwd_kernel(WaveArgs A, int item0, const int *only_dense) {
    VD_DYN_SHARED(smem_raw);
    if (only_dense && !only_dense[item0 + blockIdx.x]) return;       // solved by the banded sweeps
    const WaveCtx X = wave_ctx(A, item0 + blockIdx.x);
    if (X.skip) return;
    const int end_plane = A.out.aln_end_plane[4 * (int64_t)X.sc + X.ai];
    int beg_plane;
    u32 status;
    wave_bwd_body<TPB, K, int>(X, threadIdx.x, smem_raw, end_plane, beg_plane, status);
    if (threadIdx.x == 0) A.out.aln_beg_plane[4 * (int64_t)X.sc + X.ai] = (u8)beg_plane;
    if (status) atomicOr(&A.out.status[4 * (int64_t)X.sc + X.ai], status);
}

// ------------------------------------------------------------------------------------------
// sparse backward: the same pass as wave_bwd_kernel, restricted to the cells that are actually
// reachable from the end cell over optimal edges (what the reference's BFS visits, :549-808).
// One warp per alignment.  Per column it keeps the list of reached rows; pushes go from the
// reached cells of column c+1 into column c (atomic max on a dense int16 column in shared
// memory), the in-column INS chains are closed with a worklist, then the path flags of the
// reached cells overwrite their forward flags in HBM.  HBM traffic: ~3 bytes per REACHED cell
// instead of 2 bytes per cell of the whole matrix.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int atomic_max16(short *addr, int v) {
    unsigned *w = (unsigned *)((size_t)addr & ~(size_t)3);
    const int sh = ((size_t)addr & 2) ? 16 : 0;
    unsigned old = *w;
    while (true) {
        const int cur = (short)(old >> sh);
        if (v <= cur) return cur;
        const unsigned nw = (old & ~(0xffffu << sh)) | (((unsigned)v & 0xffffu) << sh);
        const unsigned got = atomicCAS(w, old, nw);
        if (got == old) return cur;
        old = got;
    }
}
__device__ __forceinline__ void atomic_or8(u8 *addr, int v) {
    unsigned *w = (unsigned *)((size_t)addr & ~(size_t)3);
    atomicOr(w, (unsigned)v << (((size_t)addr & 3) * 8));
}

struct SbwdPush { int n; int tgt[3]; int val[3]; int ty[3]; bool tie; };

// the (at most three) pushes of reached cell x = (row, column c+1) into column c
__device__ __forceinline__ SbwdPush sbwd_pushes(const WaveCtx &X, int row, int f, int Tx, int tch_x) {
    SbwdPush r; r.n = 0; r.tie = false;
    const bool P = row >= X.padQ;
    const int a = P ? row - X.padQ : row;
    const int tp = (!P && a > 0) ? X.tpb[a] : 0;
    if ((f & F_DIAG) && a > 0) {                                                    // :556-595, :692-731
        const int ch = (P ? X.rseq[a] : X.qstr[a]) & 0x7f;
        r.tgt[r.n] = row - 1; r.val[r.n] = Tx + tp; r.ty[r.n] = (ch == tch_x) ? PTR_MAT : PTR_SUB; r.n++;
    }
    if (f & F_DEL) { r.tgt[r.n] = row; r.val[r.n] = Tx; r.ty[r.n] = PTR_DEL; r.n++; }   // :774-804
    if ((f & F_SWP) && a > 0) {                                                     // :598-679
        const int of = P ? X.rflg[a] : X.qflg[a];
        if (!(of & P_VARIANT) || (of & P_VAR_BEG)) {
            const int *tab = P ? X.toR : X.toQ;
            const int *src = tab + (P ? X.Lr : X.Lq) + 1;
            r.tgt[r.n] = (P ? 0 : X.padQ) + src[tab[a] + (f >> F_K_SHIFT)];
            r.val[r.n] = Tx + (P ? 0 : tp); r.ty[r.n] = PTR_SWP; r.n++;
            r.tie = f & F_TIE;
        }
    }
    return r;
}

__global__ void __launch_bounds__(32) wave_sbwd_kernel(WaveArgs A, int item0, int npmax, const int *only_dense) {
    VD_DYN_SHARED(smem_raw);
    if (only_dense && !only_dense[item0 + blockIdx.x]) return;       // done by the banded backward sweep
    const WaveCtx X = wave_ctx(A, item0 + blockIdx.x);
    if (X.skip) return;
    short *T0 = (short *)smem_raw, *T1 = T0 + npmax;          // T of two columns, dense over rows
    u8 *PF0 = (u8 *)(T1 + npmax), *PF1 = PF0 + npmax;          // path flags of two columns
    __shared__ int cnt[2], wcnt[2];
    const int lane = threadIdx.x;
    // frontier lists and worklists live in the alignment's walk scratch (dead until the walk)
    int *L0 = (int *)X.walk, *L1 = L0 + X.NP, *W0 = L1 + X.NP, *W1 = W0 + X.NP;
    const int64_t oi = 4 * (int64_t)X.sc + X.ai;
    const int end_plane = A.out.aln_end_plane[oi];
    const int erow = end_plane ? X.padQ + X.Lr - 1 : X.Lq - 1;
    u32 status = 0;

    for (int r = lane; r < X.NP; r += 32) { T0[r] = -1; T1[r] = -1; PF0[r] = 0; PF1[r] = 0; }
    if (lane == 0) { cnt[0] = cnt[1] = 0; wcnt[0] = wcnt[1] = 0; }
    __syncwarp();

    short *Tn = T0, *Tc = T1;       // column c+1 ("next", already final) and column c (being built)
    u8 *PFn = PF0, *PFc = PF1;
    int *Ln = L0, *Lc = L1;
    int in = 0, ic = 1;             // indices into cnt[]
    int tch_next = 0;
    for (int c = X.Lt - 1; c >= 0; c--) {
        const u8 *Fcol = X.F + (int64_t)c * X.NP;
        // ---- phase A: seed / pushes from column c+1 ----
        if (c == X.Lt - 1) {
            if (lane == 0) { Tc[erow] = 0; PFc[erow] = PTR_MAT; Lc[0] = erow; cnt[ic] = 1; }      // :543-545
        } else {
            const int n = cnt[in];
            for (int i = lane; i < n; i += 32) {
                const int e = Ln[i], row = e & 0xffff, f = e >> 16;
                const SbwdPush pu = sbwd_pushes(X, row, f, Tn[row], tch_next);
                if (pu.tie) status |= VD_ST_TIE;
                for (int k = 0; k < pu.n; k++) {
                    const int old = atomic_max16(&Tc[pu.tgt[k]], pu.val[k]);
                    if (old < 0) Lc[atomicAdd(&cnt[ic], 1)] = pu.tgt[k];
                }
            }
        }
        __syncwarp();
        // forward flags of the rows reached so far; they also form the first worklist
        {
            const int n = cnt[ic];
            for (int i = lane; i < n; i += 32) {
                const int row = Lc[i] & 0xffff;
                const int e = row | ((int)Fcol[row] << 16);
                Lc[i] = e; W0[i] = e;
            }
            if (lane == 0) { wcnt[0] = n; wcnt[1] = 0; }
        }
        __syncwarp();
        // ---- in-column INS chains (:734-771): worklist until nothing improves ----
        {
            int *Wa = W0, *Wb = W1;
            int wa_i = 0;
            while (true) {
                const int n = wcnt[wa_i];
                if (n == 0) break;
                for (int i = lane; i < n; i += 32) {
                    const int e = Wa[i], row = e & 0xffff, f = e >> 16;
                    const bool P = row >= X.padQ;
                    const int a = P ? row - X.padQ : row;
                    if ((f & F_INS) && a > 0) {
                        const int v = Tc[row] + ((!P) ? X.tpb[a] : 0);
                        const int old = atomic_max16(&Tc[row - 1], v);
                        if (v > old) {
                            const int e2 = (row - 1) | ((int)Fcol[row - 1] << 16);
                            if (old < 0) Lc[atomicAdd(&cnt[ic], 1)] = e2;
                            Wb[atomicAdd(&wcnt[wa_i ^ 1], 1)] = e2;
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) wcnt[wa_i] = 0;
                int *t_ = Wa; Wa = Wb; Wb = t_;
                wa_i ^= 1;
                __syncwarp();
            }
        }
        // ---- phase B: path flags of column c ----
        if (c < X.Lt - 1) {
            const int n = cnt[in];
            for (int i = lane; i < n; i += 32) {
                const int e = Ln[i], row = e & 0xffff, f = e >> 16;
                const SbwdPush pu = sbwd_pushes(X, row, f, Tn[row], tch_next);
                for (int k = 0; k < pu.n; k++)
                    if (pu.val[k] == Tc[pu.tgt[k]]) atomic_or8(&PFc[pu.tgt[k]], pu.ty[k]);
            }
        }
        {
            const int n = cnt[ic];
            for (int i = lane; i < n; i += 32) {
                const int e = Lc[i], row = e & 0xffff, f = e >> 16;
                const bool P = row >= X.padQ;
                const int a = P ? row - X.padQ : row;
                if ((f & F_INS) && a > 0) {
                    const int v = Tc[row] + ((!P) ? X.tpb[a] : 0);
                    if (v == Tc[row - 1]) atomic_or8(&PFc[row - 1], PTR_INS);
                }
            }
        }
        __syncwarp();
        // ---- column c+1 is finished: its path flags replace the forward flags in HBM ----
        if (c < X.Lt - 1) {
            u8 *Fnext = X.F + (int64_t)(c + 1) * X.NP;
            const int n = cnt[in];
            for (int i = lane; i < n; i += 32) {
                const int row = Ln[i] & 0xffff;
                Fnext[row] = PFn[row];
                Tn[row] = -1; PFn[row] = 0;
            }
            __syncwarp();
            if (lane == 0) cnt[in] = 0;
        }
        tch_next = X.tinfo[c] & 0x7f;
        { short *t_ = Tn; Tn = Tc; Tc = t_; }
        { u8 *t_ = PFn; PFn = PFc; PFc = t_; }
        { int *t_ = Ln; Ln = Lc; Lc = t_; }
        in ^= 1; ic ^= 1;
        __syncwarp();
    }
    // column 0 (now "next"): origin plane (:811-814), then its path flags
    const int beg_plane = Tn[0] >= 0 ? 0 : 1;
    {
        u8 *F0 = X.F;
        const int n = cnt[in];
        for (int i = lane; i < n; i += 32) {
            const int row = Ln[i] & 0xffff;
            F0[row] = PFn[row];
        }
    }
    status = __reduce_or_sync(0xffffffffu, status);
    if (lane == 0) {
        A.out.aln_beg_plane[oi] = (u8)beg_plane;
        if (status) atomicOr(&A.out.status[oi], status);
    }
}


// ------------------------------------------------------------------------------------------
// Windowed backward sweep: calc_prec_recall_path (:486-834) in gather form over a window of rows
// that follows the reached cells.  Reached cells of column c come from reached cells x of column
// c+1 — (row-1) by a diagonal edge, (row) by a deletion edge, the recorded source row by a swap
// edge — and from the in-column insertion chain, which only runs towards lower rows.  So the rows
// to look at in column c are [lowest target - E, highest target] per plane, cut to the rows the
// banded forward sweep visited (X.band: cells outside hold no valid flags); if the chain is still
// alive at the lowest row of the window, one thread follows it further down (rare).  One row per
// thread; T and the forward flags of two columns live in shared memory (dense over rows, only
// entries of the processed windows are ever read).  The insertion chain
// T[r] = max(B[r], T[r+1] + tp(r+1)) is a link-segmented suffix max of T - S (S = X.tps), as in
// wsc_kernel.  Replaces the one-warp-per-alignment frontier kernel wherever the banded forward
// sweep succeeded: a column step is four block barriers on a few dozen rows instead of ~8
// dependent HBM/L2 round trips.
// ------------------------------------------------------------------------------------------
constexpr int BWDB_TPB = 256;
#ifndef VD_BWDB_E
#define VD_BWDB_E 24
#endif
constexpr int BWDB_E = VD_BWDB_E;             // rows below the lowest target that are looked at without the chain follower
__host__ __device__ inline int bwdb_smem(int npmax) { return 6 * npmax + 128 * 4; }

__global__ void __launch_bounds__(BWDB_TPB) wave_bwdb_kernel(WaveArgs A, int item0, int npmax, const int *need_dense) {
    VD_DYN_SHARED(smem_raw);
    if (need_dense[item0 + blockIdx.x]) return;              // full-matrix forward sweep: frontier kernel instead
    const WaveCtx X = wave_ctx(A, item0 + blockIdx.x);
    if (X.skip) return;
    short *sT0 = (short *)smem_raw, *sT1 = sT0 + npmax;      // T of column c+1 / c
    u8 *sF0 = (u8 *)(sT1 + npmax), *sF1 = sF0 + npmax;       // forward flags of column c+1 / c
    int *sW = (int *)(sF1 + npmax);                          // [NW] warp heads, [NW] all-linked flags, chunk carry
    constexpr int NW = BWDB_TPB / 32;
    int *sWin = sW + 2 * NW + 8;                             // [2 parities][lo0, hi0, lo1, hi1] targets of the next column
    int *sExt = sWin + 8;                                    // [2] new lowest processed row per plane after the chain follower
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int64_t oi = 4 * (int64_t)X.sc + X.ai;
    const int end_plane = A.out.aln_end_plane[oi];
    u32 status = 0;

    if (t == 0) {
        // targets of the last column: the end cell
        int *w = sWin + ((X.Lt - 1) & 1) * 4;
        w[0] = w[2] = INF; w[1] = w[3] = -1;
        const int ea = end_plane ? X.Lr - 1 : X.Lq - 1;
        w[2 * end_plane] = ea; w[2 * end_plane + 1] = ea;
    }
    __syncthreads();

    int nlo[2] = {1, 1}, nhi[2] = {0, 0};                    // processed rows of column c+1 (plane-local, inclusive)
    int tch_next = 0;
    for (int c = X.Lt - 1; c >= 0; c--) {
        short *sTn = (c & 1) ? sT1 : sT0, *sTc = (c & 1) ? sT0 : sT1;
        u8 *sFn = (c & 1) ? sF1 : sF0, *sFc = (c & 1) ? sF0 : sF1;
        const bool last = c == X.Lt - 1;
        const short4 bd = *(const short4 *)(X.band + 4 * (int64_t)c);
        const int blo[2] = {bd.x, bd.z}, bhi[2] = {bd.y, bd.w};                        // rows with valid forward flags
        int *wcur = sWin + (c & 1) * 4, *wnext = sWin + ((c + 1) & 1) * 4;
        int clo[2], chi[2];
#pragma unroll
        for (int P = 0; P < 2; P++) {
            clo[P] = max(max(wcur[2 * P] - BWDB_E, 0), blo[P]);
            chi[P] = min(wcur[2 * P + 1], bhi[P]);
        }
        const int nQ = max(0, chi[0] - clo[0] + 1), nR = max(0, chi[1] - clo[1] + 1);
        const int total = nQ + nR;
        const int erow = end_plane ? X.padQ + X.Lr - 1 : X.Lq - 1;
        u8 *Fcol = X.F + (int64_t)c * X.NP;
        auto Tnext = [&](int P, int a) -> int { return (a >= nlo[P] && a <= nhi[P]) ? (int)sTn[(P ? X.padQ : 0) + a] : -1; };
        auto Fnext = [&](int P, int a) -> int { return (a >= nlo[P] && a <= nhi[P]) ? (int)sFn[(P ? X.padQ : 0) + a] : 0; };
        __syncthreads();                                     // everyone has read wcur
        if (t == 0) { wnext[0] = wnext[2] = INF; wnext[1] = wnext[3] = -1; }            // becomes the target set of column c-1
        // my contribution to the targets of column c-1: reached cell (P, a) with forward flags f
        auto contribute = [&](int P, int a, int f) {
            if (a > 0 && (f & F_DIAG)) atomicMin(&wnext[2 * P], a - 1), atomicMax(&wnext[2 * P + 1], a - 1);
            if (f & F_DEL) atomicMin(&wnext[2 * P], a), atomicMax(&wnext[2 * P + 1], a);
            if ((f & F_SWP) && a > 0) {                                                // :598-679
                const int of = P ? X.rflg[a] : X.qflg[a];
                if (!(of & P_VARIANT) || (of & P_VAR_BEG)) {
                    const int *tab = P ? X.toR : X.toQ;
                    const int *src = tab + (P ? X.Lr : X.Lq) + 1;
                    const int z = src[tab[a] + (f >> F_K_SHIFT)];
                    atomicMin(&wnext[2 * (1 - P)], z), atomicMax(&wnext[2 * (1 - P) + 1], z);
                }
            }
        };
        int chunk_carry = NEG;                               // U of the lowest row of the chunk above (processed before)
        bool chunk_first = true;
        for (int base = ((total - 1) / BWDB_TPB) * BWDB_TPB; base >= 0; base -= BWDB_TPB) {
            const int v = base + t;
            const bool valid = v < total;
            const bool P = v >= nQ;
            const int a = P ? clo[1] + (v - nQ) : clo[0] + v;
            const int len = P ? X.Lr : X.Lq;
            const int row = P ? X.padQ + a : a;
            // ---- stage A: forward flags of my cell, candidates from column c+1 ----
            int fc = 0, B = -1, swv = -1, tpn = 0, S = 0, Td = -1, Fd = 0, Tn = -1, Fn = 0;
            bool below_ok = false;
            if (valid) {
                fc = Fcol[row];
                sFc[row] = (u8)fc;
                below_ok = a + 1 < len;
                if (!P) { S = X.tps[a]; if (below_ok) tpn = X.tpb[a + 1]; }
                if (last && row == erow) B = 0;                                        // :543-545
                if (!last) {
                    if (below_ok) { Td = Tnext(P, a + 1); Fd = Fnext(P, a + 1); }
                    Tn = Tnext(P, a); Fn = Fnext(P, a);
                    if (Td >= 0 && (Fd & F_DIAG)) B = max(B, Td + tpn);                // :556-595, :692-731
                    if (Tn >= 0 && (Fn & F_DEL)) B = max(B, Tn);                       // :774-804
                    const int si = P ? X.srcR[a] : X.srcQ[a];                          // my row as the recorded swap source
                    if (si & 1) {                                                      // :598-679
                        const int d = si >> 8;
                        const int T2 = Tnext(P ? 0 : 1, d), F2 = Fnext(P ? 0 : 1, d);
                        if (T2 >= 0 && (F2 & F_SWP) && (F2 >> F_K_SHIFT) == ((si >> 1) & 7)) {
                            swv = T2 + ((si >> 4) & 1);
                            B = max(B, swv);
                            if (F2 & F_TIE) status |= VD_ST_TIE;
                        }
                    }
                }
            }
            __syncthreads();                                                           // sFc of this chunk (and the one above) visible
            // ---- stage B: insertion chain (:734-771): link-segmented suffix max of T - S ----
            bool link = false;                               // (row+1, c) is in my plane, was processed and has F_INS
            if (valid && below_ok && a + 1 <= chi[P]) link = sFc[row + 1] & F_INS;
            const unsigned lm = __ballot_sync(0xffffffffu, link);
            const unsigned nm = ~(lm >> lane);
            const int run = nm ? __ffs(nm) - 1 : 32;
            int val = (valid && B >= 0) ? B - S : NEG;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int o = __shfl_down_sync(0xffffffffu, val, d);
                if (run >= d && lane + d < 32) val = max(val, o);
            }
            if (lane == 0) { sW[warp] = val; sW[NW + warp] = (lm == 0xffffffffu) ? 1 : 0; }
            __syncthreads();
            {   // carry into my warp: what arrives at its lane 31 from above
                int acc = chunk_first ? NEG : chunk_carry;
                for (int w = NW - 1; w > warp; w--) acc = max(sW[w], sW[NW + w] ? acc : NEG);
                if (run == 32 - lane) val = max(val, acc);
            }
            const int T = (valid && val > NEG / 2) ? val + S : -1;
            if (valid) sTc[row] = (short)T;
            const int head_all = __shfl_sync(0xffffffffu, val, 0);
            __syncthreads();                                                           // sTc visible; sW free again
            if (t == 0) sW[2 * NW] = head_all;                                         // U of the chunk's lowest row
            // ---- stage C: path flags (in place), targets of the next column ----
            if (valid) {
                int pf = 0;
                if (T >= 0) {
                    if (last && row == erow && T == 0) pf |= PTR_MAT;                  // :543
                    if (!last) {
                        if (Td >= 0 && (Fd & F_DIAG) && Td + tpn == T) {
                            const int chn = (P ? X.rseq[a + 1] : X.qstr[a + 1]) & 0x7f;
                            pf |= (chn == tch_next) ? PTR_MAT : PTR_SUB;
                        }
                        if (Tn >= 0 && (Fn & F_DEL) && Tn == T) pf |= PTR_DEL;
                        if (swv >= 0 && swv == T) pf |= PTR_SWP;
                    }
                    if (link) {
                        const int tb = sTc[row + 1];
                        if (tb >= 0 && tb + tpn == T) pf |= PTR_INS;
                    }
                    contribute(P, a, fc);
                }
                Fcol[row] = (u8)pf;
            }
            __syncthreads();
            chunk_carry = sW[2 * NW];
            chunk_first = false;
        }
        // ---- chain follower: the insertion chain is still alive at the lowest row of a window ----
        int plo[2] = {clo[0], clo[1]};
        {
            bool need[2];
#pragma unroll
            for (int P = 0; P < 2; P++) {
                const int r0 = (P ? X.padQ : 0) + clo[P];
                need[P] = (P ? nR : nQ) > 0 && clo[P] > max(blo[P], 0) && sTc[r0] >= 0 && (sFc[r0] & F_INS);
            }
            if (need[0] || need[1]) {                        // block-uniform
                if (t == 0) {
                    for (int P = 0; P < 2; P++) {
                        int a = clo[P];
                        if (need[P]) {
                            const int rb = P ? X.padQ : 0;
                            // cell (a-1) is reached from (a) when F[a] has F_INS (:734-771); nothing else targets it
                            while (a > max(blo[P], 0) && sTc[rb + a] >= 0 && (sFc[rb + a] & F_INS)) {
                                const int tv = sTc[rb + a] + ((!P) ? (int)X.tpb[a] : 0);
                                a--;
                                const int f = Fcol[rb + a];
                                sFc[rb + a] = (u8)f;
                                sTc[rb + a] = (short)tv;
                                Fcol[rb + a] = (u8)PTR_INS;
                                contribute(P, a, f);
                            }
                        }
                        sExt[P] = a;
                    }
                }
                __syncthreads();
                plo[0] = sExt[0]; plo[1] = sExt[1];
            }
        }
        nlo[0] = plo[0]; nhi[0] = chi[0]; nlo[1] = plo[1]; nhi[1] = chi[1];
        if (nQ == 0) { nlo[0] = 1; nhi[0] = 0; }
        if (nR == 0) { nlo[1] = 1; nhi[1] = 0; }
        tch_next = X.tinfo[c] & 0x7f;
        __syncthreads();                                     // wnext complete before the next column reads it
    }
    // origin plane (:811-814): QUERY if its origin was reached
    status = __reduce_or_sync(0xffffffffu, status);
    if (lane == 0 && status) atomicOr(&A.out.status[oi], status);
    if (t == 0) {
        const short *sTl = sT1;                              // column 0 was written as "cur" of c = 0
        const int t00 = (0 >= nlo[0] && 0 <= nhi[0]) ? (int)sTl[0] : -1;
        A.out.aln_beg_plane[oi] = (u8)(t00 >= 0 ? 0 : 1);
    }
}

// path flags of the wavefront layout
struct PFWave {
    const u8 *F; int NP, padQ; int Lt;
    __device__ __forceinline__ int get(int hi, int qri, int ti) const {
        return F[(int64_t)ti * NP + (hi ? padQ + qri : qri)];
    }
    // The walk is one dependent flag load per step, each in a new 128-byte line (a diagonal step moves one
    // column = NP bytes): pull the lines of the cells AHEAD steps further along the diagonal and along the
    // row (deletion runs) towards L1 while the current step is still being decided.
    static constexpr int AHEAD = 8;
    __device__ __forceinline__ void prefetch(int hi, int qri, int ti) const {
        if (ti + AHEAD < Lt) {
            const u8 *p = F + (int64_t)(ti + AHEAD) * NP + (hi ? padQ + qri : qri);
#ifndef VD_EMU
            asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(p + AHEAD));
#else
            (void)p;
#endif
        }
    }
};

// one thread per alignment: walk + credit
// One alignment per WARP (lane 0 walks): walks of very different lengths and move mixes in one warp
// serialise each other's branches, and there are only a few thousand long alignments in a batch.
__global__ void wave_walk_kernel(WaveArgs A, int item0, int n_items, int wpw) {
    const int lane = threadIdx.x & 31, lpi = 32 / wpw;                 // wpw alignments per warp (see band_walk_kernel)
    const int g = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * wpw + lane / lpi;
    if (g >= n_items || (lane % lpi)) return;
    if (A.bstate && A.bstate[item0 + g] > 0) return;         // walked by band_walk_kernel
    const int item = A.items[item0 + g];
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
    PFWave pfr{A.dense + A.dense_off[item0 + g], wa.NP, wa.padQ, t.len};
    u32 status = A.out.status[4 * (int64_t)sc + ai];
    const int beg_plane = A.out.aln_beg_plane[4 * (int64_t)sc + ai];
    const int end_plane = A.out.aln_end_plane[4 * (int64_t)sc + ai];
    walk_credit<GMem, 4, int>(mem, L, pfr, q, qm, t, rseq, p.lr, beg_plane, end_plane, A.in, A.out, sc, ai, status);
    A.out.status[4 * (int64_t)sc + ai] = status;
}

// Alignments no block kernel takes (more than 32768 rows over both planes, or a score that could overflow the
// 16-bit columns) and the banded kernels did not solve either: one thread per alignment with every matrix in
// the dense-phase scratch (the reference only WARNs about the RAM and computes such a supercluster,
// src/cluster.cpp:102-107).  Slow, but it never refuses.
__global__ void wave_oversize_kernel(WaveArgs A, int item0, int n_items) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_items) return;
    if (A.bstate && A.bstate[item0 + g] > 0) return;
    const int item = A.items[item0 + g];
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
    GMem mem{A.dense + A.dense_off[item0 + g]};
    const AlnLayout<int64_t> L = make_layout<int64_t, 4, false>(q.len + p.lr, t.len, p.lr);
    u32 status = 0;
    int score, end_plane;
    forward_scalar<GMem, 4, int>(mem, L, q, qm, t, rseq, p.lr, score, end_plane);
    const int beg_plane = backward_scalar<GMem, 4, int>(mem, L, q, qm, t, rseq, p.lr, end_plane, status);
    PFScalar<GMem> pfr{&mem, L.oPF, q.len + p.lr, q.len};
    walk_credit<GMem, 4, int>(mem, L, pfr, q, qm, t, rseq, p.lr, beg_plane, end_plane, A.in, A.out, sc, ai, status);
    A.out.aln_score[4 * (int64_t)sc + ai] = score;
    A.out.aln_end_plane[4 * (int64_t)sc + ai] = (u8)end_plane;
    A.out.aln_beg_plane[4 * (int64_t)sc + ai] = (u8)beg_plane;
    A.out.status[4 * (int64_t)sc + ai] = status;
}

// homozygous long superclusters: alignment 0 was computed, its records go to the other haplotype and slot
__global__ void wave_hom_replicate_kernel(BatchDev in, OutDev out, const ScPlan *plan, const int *list, int i0, int i1, const int *hap_ok) {
    const int i = i0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= i1) return;
    const int sc = list[i];
    const ScPlan p = plan[sc];
    if (p.cls != CLS_WAVE || !p.hom) return;
    const int *okp = hap_ok + 4 * (int64_t)(i - i0);
    if (!(okp[0] && okp[1] && okp[2] && okp[3])) return;
    const int64_t oi = 4 * (int64_t)sc;
    for (int k = 1; k < 4; k++) {
        out.aln_score[oi + k] = out.aln_score[oi];
        out.aln_end_plane[oi + k] = out.aln_end_plane[oi];
        out.aln_beg_plane[oi + k] = out.aln_beg_plane[oi];
        out.status[oi + k] = out.status[oi];
    }
    replicate_hom(in, out, sc);
}

// ---- host side --------------------------------------------------------------------------------
template <int TPB, int K> constexpr int wave_fwd_smem() { return 2 * TPB * K * 2 + 32 + 2 * 64 * 4; }
template <int TPB, int K> constexpr int wave_bwd_smem() { return 2 * TPB * K * 2 + 2 * TPB * K + 2 * 64 * 4; }

template <int TPB, int K> inline void wave_configure_one() {
    cudaFuncSetAttribute(wave_fwd_kernel<TPB, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, wave_fwd_smem<TPB, K>());
    cudaFuncSetAttribute(wave_bwd_kernel<TPB, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, wave_bwd_smem<TPB, K>());
}
inline void wave_configure() {
    wave_configure_one<32, 1>(); wave_configure_one<32, 2>(); wave_configure_one<32, 4>();
    wave_configure_one<128, 4>(); wave_configure_one<256, 8>(); wave_configure_one<512, 16>();
    wave_configure_one<1024, 16>(); wave_configure_one<1024, 32>();
    cudaFuncSetAttribute(wave_sbwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * 32768);
    cudaFuncSetAttribute(wave_fwdb_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, fwdb_smem(32768));
    cudaFuncSetAttribute(wave_bwdb_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bwdb_smem(32768));
}

constexpr int FWDB_MIN_CLASS = 4;      // classes with more than 512 rows go through the banded sweep first
template <int TPB, int K>
inline void wave_launch_pair(cudaStream_t st, const WaveArgs &A, int item0, int n, bool fwd, int *need_dense = nullptr) {
    if (n <= 0) return;
    auto kf = wave_fwd_kernel<TPB, K>;
    auto kb = wave_bwd_kernel<TPB, K>;
    const int smf = wave_fwd_smem<TPB, K>(), smb = wave_bwd_smem<TPB, K>();
    if (fwd && need_dense) {
        VD_LAUNCH(wave_fwdb_kernel, n, FWDB_TPB, fwdb_smem(TPB * K), st, A, item0, TPB * K, need_dense);
        VD_LAUNCH(kf, n, TPB, smf, st, A, item0, (const int *)need_dense);
    } else if (fwd) VD_LAUNCH(kf, n, TPB, smf, st, A, item0, (const int *)nullptr);
    else VD_LAUNCH(kb, n, TPB, smb, st, A, item0, (const int *)need_dense);
}
// backward of a class that went through the banded forward sweep: the windowed sweep for the alignments
// it solved; the DENSE sweep for the rest (score above the last bound: a structural variant that only one
// side carries - nearly every cell of such a matrix lies on some optimal path, so the frontier kernel
// would crawl over the whole matrix with one warp).  banded_bwd = false: frontier kernel for all of them
// (VD_SPARSE_BWD=1, testing).  Classes without the banded forward sweep: dense or frontier kernel.
inline void wave_launch(cudaStream_t st, const WaveArgs &A, int cls, int item0, int n, bool fwd, bool sparse_bwd = true,
                        int *need_dense = nullptr, bool banded_bwd = true) {
    if (cls < FWDB_MIN_CLASS) need_dense = nullptr;
    const int npmax = wave_tpb(cls) * wave_k(cls);
    if (!fwd && sparse_bwd && need_dense && banded_bwd)
        VD_LAUNCH(wave_bwdb_kernel, n, BWDB_TPB, bwdb_smem(npmax), st, A, item0, npmax, (const int *)need_dense);   // then the dense kernel below
    else if (!fwd && sparse_bwd) {
        VD_LAUNCH(wave_sbwd_kernel, n, 32, 6 * npmax, st, A, item0, npmax, (const int *)nullptr);
        return;
    }
    if (!fwd && !(sparse_bwd && banded_bwd)) need_dense = nullptr;                               // dense sweep for everything
    switch (cls) {
        case 0: wave_launch_pair<32, 1>(st, A, item0, n, fwd, need_dense); break;
        case 1: wave_launch_pair<32, 2>(st, A, item0, n, fwd, need_dense); break;
        case 2: wave_launch_pair<32, 4>(st, A, item0, n, fwd, need_dense); break;
        case 3: wave_launch_pair<128, 4>(st, A, item0, n, fwd, need_dense); break;
        case 4: wave_launch_pair<256, 8>(st, A, item0, n, fwd, need_dense); break;
        case 5: wave_launch_pair<512, 16>(st, A, item0, n, fwd, need_dense); break;
        case 6: wave_launch_pair<1024, 16>(st, A, item0, n, fwd, need_dense); break;
        case 7: wave_launch_pair<1024, 32>(st, A, item0, n, fwd, need_dense); break;
    }
}

}  // namespace vd
