// One-thread-per-alignment pipeline: haplotype expansion (generate_ptrs_strs), forward
// two-plane DP, backward max-TP pass, path walk, sync sections and integer credit.
//
// The same code runs in two places, selected by the memory accessor:
//   * the fused short-supercluster kernels (small_kernel / small_hom_kernel, vd_kernels.cuh): every matrix of the alignment lives
//     in shared memory, one slice per thread with an odd word stride (bank-conflict free
//     when lanes touch the same offset);
//   * the scalar fallback kernel (slab_align_kernel, vd_kernels.cuh): matrices in an HBM scratch slab, for shapes
//     the wavefront kernels do not take (e.g. more than two swap sources per row).
// The walk / credit / Levenshtein tail is also what the long-supercluster path runs after
// its wavefront forward and backward kernels.
//
// Reference lines cited are in /root/reference/src/dist.cpp unless stated.
#pragma once
#include "vd_common.cuh"

namespace vd {

// ---------------------------------------------------------------------------------------
// memory accessors for thread-private scratch
// ---------------------------------------------------------------------------------------
struct GMem {                      // plain HBM slab
    typedef int64_t off_t;
    u8 *p;
    __device__ __forceinline__ int ld8(off_t o) const { return p[o]; }
    __device__ __forceinline__ void st8(off_t o, int v) const { p[o] = (u8)v; }
    __device__ __forceinline__ int ld32(off_t o) const { return *(const int *)(p + o); }
    __device__ __forceinline__ void st32(off_t o, int v) const { *(int *)(p + o) = v; }
};

struct SMemIL {                    // shared memory: one contiguous slice per thread whose stride is an
    typedef int off_t;             // ODD number of 32-bit words, so that lanes touching the same offset
    u8 *base;                      // hit 32 different banks (no interleaving arithmetic per access)
    __device__ __forceinline__ u8 *at(int o) const { return base + o; }
    __device__ __forceinline__ int ld8(int o) const { return base[o]; }
    __device__ __forceinline__ void st8(int o, int v) const { base[o] = (u8)v; }
    __device__ __forceinline__ int ld16(int o) const { return *(const short *)(base + o); }
    __device__ __forceinline__ void st16(int o, int v) const { *(short *)(base + o) = (short)v; }
};

struct GMemIL {                    // the same byte/halfword interface over a region of HBM (path scratch of wsc_walk_kernel)
    typedef int off_t;
    u8 *base;
    __device__ __forceinline__ u8 *at(int o) const { return base + o; }
    __device__ __forceinline__ int ld8(int o) const { return base[o]; }
    __device__ __forceinline__ void st8(int o, int v) const { base[o] = (u8)v; }
    __device__ __forceinline__ int ld16(int o) const { return *(const short *)(base + o); }
    __device__ __forceinline__ void st16(int o, int v) const { *(short *)(base + o) = (short)v; }
};

// value accessors: W = element width of the D / T / path / lev arrays (4 in HBM, 2 in smem)
template <class Mem, int W> struct Val;
template <> struct Val<GMem, 4> {
    static __device__ __forceinline__ int ld(const GMem &m, int64_t base, int i) { return m.ld32(base + 4 * (int64_t)i); }
    static __device__ __forceinline__ void st(const GMem &m, int64_t base, int i, int v) { m.st32(base + 4 * (int64_t)i, v); }
};
template <> struct Val<SMemIL, 2> {
    static __device__ __forceinline__ int ld(const SMemIL &m, int base, int i) { return m.ld16(base + 2 * i); }
    static __device__ __forceinline__ void st(const SMemIL &m, int base, int i, int v) { m.st16(base + 2 * i, v); }
};
template <> struct Val<GMemIL, 2> {
    static __device__ __forceinline__ int ld(const GMemIL &m, int base, int i) { return m.ld16(base + 2 * i); }
    static __device__ __forceinline__ void st(const GMemIL &m, int base, int i, int v) { m.st16(base + 2 * i, v); }
};

// expanded haplotype (generate_ptrs_strs output), PT = pointer element type
template <class PT> struct Hap {
    int len;
    const u8 *str;     // bases
    const u8 *flg;     // hap -> ref flags
    const PT *ptr;     // hap -> ref pointers
    const u8 *ins;     // [Lr] 1 where this hap has an insertion after that ref base (:886-894)
};
template <class PT> struct QMaps {     // per query hap
    const PT *rptr;    // ref -> query pointers
    const u8 *rflg;    // ref -> query flags
    const PT *toQ;     // CSR [Lq+1 | <=Lr]: swap sources (REF rows) of QUERY-plane destination rows
    const PT *toR;     // CSR [Lr+1 | <=Lq]: swap sources (QUERY rows) of REF-plane destination rows
};

// byte offsets of one alignment's private scratch
template <class OffT> struct AlnLayout {
    OffT oPF, oF, oD0, oD1, oT0, oT1, oPQ, oPT, oPS, oLev, total;
};

// W = width of value arrays.  ALIAS: the path and Levenshtein row reuse the F / column
// region, which is dead once the backward pass is done (smem is scarce, HBM is not).
template <class OffT, int W, bool ALIAS>
__host__ __device__ inline AlnLayout<OffT> make_layout(int N, int Lt, int Lr) {
    AlnLayout<OffT> L;
    OffT cells = (OffT)N * Lt;
    OffT a4 = (cells + 3) & ~(OffT)3;
    L.oPF = 0;
    L.oF = a4;
    L.oD0 = L.oF + a4;
    L.oD1 = L.oD0 + (OffT)((N * W + 3) & ~3);
    L.oT0 = L.oD1 + (OffT)((N * W + 3) & ~3);
    L.oT1 = L.oT0 + (OffT)((N * W + 3) & ~3);
    OffT end_cols = L.oT1 + (OffT)((N * W + 3) & ~3);
    int np = N + Lt + 4;
    OffT pbase = ALIAS ? L.oF : end_cols;
    L.oPQ = pbase;
    L.oPT = L.oPQ + (OffT)((np * W + 3) & ~3);
    L.oPS = L.oPT + (OffT)((np * W + 3) & ~3);
    L.oLev = L.oPS + (OffT)((np + 3) & ~3);
    int mn = (Lr < Lt ? Lr : Lt) + 1;
    OffT end_path = L.oLev + (OffT)((mn * W + 3) & ~3);
    L.total = ALIAS ? (end_cols > end_path ? end_cols : end_path) : end_path;
    return L;
}

// ---------------------------------------------------------------------------------------
// generate_ptrs_strs, :145-242.  Writes str/flg/ptr (hap side) and, when rptr != nullptr,
// the ref side.  Returns the hap length, or -1 on input the reference cannot process.
// ---------------------------------------------------------------------------------------
// Where the expansion reads a supercluster's compact input from: straight from the batch in HBM, or from a copy the
// warp staged in shared memory with coalesced loads (wsc_block_kernel) - the expansion is one lane following offsets,
// i.e. a chain of dependent loads, and each link costs an HBM/L2 round trip or a shared-memory one.
struct HapSrcGlobal {
    const BatchDev &in; int sc, h;
    int64_t r0, v0, ve;
    __device__ HapSrcGlobal(const BatchDev &in_, int sc_, int h_) : in(in_), sc(sc_), h(h_) {
        r0 = in.ref_off[sc];
        v0 = in.var_off[4 * (int64_t)sc + h];
        ve = in.var_off[4 * (int64_t)sc + h + 1];
    }
    __device__ int win() const { return (int)(in.ref_off[sc + 1] - r0); }
    __device__ int nvar() const { return (int)(ve - v0); }
    __device__ int fa(int k) const { return in.ref_seq[r0 + k]; }
    __device__ int pos(int v) const { return in.var_pos[v0 + v]; }
    __device__ int rlen(int v) const { return in.var_rlen[v0 + v]; }
    __device__ int type(int v) const { return in.var_type[v0 + v]; }
    __device__ const u8 *alt(int v, int &alen) const {
        const int64_t a0 = in.alt_off[v0 + v];
        alen = (int)(in.alt_off[v0 + v + 1] - a0);
        return in.alt_seq + a0;
    }
};
struct HapSrcStaged {        // variant v of the haplotype is entry vb + v of the supercluster's staged arrays
    const u8 *fa_; int win_; int vb, n;
    const int *pos_, *rlen_, *aoff_;        // aoff_: low words of alt_off, nv + 1 of them
    const u8 *type_, *alt_;                 // alt_: ALT bytes from the supercluster's first variant on
    __device__ int win() const { return win_; }
    __device__ int nvar() const { return n; }
    __device__ int fa(int k) const { return fa_[k]; }
    __device__ int pos(int v) const { return pos_[vb + v]; }
    __device__ int rlen(int v) const { return rlen_[vb + v]; }
    __device__ int type(int v) const { return type_[vb + v]; }
    __device__ const u8 *alt(int v, int &alen) const {
        alen = aoff_[vb + v + 1] - aoff_[vb + v];
        return alt_ + (aoff_[vb + v] - aoff_[0]);
    }
};

template <class PT, class SRC>
__device__ int expand_hap_src(const SRC &src, u8 *str, u8 *flg, PT *ptr, PT *rptr, u8 *rflg, u8 *ins, int cap) {
    const int win = src.win();
    int v = 0;
    const int ve = src.nvar();
    int Q = 0, R = 0, ref_pos = 0;
    if (ins) for (int k = 0; k < win; k++) ins[k] = 0;
    while (ref_pos < win) {                                      // :163
        if (v < ve && ref_pos == src.pos(v)) {                   // :165-166
            int alen;
            const u8 *alt = src.alt(v, alen);
            const int rl = src.rlen(v);
            const int ty = src.type(v);
            if (ty == VD_TYPE_INS) {                             // :169-178
                if (alen < 1 || Q + alen > cap) return -1;
                for (int k = 0; k < alen; k++) {
                    ptr[Q + k] = (PT)(R - 1);
                    flg[Q + k] = P_VARIANT;
                    str[Q + k] = alt[k];
                }
                flg[Q + alen - 1] |= P_VAR_END;
                flg[Q] |= P_VAR_BEG | P_INS_LOC;
                if (ins && R - 1 >= 0) ins[R - 1] = 1;
                Q += alen;
            } else if (ty == VD_TYPE_DEL) {                      // :179-189
                if (rl < 1 || R + rl > win) return -1;
                if (rptr) {
                    for (int k = 0; k < rl; k++) { rptr[R + k] = (PT)(Q - 1); rflg[R + k] = P_VARIANT; }
                    rflg[R + rl - 1] |= P_VAR_END;
                    rflg[R] |= P_VAR_BEG;
                }
                R += rl; ref_pos += rl;
            } else if (ty == VD_TYPE_SUB) {                      // :190-198
                if (alen != 1 || rl != 1 || Q + 1 > cap) return -1;
                if (rptr) { rptr[R] = (PT)Q; rflg[R] = P_VARIANT | P_VAR_BEG | P_VAR_END; }
                ptr[Q] = (PT)R; flg[Q] = P_VARIANT | P_VAR_BEG | P_VAR_END;
                str[Q] = alt[0];
                R++; Q++; ref_pos++;
            } else {
                return -1;                                       // :199-201
            }
            v++;                                                 // :204
        } else {                                                 // :206-235
            const int ref_end = (v < ve) ? src.pos(v) : win;
            if (ref_end < ref_pos || ref_end > win || Q + (ref_end - ref_pos) > cap) return -1;
            const int n = ref_end - ref_pos;
            for (int k = 0; k < n; k++) {
                ptr[Q + k] = (PT)(R + k); flg[Q + k] = 0;
                if (rptr) { rptr[R + k] = (PT)(Q + k); rflg[R + k] = 0; }
                str[Q + k] = (u8)src.fa(ref_pos + k);
            }
            Q += n; R += n; ref_pos = ref_end;
        }
    }
    return Q;
}

template <class PT>
__device__ int expand_hap(const BatchDev &in, int sc, int h, u8 *str, u8 *flg, PT *ptr,
                          PT *rptr, u8 *rflg, u8 *ins, int cap) {
    return expand_hap_src<PT>(HapSrcGlobal(in, sc, h), str, flg, ptr, rptr, rflg, ins, cap);
}

// swap-source table of one destination plane, CSR: tab[0..ndst] = offsets, then the source
// rows.  Sources of destination row a are the rows b of the other plane with src(b) and
// ptr[b]+1 == a (:335-337, :364-366); ptr is non-decreasing in b, so they are contiguous and
// ascending.  A base before an insertion plus the insertion's last base give two sources; k
// back-to-back deletions give k+1.  The chosen source index is kept in 3 bits of the flag
// byte, so more than SW_MAX sources per row are rejected (VD_E_BADINPUT).
constexpr int SW_MAX = 8;
template <class PT>
__device__ bool build_swsrc(const PT *ptr, const u8 *flg, int nsrc, PT *tab, int ndst) {
    PT *off = tab, *src = tab + ndst + 1;
    for (int a = 0; a <= ndst; a++) off[a] = 0;
    int n = 0;
    for (int b = 0; b < nsrc; b++) {
        const int f = flg[b];
        if ((f & P_VARIANT) && !(f & P_VAR_END)) continue;
        const int d = (int)ptr[b] + 1;
        if (d < 0 || d >= ndst) continue;
        off[d + 1] = (PT)((int)off[d + 1] + 1);
        src[n++] = (PT)b;
    }
    bool ok = true;
    for (int a = 0; a < ndst; a++) {
        if ((int)off[a + 1] > SW_MAX) ok = false;
        off[a + 1] = (PT)((int)off[a + 1] + (int)off[a]);
    }
    return ok;
}

// ---------------------------------------------------------------------------------------
// plain Levenshtein, what wf_ed (:1406-1506) returns in `s`.
// ---------------------------------------------------------------------------------------
// Bit-parallel unit-cost global edit distance (Myers 1999 in Hyyro's formulation for the global
// distance: the horizontal delta shifted into row 0 is +1).  b is the pattern, n <= 64; a is the
// text.  O(m) word operations, no memory traffic except reading the strings: the section edit
// distances of structural variants are (thousands) x (a few) and made the one-thread row DP the
// slowest thing in the whole long-supercluster path.
__host__ __device__ inline int lev_myers64(const u8 *a, int m, const u8 *b, int n) {
    typedef unsigned long long u64;
    u64 eqA = 0, eqC = 0, eqG = 0, eqT = 0;
    for (int j = 0; j < n; j++) {
        const u64 bit = 1ull << j;
        const int c = b[j];
        if (c == 'A') eqA |= bit; else if (c == 'C') eqC |= bit; else if (c == 'G') eqG |= bit; else if (c == 'T') eqT |= bit;
    }
    const u64 top = 1ull << (n - 1);
    u64 VP = ~0ull, VN = 0;
    int score = n;
    for (int i = 0; i < m; i++) {
        const int c = a[i];
        u64 Eq;
        if (c == 'A') Eq = eqA; else if (c == 'C') Eq = eqC; else if (c == 'G') Eq = eqG; else if (c == 'T') Eq = eqT;
        else { Eq = 0; for (int j = 0; j < n; j++) if (b[j] == c) Eq |= 1ull << j; }     // N / IUPAC: rare
        const u64 X = Eq | VN;
        const u64 D0 = ((VP + (X & VP)) ^ VP) | X;
        const u64 HN = VP & D0;
        const u64 HP = VN | ~(VP | D0);
        if (HP & top) score++; else if (HN & top) score--;
        const u64 Xs = (HP << 1) | 1ull;
        VN = Xs & D0;
        VP = (HN << 1) | ~(Xs | D0);
    }
    return score;
}

// Row buffer in `mem` at oLev for the general case.
template <class Mem, int W>
__device__ int lev_scalar(const Mem &mem, typename Mem::off_t oLev,
                          const u8 *a, int m, const u8 *b, int n) {
    if (!m) return n;                                            // :1417
    if (!n) return m;                                            // :1418
    if (m == n) {                                                // identical strings: 0
        int k = 0;
        while (k < m && a[k] == b[k]) k++;
        if (k == m) return 0;
    }
    if (n > m) { const u8 *t = a; a = b; b = t; int x = m; m = n; n = x; }   // row over the shorter
    if (n <= 64 && m * n > 64) return lev_myers64(a, m, b, n);
    typedef Val<Mem, W> V;
    for (int j = 0; j <= n; j++) V::st(mem, oLev, j, j);
    for (int i = 1; i <= m; i++) {
        int diag = V::ld(mem, oLev, 0);
        int left = i;
        V::st(mem, oLev, 0, i);
        const int ai = a[i - 1];
        for (int j = 1; j <= n; j++) {
            const int up = V::ld(mem, oLev, j);
            int best = diag + (ai != b[j - 1]);
            best = min(best, up + 1);
            best = min(best, left + 1);
            V::st(mem, oLev, j, best);
            diag = up; left = best;
        }
    }
    return V::ld(mem, oLev, n);
}

// ---------------------------------------------------------------------------------------
// forward pass, calc_prec_recall_aln :251-443, as the dense column sweep of SURVEY.md 8a.
// Rows 0..Lq-1 are the QUERY plane, Lq..Lq+Lr-1 the REF plane; F is [column][row].
// ---------------------------------------------------------------------------------------
template <class Mem, int W, class PT>
__device__ void forward_scalar(const Mem &mem, const AlnLayout<typename Mem::off_t> &L,
                               const Hap<PT> &q, const QMaps<PT> &qm, const Hap<PT> &t,
                               const u8 *rseq, int Lr, int &score, int &end_plane) {
    typedef Val<Mem, W> V;
    typedef typename Mem::off_t off_t;
    const int Lq = q.len, Lt = t.len, N = Lq + Lr;
    off_t oPrev = L.oD0, oCur = L.oD1;
    for (int c = 0; c < Lt; c++) {
        const int tch = t.str[c];
        const bool tok = c > 0 && (!(t.flg[c - 1] & P_VARIANT) || (t.flg[c - 1] & P_VAR_END));   // :338-339
        int up_prev = 0;       // D[row-1][c-1]
        int up_cur = 0;        // D[row-1][c]
        for (int row = 0; row < N; row++) {
            const bool P = row >= Lq;
            const int a = P ? row - Lq : row;
            const int dprev = c > 0 ? V::ld(mem, oPrev, row) : INF;
            int d, f;
            if (a == 0 && c == 0) { d = 0; f = F_DIAG; }                                     // :299-305
            else {
                const bool m = (P ? rseq[a] : q.str[a]) == tch;
                const int diag = (a > 0 && c > 0) ? up_prev + (m ? 0 : 1) : INF;             // :324-332, :415-422
                const int ins = a > 0 ? up_cur + 1 : INF;                                    // :397-404
                const int del = c > 0 ? dprev + 1 : INF;                                     // :406-413
                int swp = INF, sbits = 0;
                if (tok && m) {                                                              // :334-349, :363-378
                    const PT *tab = P ? qm.toR : qm.toQ;
                    const PT *src = tab + (P ? Lr : Lq) + 1;
                    const int ob = P ? 0 : Lq;
                    const int k0 = tab[a], k1 = tab[a + 1];
                    for (int k = k0; k < k1; k++) {
                        const int v = V::ld(mem, oPrev, ob + (int)src[k]);
                        if (v < swp) { swp = v; sbits = (k - k0) << F_K_SHIFT; }
                        else if (v == swp) sbits = ((k - k0) << F_K_SHIFT) | F_TIE;   // keep the larger row
                    }
                }
                d = min(min(diag, ins), min(del, swp));
                f = 0;
                if (diag == d) f |= F_DIAG;
                if (ins == d) f |= F_INS;
                if (del == d) f |= F_DEL;
                if (swp == d) f |= F_SWP | sbits;
            }
            V::st(mem, oCur, row, d);
            mem.st8(L.oF + (off_t)c * N + row, f);
            up_prev = dprev;
            up_cur = d;
        }
        off_t x = oPrev; oPrev = oCur; oCur = x;
    }
    const int dq = V::ld(mem, oPrev, Lq - 1), dr = V::ld(mem, oPrev, N - 1);                // :390-391
    score = min(dq, dr);
    end_plane = (dq == score) ? 0 : 1;                                                        // :436-440
}

// ---------------------------------------------------------------------------------------
// backward pass, calc_prec_recall_path :486-834, reverse column sweep.  Returns the origin
// plane (:811-814); ORs VD_ST_TIE into status when an ambiguous swap edge is followed.
// ---------------------------------------------------------------------------------------
template <class Mem, int W, class PT>
__device__ int backward_scalar(const Mem &mem, const AlnLayout<typename Mem::off_t> &L,
                               const Hap<PT> &q, const QMaps<PT> &qm, const Hap<PT> &t,
                               const u8 *rseq, int Lr, int end_plane, u32 &status) {
    typedef Val<Mem, W> V;
    typedef typename Mem::off_t off_t;
    const int Lq = q.len, Lt = t.len, N = Lq + Lr;
    off_t oCur = L.oT0, oNext = L.oT1;          // T of column c, and of column c-1 being built
    for (int r = 0; r < N; r++) V::st(mem, oCur, r, -1);
    {
        const int erow = end_plane ? N - 1 : Lq - 1;
        V::st(mem, oCur, erow, 0);                                                            // :543-545
        mem.st8(L.oPF + (off_t)(Lt - 1) * N + erow, PTR_MAT);
    }
#define VD_RELAX(obase, col, yrow, val, ty) do {                                  \
        const int cur_ = V::ld(mem, obase, yrow);                                   \
        const off_t pi_ = L.oPF + (off_t)(col) * N + (yrow);                        \
        if ((val) > cur_) { V::st(mem, obase, yrow, (val)); mem.st8(pi_, (ty)); }   \
        else if ((val) == cur_) mem.st8(pi_, mem.ld8(pi_) | (ty)); } while (0)
    for (int c = Lt - 1; c >= 0; c--) {
        if (c > 0) for (int r = 0; r < N; r++) V::st(mem, oNext, r, -1);
        for (int row = N - 1; row >= 0; row--) {
            const int tx = V::ld(mem, oCur, row);
            if (tx < 0) continue;
            const bool P = row >= Lq;
            const int a = P ? row - Lq : row;
            const int f = mem.ld8(L.oF + (off_t)c * N + row);
            int tp = 0;                                                                       // :572-574
            if (!P && a > 0) tp = ((int)q.ptr[a] != (int)q.ptr[a - 1] + 1) || (q.flg[a] & P_VAR_BEG);
            if ((f & F_DIAG) && a > 0 && c > 0) {                                              // :556-595, :692-731
                const bool m = (P ? rseq[a] : q.str[a]) == t.str[c];
                VD_RELAX(oNext, c - 1, row - 1, tx + tp, m ? PTR_MAT : PTR_SUB);
            }
            if ((f & F_SWP) && a > 0 && c > 0) {                                              // :598-679
                const int of = P ? qm.rflg[a] : q.flg[a];
                if (!(of & P_VARIANT) || (of & P_VAR_BEG)) {
                    const PT *tab = P ? qm.toR : qm.toQ;
                    const PT *src = tab + (P ? Lr : Lq) + 1;
                    const int zrow = (P ? 0 : Lq) + (int)src[(int)tab[a] + (f >> F_K_SHIFT)];
                    if (f & F_TIE) status |= VD_ST_TIE;
                    VD_RELAX(oNext, c - 1, zrow, tx + (P ? 0 : tp), PTR_SWP);
                }
            }
            if ((f & F_INS) && a > 0) VD_RELAX(oCur, c, row - 1, tx + tp, PTR_INS);            // :734-771
            if ((f & F_DEL) && c > 0) VD_RELAX(oNext, c - 1, row, tx, PTR_DEL);                // :774-804
        }
        if (c > 0) { off_t x = oCur; oCur = oNext; oNext = x; }
    }
#undef VD_RELAX
    return V::ld(mem, oCur, 0) >= 0 ? 0 : 1;                                                  // :811-814
}

// ---------------------------------------------------------------------------------------
// Flag-matrix reader for the walk: PF byte of (plane, row, column).
// ---------------------------------------------------------------------------------------
template <class Mem> struct PFScalar {          // [column][row], rows = Lq + Lr
    const Mem *mem; typename Mem::off_t oPF; int N, Lq;
    __device__ __forceinline__ int get(int hi, int qri, int ti) const {
        return mem->ld8(oPF + (typename Mem::off_t)ti * N + (hi ? Lq + qri : qri));
    }
    __device__ __forceinline__ void prefetch(int, int, int) const {}
};

// ---------------------------------------------------------------------------------------
// walk (get_prec_recall_path_sync :842-999) + integer credit (calc_prec_recall :1005-1401).
// ---------------------------------------------------------------------------------------
#ifdef VD_PHASE_PROF
template <class A, class B> struct vd_same { static constexpr int v = 0; };
template <class A> struct vd_same<A, A> { static constexpr int v = 1; };
__device__ unsigned long long g_walk_phase[8];      // walk loop, credit loop, walk steps, sync sections (thread 0 of a block only)
#endif
template <class Mem, int W, class PT, class PFR>
__device__ void walk_credit(const Mem &mem, const AlnLayout<typename Mem::off_t> &L, const PFR &pfr,
                            const Hap<PT> &q, const QMaps<PT> &qm, const Hap<PT> &t,
                            const u8 *rseq, int Lr, int beg_plane, int end_plane,
                            const BatchDev &in, const OutDev &out, int sc, int ai, u32 &status) {
    typedef Val<Mem, W> V;
    constexpr int HB = (W == 2) ? 14 : 30;      // bit holding the plane in a packed path entry
    const int Lq = q.len, Lt = t.len;
#ifdef VD_PHASE_PROF
    long long ph_w0 = clock64();
    int ph_nsync = 0;
    constexpr int ph_o = vd_same<PFR, PFScalar<Mem>>::v ? 0 : 4;     // thread-per-alignment kernels / everything else
#endif
    const int maxpath = Lq + Lr + Lt + 3;
    int np = 0, ns = 0;
    int last_edit = 0;          // edits[] entry of the move that hit the :941 break, if any
    {
        int hi = beg_plane, qri = 0, ti = 0;
        V::st(mem, L.oPQ, 0, qri | (hi << HB)); V::st(mem, L.oPT, 0, ti);
        mem.st8(L.oPS, 1);                                                                    // :897-899
        np = 1; ns = 1;
        // The move is picked with selects and the loop is left through `go` only: the lanes of a warp walk different
        // alignments, and a chain of branches with early exits would keep them apart for the rest of each iteration.
        bool go = true, err = false;
        while (go && ((hi == 1 && qri < Lr - 1) || (hi == 0 && qri < Lq - 1) || ti < Lt - 1)) {  // :905
            const int pf = pfr.get(hi, qri, ti);
            const bool sw_r = hi == 1 && (pf & PTR_SWP);                                        // :907-912
            const int ty = sw_r ? PTR_SWP : (pf & PTR_MAT) ? PTR_MAT : (pf & PTR_SUB) ? PTR_SUB      // :914-920
                         : (pf & PTR_INS) ? PTR_INS : (pf & PTR_DEL) ? PTR_DEL                  // :922-928
                         : (hi == 0 && (pf & PTR_SWP)) ? PTR_SWP : 0;                           // :930-934, else :936-939
            const bool sw_q = !sw_r && ty == PTR_SWP;
            const int ed = (ty & (PTR_SUB | PTR_INS | PTR_DEL)) ? 1 : 0;
            int jump = 0;
            if (sw_r) jump = (int)qm.rptr[qri] + 1;
            if (sw_q) jump = (int)q.ptr[qri] + 1;
            qri = (ty == PTR_SWP) ? jump : qri + ((ty & (PTR_MAT | PTR_SUB | PTR_INS)) ? 1 : 0);
            ti += (ty & (PTR_MAT | PTR_SUB | PTR_DEL | PTR_SWP)) ? 1 : 0;
            hi = sw_r ? 0 : (sw_q ? 1 : hi);
            const bool past = (hi == 0 && qri >= Lq) || (hi == 1 && qri >= Lr) || ti >= Lt;       // :941
            if (!ty || (!past && np >= maxpath)) { err = true; go = false; continue; }
            if (past) { last_edit = ed; ns++; go = false; continue; }
            pfr.prefetch(hi, qri, ti);                  // flag bytes a few steps ahead (HBM-resident matrices only)
            const int tf = t.flg[ti];
            bool in_truth_var = tf & P_VARIANT;                                                 // :949-951
            if (ty & (PTR_MAT | PTR_SWP | PTR_SUB | PTR_DEL)) in_truth_var = in_truth_var && !(tf & P_VAR_BEG);
            bool in_query_var = false;                                                          // :953-956
            if (hi == 0) {
                const int qf = q.flg[qri];
                in_query_var = qf & P_VARIANT;
                if (ty & (PTR_MAT | PTR_SWP | PTR_SUB | PTR_DEL)) in_query_var = in_query_var && !(qf & P_VAR_BEG);
            }
            const int tref = (int)t.ptr[ti];
            const int qref = (hi == 1) ? qri : (int)q.ptr[qri];
            const bool is_ins_loc =                                                             // :958-960
                (tref >= 0 && tref < Lr && (q.ins[tref] | t.ins[tref])) ||
                (qref >= 0 && qref < Lr && (q.ins[qref] | t.ins[qref]));
            const bool is_sync = !in_truth_var && !in_query_var && !is_ins_loc && tref == qref &&   // :964-967
                                 (ty & (PTR_MAT | PTR_SWP | PTR_SUB));
            V::st(mem, L.oPQ, np, qri | (hi << HB)); V::st(mem, L.oPT, np, ti);
            mem.st8(L.oPS + np, (is_sync ? 1 : 0) | (ed ? 2 : 0));
            np++; ns++;
        }
        if (err) { status |= VD_ST_ERR_NO_POINTER; return; }
    }
    // sync[] has np+1 entries (the last forced true, :995), edits[] index k <= np:
    //   k < np  -> bit 1 of path flag k;  k == np -> the :941 break move's edit, else false
    const bool broke = (ns == np + 1);
#ifdef VD_PHASE_PROF
    if (threadIdx.x == 0) { const long long n_ = clock64(); atomicAdd(&g_walk_phase[ph_o + 0], (unsigned long long)(n_ - ph_w0)); ph_w0 = n_; atomicAdd(&g_walk_phase[ph_o + 2], (unsigned long long)np); }
#endif

    const int swap = (ai == 1 || ai == 2);                                                     // :1037
    const int qh = ai >> 1, th = 2 + (ai & 1);
    const int64_t qb = in.var_off[4 * (int64_t)sc + qh], qe = in.var_off[4 * (int64_t)sc + qh + 1];
    const int64_t tb = in.var_off[4 * (int64_t)sc + th], te = in.var_off[4 * (int64_t)sc + th + 1];
    u8 *asg = out.assigned + (int64_t)swap * in.n_var;
    int32_t *sg = out.sync_group + (int64_t)swap * in.n_var;
    int32_t *red = out.ref_ed + (int64_t)swap * in.n_var;
    int32_t *qed = out.query_ed + (int64_t)swap * in.n_var;
    float *cq = out.callq + (int64_t)swap * in.n_var;

    int sync_group = 0;                                                                        // :1059
    int hi = end_plane;                                                                        // :1061
    int prev_hi = hi, prev_qri = (hi == 0 ? Lq : Lr) - 1, prev_ti = Lt - 1;                     // :1062-1070
    int prev_sync_ref_idx = Lr, prev_sync_truth_idx = Lt;                                      // :1066-1072
    int query_ed = 0;
    int64_t qvp = qe - 1, prev_qvp = qvp;                                                      // :1074-1077
    int q_pos = (qvp >= qb) ? in.var_pos[qvp] : 0;
    int64_t tvp = te - 1, prev_tvp = tvp;                                                      // :1078-1081
    int t_pos = (tvp >= tb) ? in.var_pos[tvp] : 0;
    int sync_idx = np;                                                                         // :1082

    // The reference's loop (:1136-1399) visits every path entry from the end; at a sync point it closes the section since
    // the previous one (:1190-1380).  Most sections are a run of matched bases between variants: nothing is assigned and
    // the section's edit distance is 0.  Those are closed inline; a section that needs the full treatment (variants to
    // credit, or strings that differ) stops the inner loop instead, so that the lanes of a warp - each on its own
    // alignment - reach the expensive part together rather than one after the other at their own path positions.
    bool more = true;
    while (more) {
        int sync_ref_idx = 0, sync_truth_idx = 0, rn = 0, tn = 0;
        bool full = false;
        while (sync_idx >= 0) {                                                                // :1136
            const int query_ref_pos = (prev_hi == 1) ? prev_qri : (int)q.ptr[prev_qri];        // :1139-1144
            while (query_ref_pos < q_pos && qvp >= qb) {                                       // :1147
                if (hi == 1) {                                                                 // :1157-1168
                    asg[qvp] = VD_ASSIGN_REF_FP;
                    sg[qvp] = sync_group++;
                    red[qvp] = 0; qed[qvp] = 0;
                    cq[qvp] = in.var_qual[qvp];
                }
                qvp--;
                q_pos = (qvp < qb) ? -1 : in.var_pos[qvp];                                     // :1175-1176
            }
            const int truth_ref_pos = (int)t.ptr[prev_ti];                                     // :1180
            while (truth_ref_pos < t_pos && tvp >= tb) {                                       // :1181-1187
                tvp--;
                t_pos = (tvp < tb) ? -1 : in.var_pos[tvp];
            }
            const bool is_sync = (sync_idx == np) ? true : (mem.ld8(L.oPS + sync_idx) & 1);
            if (is_sync) {                                                                     // :1190
                sync_ref_idx = query_ref_pos + 1;                                              // :1194
                sync_truth_idx = prev_ti + 1;                                                  // :1195
                rn = prev_sync_ref_idx - sync_ref_idx;                                         // substr clipping
                if (rn < 0 || sync_ref_idx + rn > Lr) rn = Lr - sync_ref_idx;
                tn = prev_sync_truth_idx - sync_truth_idx;
                if (tn < 0 || sync_truth_idx + tn > Lt) tn = Lt - sync_truth_idx;
                bool plain = prev_qvp == qvp && prev_tvp == tvp && rn == tn && rn <= 8;        // no variant passed, equal lengths
                for (int k = 0; k < rn && plain; k++) plain = rseq[sync_ref_idx + k] == t.str[sync_truth_idx + k];
                if (!plain) { full = true; break; }
                // ref_ed = 0 here (:1197-1199 on identical strings), so of :1203-1380 only these remain
                if (query_ed != 0) status |= VD_ST_WARN_QED_NOQUERY | VD_ST_WARN_QED_GT_REFED;  // :1207, :1211
                prev_sync_ref_idx = sync_ref_idx;
                prev_sync_truth_idx = sync_truth_idx;
                query_ed = 0;
            }
            if (sync_idx == np) query_ed += broke ? last_edit : 0;                             // :1382
            else query_ed += (mem.ld8(L.oPS + sync_idx) >> 1) & 1;
            sync_idx--;
            if (sync_idx < 0) break;
            hi = prev_hi;                                                                      // :1387-1392
            const int pq = V::ld(mem, L.oPQ, sync_idx);
            prev_qri = pq & ((1 << HB) - 1); prev_hi = (pq >> HB) & 1;
            prev_ti = V::ld(mem, L.oPT, sync_idx);
            // :1395-1398 reload q_pos / t_pos here; they already hold var_pos[qvp] / var_pos[tvp] (set where qvp and tvp
            // move), and are not looked at once qvp < qb / tvp < tb: no load from HBM on the per-entry path
        }
        more = full;
        if (full) {
#ifdef VD_PHASE_PROF
            ph_nsync++;
#endif
            int ref_ed = lev_scalar<Mem, W>(mem, L.oLev, rseq + sync_ref_idx, rn,               // :1197-1199
                                            t.str + sync_truth_idx, tn);
            if (prev_tvp == tvp && ref_ed != 0) status |= VD_ST_WARN_REFED_NOTRUTH;            // :1203
            if (prev_qvp == qvp && query_ed != ref_ed) status |= VD_ST_WARN_QED_NOQUERY;       // :1207
            if (query_ed > ref_ed) status |= VD_ST_WARN_QED_GT_REFED;                          // :1211
            if (ref_ed == 0 && tvp != prev_tvp) { status |= VD_ST_WARN_ZERO_REFED; ref_ed = 1; }   // :1219-1223
            float callq = in.max_qual;                                                         // :1284-1288
            for (int64_t v = prev_qvp; v > qvp; v--) callq = fminf(callq, in.var_qual[v]);
            for (int64_t v = prev_qvp; v > qvp; v--) {                                         // :1291-1322
                if (asg[v] == VD_ASSIGN_NONE) {
                    asg[v] = VD_ASSIGN_SYNC;
                    sg[v] = sync_group; red[v] = ref_ed; qed[v] = query_ed; cq[v] = callq;
                }
            }
            for (int64_t v = prev_tvp; v > tvp; v--) {                                         // :1325-1353
                asg[v] = VD_ASSIGN_SYNC;
                sg[v] = sync_group; red[v] = ref_ed; qed[v] = query_ed; cq[v] = callq;
            }
            if (qvp != prev_qvp || tvp != prev_tvp) sync_group++;                              // :1364-1367
            prev_qvp = qvp; prev_tvp = tvp;
            prev_sync_ref_idx = sync_ref_idx;
            prev_sync_truth_idx = sync_truth_idx;
            query_ed = 0;
            // the rest of this path entry (:1382-1398), then on with the next one
            if (sync_idx == np) query_ed += broke ? last_edit : 0;                             // :1382
            else query_ed += (mem.ld8(L.oPS + sync_idx) >> 1) & 1;
            sync_idx--;
            if (sync_idx < 0) more = false;
            else {
                hi = prev_hi;                                                                  // :1387-1392
                const int pq = V::ld(mem, L.oPQ, sync_idx);
                prev_qri = pq & ((1 << HB) - 1); prev_hi = (pq >> HB) & 1;
                prev_ti = V::ld(mem, L.oPT, sync_idx);
            }
        }
    }
#ifdef VD_PHASE_PROF
    if (threadIdx.x == 0) { atomicAdd(&g_walk_phase[ph_o + 1], (unsigned long long)(clock64() - ph_w0)); atomicAdd(&g_walk_phase[ph_o + 3], (unsigned long long)ph_nsync); }
#endif
}

// ---------------------------------------------------------------------------------------
// Homozygous superclusters: when the two query haplotypes carry the same variants (same
// qualities) and the two truth haplotypes too, the four alignments Q1T1, Q1T2, Q2T1, Q2T2 are
// the same problem.  It is solved once (as Q1T1: query-hap-1 / truth-hap-1 records of phasing
// slot 0) and the records are replicated to the other haplotype and the other slot — every
// (haplotype, slot) pair is written by exactly one of the four alignments (:1037-1048), all
// with these values.
// ---------------------------------------------------------------------------------------
__device__ inline void replicate_hom(const BatchDev &in, const OutDev &out, int sc) {
    const int64_t nv = in.n_var;
#pragma unroll 1
    for (int side = 0; side < 2; side++) {                // query haps 0,1 then truth haps 2,3
        const int64_t b1 = in.var_off[4 * (int64_t)sc + 2 * side], b2 = in.var_off[4 * (int64_t)sc + 2 * side + 1];
        const int n = (int)(b2 - b1);
        for (int j = 0; j < n; j++) {
            const int64_t s = b1 + j, d = b2 + j;
            const u8 a = out.assigned[s];
            const int32_t g = out.sync_group[s], r = out.ref_ed[s], q = out.query_ed[s];
            const float c = out.callq[s];
            out.assigned[d] = a; out.sync_group[d] = g; out.ref_ed[d] = r; out.query_ed[d] = q; out.callq[d] = c;
            out.assigned[nv + s] = a; out.sync_group[nv + s] = g; out.ref_ed[nv + s] = r; out.query_ed[nv + s] = q; out.callq[nv + s] = c;
            out.assigned[nv + d] = a; out.sync_group[nv + d] = g; out.ref_ed[nv + d] = r; out.query_ed[nv + d] = q; out.callq[nv + d] = c;
        }
    }
}

}  // namespace vd
