// Warp-cooperative setup of a long supercluster: one WARP per (supercluster, haplotype) expands the haplotype
// (generate_ptrs_strs, src/dist.cpp:145-242), builds the CSR swap-source tables and every per-row table of the
// long-path kernels.  Same outputs, bit for bit, as the one-thread-per-haplotype kernels it replaces
// (slab_setup_kernel + wave_tables_kernel: a 10 kb haplotype is ~15 dependent passes over 20 k rows there,
// which made the setup the longest phase of an SV-bearing batch): every pass is a map or a scan over rows -
// lanes stride over the rows, scans carry across chunks of 32 - except the row-step hulls
// (wave_row_hulls), a two-pointer walk that lane 0 does alone.
#pragma once
#include "vd_wave.cuh"

namespace vd {

constexpr unsigned SETUP_FULL = 0xffffffffu;

// generate_ptrs_strs by a warp: the loop over variants is serial (a handful), the copies are parallel.
// Returns the haplotype length, -1 on input the reference cannot process (all lanes agree).
__device__ inline int expand_hap_warp(const BatchDev &in, int sc, int h, u8 *str, u8 *flg, int *ptr, int *rptr, u8 *rflg, u8 *ins,
                                      int cap, int lane) {
    const int64_t r0 = in.ref_off[sc];
    const int win = (int)(in.ref_off[sc + 1] - r0);
    const u8 *fa = in.ref_seq + r0;
    int64_t v = in.var_off[4 * (int64_t)sc + h];
    const int64_t ve = in.var_off[4 * (int64_t)sc + h + 1];
    int Q = 0, R = 0, ref_pos = 0;
    for (int k = lane; k < win; k += 32) ins[k] = 0;
    __syncwarp();
    while (ref_pos < win) {                                      // :163
        if (v < ve && ref_pos == in.var_pos[v]) {                // :165-166
            const int64_t a0 = in.alt_off[v];
            const int alen = (int)(in.alt_off[v + 1] - a0);
            const int rl = in.var_rlen[v];
            const int ty = in.var_type[v];
            if (ty == VD_TYPE_INS) {                             // :169-178
                if (alen < 1 || Q + alen > cap) return -1;
                for (int k = lane; k < alen; k += 32) {
                    ptr[Q + k] = R - 1;
                    flg[Q + k] = (u8)(P_VARIANT | (k == alen - 1 ? P_VAR_END : 0) | (k == 0 ? (P_VAR_BEG | P_INS_LOC) : 0));
                    str[Q + k] = in.alt_seq[a0 + k];
                }
                if (lane == 0 && R - 1 >= 0) ins[R - 1] = 1;
                Q += alen;
            } else if (ty == VD_TYPE_DEL) {                      // :179-189
                if (rl < 1 || R + rl > win) return -1;
                if (rptr) {
                    for (int k = lane; k < rl; k += 32) {
                        rptr[R + k] = Q - 1;
                        rflg[R + k] = (u8)(P_VARIANT | (k == rl - 1 ? P_VAR_END : 0) | (k == 0 ? P_VAR_BEG : 0));
                    }
                }
                R += rl; ref_pos += rl;
            } else if (ty == VD_TYPE_SUB) {                      // :190-198
                if (alen != 1 || rl != 1 || Q + 1 > cap) return -1;
                if (lane == 0) {
                    if (rptr) { rptr[R] = Q; rflg[R] = P_VARIANT | P_VAR_BEG | P_VAR_END; }
                    ptr[Q] = R; flg[Q] = P_VARIANT | P_VAR_BEG | P_VAR_END;
                    str[Q] = in.alt_seq[a0];
                }
                R++; Q++; ref_pos++;
            } else {
                return -1;                                       // :199-201
            }
            v++;                                                 // :204
        } else {                                                 // :206-235
            const int ref_end = (v < ve) ? in.var_pos[v] : win;
            if (ref_end < ref_pos || ref_end > win || Q + (ref_end - ref_pos) > cap) return -1;
            const int n = ref_end - ref_pos;
            for (int k = lane; k < n; k += 32) {
                ptr[Q + k] = R + k; flg[Q + k] = 0;
                if (rptr) { rptr[R + k] = Q + k; rflg[R + k] = 0; }
                str[Q + k] = fa[ref_pos + k];
            }
            Q += n; R += n; ref_pos = ref_end;
        }
    }
    __syncwarp();
    return Q;
}

// CSR swap-source table by a warp (build_swsrc of vd_scalar.cuh): tab[0..ndst] offsets, then the admitted
// source rows in ascending order.  Returns false when a destination has more than SW_MAX sources.
__device__ inline bool build_swsrc_warp(const int *ptr, const u8 *flg, int nsrc, int *tab, int ndst, int lane) {
    int *off = tab, *src = tab + ndst + 1;
    for (int a = lane; a <= ndst; a += 32) off[a] = 0;
    __syncwarp();
    int n = 0;                                                   // admitted so far (uniform)
    for (int b0 = 0; b0 < nsrc; b0 += 32) {
        const int b = b0 + lane;
        bool adm = false;
        int d = -1;
        if (b < nsrc) {
            const int f = flg[b];
            d = ptr[b] + 1;
            adm = (!(f & P_VARIANT) || (f & P_VAR_END)) && d >= 0 && d < ndst;
        }
        const unsigned m = __ballot_sync(SETUP_FULL, adm);
        if (adm) {
            src[n + __popc(m & ((1u << lane) - 1u))] = b;
            atomicAdd(&off[d + 1], 1);
        }
        n += __popc(m);
    }
    __syncwarp();
    // exclusive prefix sum of the counts, in place; a count above SW_MAX invalidates the supercluster
    bool ok = true;
    int carry = 0;
    for (int a0 = 1; a0 <= ndst; a0 += 32) {
        const int a = a0 + lane;
        const int c = a <= ndst ? off[a] : 0;
        if (c > SW_MAX) ok = false;
        int incl = c;
#pragma unroll
        for (int dd = 1; dd < 32; dd <<= 1) {
            const int o = __shfl_up_sync(SETUP_FULL, incl, dd);
            if (lane >= dd) incl += o;
        }
        if (a <= ndst) off[a] = carry + incl;
        carry += __shfl_sync(SETUP_FULL, incl, 31);
    }
    __syncwarp();
    return __all_sync(SETUP_FULL, ok);
}

// srcinfo (build_srcinfo of vd_wave.cuh) by a warp: zero everywhere, then one entry per admitted source
__device__ inline void build_srcinfo_warp(const int *ptr, int nsrc, const int *tab, int ndst, const u8 *dflg, const int *dptr, int *info, int lane) {
    for (int b = lane; b < nsrc; b += 32) info[b] = 0;
    __syncwarp();
    const int *src = tab + ndst + 1;
    const int total = tab[ndst];
    for (int n = lane; n < total; n += 32) {
        const int b = src[n];
        const int d = ptr[b] + 1;
        const int k = n - tab[d];
        const int df = dflg[d];
        int v = 0;
        if (d > 0 && (!(df & P_VARIANT) || (df & P_VAR_BEG))) {
            int tp = 0;
            if (dptr) tp = (dptr[d] != dptr[d - 1] + 1) || (df & P_VAR_BEG);
            v = 1 | (k << 1) | (tp << 4) | (d << 8);
        }
        info[b] = v;
    }
    __syncwarp();
}

// one warp per (entry, haplotype); CLS_WAVE superclusters only (the scalar-slab class keeps slab_setup_kernel)
__global__ void __launch_bounds__(128) long_setup_kernel(BatchDev in, const ScPlan *plan, const int *list, int i0, int i1,
                                                         const int64_t *offs, u8 *slab, int *hap_ok) {
    const int lane = threadIdx.x & 31, h = threadIdx.x >> 5;
    const int i = i0 + blockIdx.x;
    if (i >= i1) return;
    const int sc = list[i];
    const ScPlan p = plan[sc];
    if (p.cls != CLS_WAVE) return;
    int *okp = hap_ok + 4 * (int64_t)(i - i0);
    if (p.hom && (h & 1)) { if (lane == 0) okp[h] = 1; return; }      // same as haplotype h-1, never read
    const WaveSlab W = make_wave_slab(p);
    u8 *base = slab + (offs[i] - offs[i0]);
    const int L = p.len[h], Lr = p.lr;
    SlabHap H(base + W.base.hap[h], L, Lr);
    if (h >= 2) {                                                     // truth haplotype: expansion + tinfo
        const int len = expand_hap_warp(in, sc, h, H.str, H.flg, H.ptr, nullptr, nullptr, H.ins, L, lane);
        const bool ok = len == L;
        if (lane == 0) okp[h] = ok ? 1 : 0;
        if (!ok) return;
        u8 *tinfo = base + W.ht[h - 2];
        for (int c = lane; c < L; c += 32) {
            const bool tok = c > 0 && (!(H.flg[c - 1] & P_VARIANT) || (H.flg[c - 1] & P_VAR_END));   // :338-339
            tinfo[c] = (u8)((H.str[c] & 0x7f) | (tok ? 0x80 : 0));
        }
        return;
    }
    // ---- query haplotype ----
    SlabQm M(base + W.base.qm[h], L, Lr);
    WaveHapQ X(base + W.hq[h], L, Lr);
    const int len = expand_hap_warp(in, sc, h, H.str, H.flg, H.ptr, M.rptr, M.rflg, H.ins, L, lane);
    bool ok = len == L;
    if (ok) ok = build_swsrc_warp(H.ptr, H.flg, L, M.toR, Lr, lane);
    if (ok) ok = build_swsrc_warp(M.rptr, M.rflg, Lr, M.toQ, L, lane);
    if (lane == 0) okp[h] = ok ? 1 : 0;
    if (!ok) return;
    const u8 *rseq = in.rplane_seq + in.ref_off[sc];
    // QUERY rows as sources -> destinations on REF (CSR toR); dest flags = rflg, no tp on REF
    build_srcinfo_warp(H.ptr, L, M.toR, Lr, M.rflg, nullptr, X.srcQ, lane);
    // REF rows as sources -> destinations on QUERY (CSR toQ); dest flags = hap flags, tp from hap ptrs
    build_srcinfo_warp(M.rptr, Lr, M.toQ, L, H.flg, H.ptr, X.srcR, lane);
    // tp(a) (:572-574) and, for now, the number of tp rows BELOW each row in BwdRow::tw
    int carry = 0;
    for (int a0 = 0; a0 < L; a0 += 32) {
        const int a = a0 + lane;
        const int tp = (a < L && a > 0 && ((H.ptr[a] != H.ptr[a - 1] + 1) || (H.flg[a] & P_VAR_BEG))) ? 1 : 0;
        const unsigned m = __ballot_sync(SETUP_FULL, tp);
        if (a < L) { X.tpb[a] = (u8)tp; X.bwdQ[a].tw = (u32)(carry + __popc(m & ((1u << lane) - 1u))); }
        carry += __popc(m);
    }
    __syncwarp();
    const int tp_total = carry;
    if (lane == 0) wave_row_hulls(H.ptr, H.flg, L, M.rptr, M.rflg, Lr, X.fwdQ, X.fwdR);
    for (int a = lane; a < L; a += 32) {
        const int k0 = M.toQ[a], k1 = M.toQ[a + 1];
        const u32 first = k1 > k0 ? (u32)M.toQ[L + 1 + k0] : 0u;
        X.swiQ[a] = k1 > k0 ? (first | ((u32)(k1 - k0) << 16)) : 0u;
        X.fwdQ[a].w0 = (k1 > k0 ? ((first & 0xfffffu) | ((u32)(k1 - k0) << 20)) : 0u) | ((u32)(H.str[a] & 0x7f) << 24);
        const u32 tpa = X.tpb[a], tpn = a + 1 < L ? X.tpb[a + 1] : 0;
        const int above = tp_total - (int)X.bwdQ[a].tw - (int)tpa;   // rows a' > a with tp(a')
        X.tps[a] = (u16)above;
        BwdRow r;
        r.si = (u32)X.srcQ[a];
        r.tw = (u32)above | (tpa << 24) | (tpn << 25);
        r.chn = a + 1 < L ? (u32)(H.str[a + 1] & 0x7f) : 0xffu;
        r.pad = 0;
        X.bwdQ[a] = r;
    }
    for (int a = lane; a < Lr; a += 32) {
        const int k0 = M.toR[a], k1 = M.toR[a + 1];
        const u32 first = k1 > k0 ? (u32)M.toR[Lr + 1 + k0] : 0u;
        X.swiR[a] = k1 > k0 ? (first | ((u32)(k1 - k0) << 16)) : 0u;
        X.fwdR[a].w0 = (k1 > k0 ? ((first & 0xfffffu) | ((u32)(k1 - k0) << 20)) : 0u) | ((u32)(rseq[a] & 0x7f) << 24);
        BwdRow r;
        r.si = (u32)X.srcR[a];
        r.tw = 0;                                                     // tp is zero on the REF plane
        r.chn = a + 1 < Lr ? (u32)(rseq[a + 1] & 0x7f) : 0xffu;
        r.pad = 0;
        X.bwdR[a] = r;
    }
}

}  // namespace vd
