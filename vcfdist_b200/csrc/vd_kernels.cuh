// Kernels of the precision/recall path (sm_100a):
//   plan_kernel        sizes, kernel choice (launch rank) and homozygosity of every supercluster, work list
//                      of the long ones
//   small_kernel<K>    fused: one thread per alignment, one quad per supercluster, all
//                      matrices in shared memory; three footprint classes, cost-sorted order
//   slab_size_kernel / slab_setup_kernel / slab_align_kernel
//                      thread-per-alignment path with matrices in an HBM slab
// The wavefront kernels for long superclusters are in vd_wave.cuh.
#pragma once
#include "vd_scalar.cuh"
#include "vd_warp.cuh"

namespace vd {

// ---- small-supercluster classes -----------------------------------------------------------
// One thread per alignment, one quad per supercluster, every matrix in shared memory.  The class
// fixes the static sizes of the per-supercluster shared area (haplotype strings / pointer maps,
// TL = max haplotype length, TR = max window length) and of the per-alignment private slice
// (CAP bytes).  Small classes have small footprints, so most of the batch (91 % of a WGS
// small-variant batch fits class 0) runs at high occupancy; bigger classes trade occupancy for
// keeping the alignment on chip.  Whatever fits none goes to the HBM-slab wavefront path.
constexpr int N_SMALL = 2;
constexpr int N_SBIN = 8;              // cost bins per class (log2 of the supercluster's cells)
#ifndef VD_MINB0
#define VD_MINB0 8
#endif
template <int K> struct SmallCfg;      // MINB: resident blocks per SM the register allocation must allow
template <> struct SmallCfg<0> { static constexpr int TL = 8,  TR = 8,  CAP = 128,  TPB = 128, SHIFT = 3, MINB = VD_MINB0; };
template <> struct SmallCfg<1> { static constexpr int TL = 16, TR = 20, CAP = 320,  TPB = 128, SHIFT = 5, MINB = 4; };

__host__ __device__ inline int small_tl(int k) { const int v[N_SMALL] = {SmallCfg<0>::TL, SmallCfg<1>::TL}; return v[k]; }
__host__ __device__ inline int small_tr(int k) { const int v[N_SMALL] = {SmallCfg<0>::TR, SmallCfg<1>::TR}; return v[k]; }
__host__ __device__ inline int small_cap(int k) { const int v[N_SMALL] = {SmallCfg<0>::CAP, SmallCfg<1>::CAP}; return v[k]; }
__host__ __device__ inline int small_shift(int k) { const int v[N_SMALL] = {SmallCfg<0>::SHIFT, SmallCfg<1>::SHIFT}; return v[k]; }

// keys of the order array: (small class, cost bin) first, then the warp kernel's (slots, smem bin);
// launch groups: one per small class, one per (slots, smem bin)
// Launch groups and ranks.  A group is one kernel launch: (small class k, homozygous bit) for the
// thread-per-alignment kernels, (slots, shared-memory bin, homozygous bit) for the warp kernel.
// The order array is the batch STABLY sorted by rank (cub radix sort), rank = the group's position
// with the cost bin of the small classes as minor key, so that a group is a contiguous range, the
// lanes of a warp get similar costs, and superclusters keep their batch order inside a rank (their
// inputs stay neighbours in HBM).
constexpr int N_GROUP0 = N_SMALL + WSC_MAXSLOT * N_WBIN;
constexpr int N_GROUP = 2 * N_GROUP0;
constexpr int N_KEY = 2 * N_SMALL * N_SBIN + 2 * WSC_MAXSLOT * N_WBIN;      // ranks
constexpr int RANK_NONE = 255;                                              // not a short supercluster
static_assert(N_KEY < RANK_NONE, "ranks are sorted as 8-bit keys");
// g0: group without the homozygous bit (small class k, or N_SMALL + (slots-1)*N_WBIN + reversed bin)
__host__ __device__ inline int rank_of(int g0, int hom, int sbin) {
    return g0 < N_SMALL ? (2 * g0 + hom) * N_SBIN + sbin : 2 * N_SMALL * N_SBIN + 2 * (g0 - N_SMALL) + hom;
}
__host__ __device__ inline int group_of_key(int rank) {
    return rank < 2 * N_SMALL * N_SBIN ? rank / N_SBIN : 2 * N_SMALL + (rank - 2 * N_SMALL * N_SBIN);
}

template <int TL, int TR> struct SmallDims {
    static constexpr int SW = (TL + TR + 1 + 3) & ~3;                 // one CSR swap table
    static constexpr int HAP = 3 * TL + TR;                           // str TL | flg TL | ptr TL | ins TR
    static constexpr int QM = 2 * TR + 2 * SW;                        // rptr TR | rflg TR | toQ SW | toR SW
    static constexpr int RAW = 4 * HAP + 2 * QM + TR + 8;             // ... | rseq TR | hlen 4 x int16
    // stride of the per-supercluster areas: an ODD number of words, so that the quads of a warp that
    // read the same offset of their own area hit different banks
    static constexpr int SC_BYTES = ((RAW / 4) & 1) ? RAW : RAW + 4;
    static_assert((TR % 4) == 0 && (TL % 4) == 0 && (RAW % 4) == 0, "alignment of the shared area");
};

struct PlanCounters {
    int n_list;                        // superclusters not handled by the small kernels
    int n_bad;
    unsigned long long cells;
    unsigned long long cells_list;
    unsigned status_or;
    int n_key[N_KEY];                  // superclusters per rank
    int grp_first[N_GROUP], grp_count[N_GROUP];
    unsigned long long io_grp[N_GROUP];     // algorithmic input+output bytes per launch group
};

// small control blocks go to the host through mapped pinned memory (stores from the SM), not through
// a copy engine: a D2H memcpy would wait behind the bulk result copies of the previous chunk
__global__ void publish_kernel(const u32 *src, u32 *dst_host, int n_words) {
    for (int i = threadIdx.x; i < n_words; i += blockDim.x) dst_host[i] = src[i];
    __threadfence_system();
}

// 16-bit result records (vd_packed_out): narrowing of one chunk's result arrays before the copy out
__global__ void pack_out_kernel(OutDev o, int64_t n_aln, int64_t n_var, u16 *score16, u8 *planes, u16 *status16,
                                u16 *sg16, u16 *red16, u16 *qed16, unsigned *range_flag_host) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    bool over = false;
    if (i < n_aln) {
        const int sc = o.aln_score[i];
        over |= sc >= 0xffff;
        score16[i] = (u16)(sc < 0 ? 0xffff : sc);
        planes[i] = (u8)((o.aln_end_plane[i] & 1) | ((o.aln_beg_plane[i] & 1) << 1));
        over |= (o.status[i] >> 16) != 0;
        status16[i] = (u16)o.status[i];
    }
    if (i < 2 * n_var) {
        const int sg = o.sync_group[i], r = o.ref_ed[i], q = o.query_ed[i];
        over |= sg < 0 || sg >= (1 << 14) || r < 0 || r >= 0xffff || q < 0 || q >= 0xffff;
        sg16[i] = (u16)(((int)o.assigned[i] << 14) | (sg & 0x3fff));
        red16[i] = (u16)r; qed16[i] = (u16)q;
    }
    if (over) { *range_flag_host = 1u; __threadfence_system(); }
}

// compact input (vd_compact_in): sizes widened to 64 bits where the prefix sums will stand (entry n = 0), positions
// and REF lengths to 32 bits
__global__ void unpack_sizes_kernel(const u16 *ref_len, int64_t ns, const u8 *hap_nvar, const u16 *alt_len, int64_t nv,
                                    int64_t *ref_off, int64_t *var_off, int64_t *alt_off,
                                    const u16 *pos16, const u16 *rlen16, int32_t *var_pos, int32_t *var_rlen) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i <= ns) ref_off[i] = i < ns ? ref_len[i] : 0;
    if (i <= 4 * ns) var_off[i] = i < 4 * ns ? hap_nvar[i] : 0;
    if (i <= nv) alt_off[i] = i < nv ? alt_len[i] : 0;
    if (i < nv) { var_pos[i] = pos16[i]; var_rlen[i] = rlen16[i]; }
}

// OR of all status words (so that the host only scans them when an error bit is set)
__global__ void status_or_kernel(const u32 *status, int64_t n, unsigned *dst) {
    unsigned v = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) v |= status[i];
    v = __reduce_or_sync(0xffffffffu, v);
    if ((threadIdx.x & 31) == 0 && v) atomicOr(dst, v);
}

__device__ __forceinline__ int small_need(int Lq, int Lr, int Lt) {
    return make_layout<int, 2, true>(Lq + Lr, Lt, Lr).total;
}

// One thread per supercluster.  small_lo / small_hi: range of small classes in use (testing hooks
// VD_SMALL_MIN / VD_SMALL_MAX; hi < lo disables the small kernels).
__global__ void plan_kernel(BatchDev in, OutDev out, ScPlan *plan, int *list, u8 *ranks, int *iota, PlanCounters *cnt, int force_class,
                            int big_class, int small_lo, int small_hi, int use_wsc, int use_hom) {
    const int sc0 = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = sc0 < in.n_sc;
    const int sc = live ? sc0 : in.n_sc - 1;      // dead lanes recompute the last one and discard it
    ScPlan p;
    const int lr = (int)(in.ref_off[sc + 1] - in.ref_off[sc]);
    p.lr = lr;
    bool bad = lr < 1;
    int maxlen = 0;
    for (int h = 0; h < 4; h++) {
        int len = lr, prev_end = 0;
        for (int64_t v = in.var_off[4 * (int64_t)sc + h]; v < in.var_off[4 * (int64_t)sc + h + 1]; v++) {
            const int alen = (int)(in.alt_off[v + 1] - in.alt_off[v]);
            const int rl = in.var_rlen[v], ty = in.var_type[v], pos = in.var_pos[v];
            len += alen - rl;
            if (ty < VD_TYPE_SUB || ty > VD_TYPE_DEL || pos < prev_end || pos + rl > lr || rl < 0) bad = true;
            prev_end = pos + rl;
        }
        p.len[h] = len;
        maxlen = max(maxlen, len);
        if (len < 1) bad = true;
    }
    // homozygous: query haps 0/1 carry the same variants with the same qualities, truth haps 2/3 the same variants
    bool hom = use_hom && !bad;
    for (int side = 0; side < 2 && hom; side++) {
        const int64_t b1 = in.var_off[4 * (int64_t)sc + 2 * side], b2 = in.var_off[4 * (int64_t)sc + 2 * side + 1];
        const int64_t e2 = in.var_off[4 * (int64_t)sc + 2 * side + 2];
        if (b2 - b1 != e2 - b2) { hom = false; break; }
        for (int64_t j = 0; j < b2 - b1 && hom; j++) {
            const int64_t x = b1 + j, y = b2 + j;
            const int64_t ax = in.alt_off[x], ay = in.alt_off[y];
            const int al = (int)(in.alt_off[x + 1] - ax);
            hom = in.var_pos[x] == in.var_pos[y] && in.var_type[x] == in.var_type[y] && in.var_rlen[x] == in.var_rlen[y] &&
                  al == (int)(in.alt_off[y + 1] - ay) &&
                  (side == 1 || __float_as_uint(in.var_qual[x]) == __float_as_uint(in.var_qual[y]));
            for (int k = 0; k < al && hom; k++) hom = in.alt_seq[ax + k] == in.alt_seq[ay + k];
        }
    }
    unsigned long long cells = 0;
    int cls = big_class;
    int sbin = -1;                                 // rank of the short kernels
    if (bad) cls = CLS_BAD;
    else {
        int need = 0;
        for (int ai = 0; ai < 4; ai++) {
            const int lq = p.len[ai >> 1], lt = p.len[2 + (ai & 1)];
            cells += (unsigned long long)(lq + lr) * lt;
            if (maxlen <= SmallCfg<N_SMALL - 1>::TL && lr <= SmallCfg<N_SMALL - 1>::TR) need = max(need, small_need(lq, lr, lt));
            else need = 1 << 30;
        }
        if (force_class > CLS_TINY) cls = force_class;
        else {
            for (int k = max(small_lo, 0); k <= small_hi && k < N_SMALL; k++) {
                if (maxlen <= small_tl(k) && lr <= small_tr(k) && need <= small_cap(k)) {
                    cls = CLS_TINY;
                    const int lg = 31 - __clz((int)cells | 1);
                    // bins are laid out most expensive first (the tail of a launch is its cheap work)
                    sbin = rank_of(k, hom, N_SBIN - 1 - min(max(lg - small_shift(k), 0), N_SBIN - 1));
                    break;
                }
            }
            if (sbin < 0 && use_wsc) {             // warp-per-supercluster kernel: four flag matrices in shared memory
                const int ws = wsc_slots(p);
                const int wb = ws > 0 ? wsc_bin(wsc_layout(p, hom).total) : -1;
                if (wb >= 0) { cls = CLS_TINY; sbin = rank_of(N_SMALL + (ws - 1) * N_WBIN + (N_WBIN - 1 - wb), hom, 0); }
            }
        }
    }
    p.cls = cls | (sbin >= 0 ? (sbin << 8) : 0);
    p.hom = hom ? 1 : 0;
    if (live) { plan[sc] = p; ranks[sc] = (u8)(sbin >= 0 ? sbin : RANK_NONE); iota[sc] = sc; }
    if (live && cls == CLS_BAD)                        // malformed supercluster: flagged per alignment, never computed
        for (int k = 0; k < 4; k++) { out.status[4 * (int64_t)sc + k] = VD_ST_ERR_BADINPUT; out.aln_score[4 * (int64_t)sc + k] = -1; }
    // Counters: aggregated per warp (shuffles, match), then per block in shared memory, then ONE global atomic per block and
    // counter - the batch's 14 k blocks would otherwise send half a million atomics to the same few addresses, which the L2
    // serialises (that, not the loads, was most of this kernel's time).
    __shared__ unsigned long long s_cells, s_io[N_GROUP];
    __shared__ unsigned s_key[N_KEY];
    for (int i = threadIdx.x; i < N_KEY; i += blockDim.x) s_key[i] = 0;
    for (int i = threadIdx.x; i < N_GROUP; i += blockDim.x) s_io[i] = 0;
    if (threadIdx.x == 0) s_cells = 0;
    __syncthreads();
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const bool is_bad = live && cls == CLS_BAD, is_list = live && (cls == CLS_WAVE || cls == CLS_SCALAR);
    unsigned long long c_all = live ? cells : 0ull, c_list = is_list ? cells : 0ull;
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        c_all += __shfl_down_sync(full, c_all, d);
        c_list += __shfl_down_sync(full, c_list, d);
    }
    const unsigned m_list = __ballot_sync(full, is_list), m_bad = __ballot_sync(full, is_bad);
    int base = 0;
    if (lane == 0) {
        if (c_all) atomicAdd(&s_cells, c_all);
        if (m_list) { base = atomicAdd(&cnt->n_list, __popc(m_list)); atomicAdd(&cnt->cells_list, c_list); }     // long superclusters: rare
        if (m_bad) atomicAdd(&cnt->n_bad, __popc(m_bad));
    }
    base = __shfl_sync(full, base, 0);
    if (is_list) list[base + __popc(m_list & ((1u << lane) - 1))] = sc;
    // per-bin counts of the small classes: one shared-memory atomic per distinct bin in the warp
    const int key = (live && sbin >= 0) ? sbin : -1;
    const unsigned peers = __match_any_sync(full, key);
    if (key >= 0 && lane == __ffs(peers) - 1) atomicAdd(&s_key[key], (unsigned)__popc(peers));
    {   // algorithmic bytes of each launch group's superclusters (io_bytes_of in vd_api.cu)
        const int grp = key >= 0 ? group_of_key(key) : -1;
        unsigned long long io = 0;
        if (grp >= 0) {
            const int64_t v0 = in.var_off[4 * (int64_t)sc], v1 = in.var_off[4 * (int64_t)sc + 4];
            const int64_t nv = v1 - v0;
            io = (unsigned long long)(lr + 8 + 32 + nv * (4 + 4 + 1 + 4 + 8) + (in.alt_off[v1] - in.alt_off[v0])
                                      + 4 * (4 + 1 + 1 + 4) + 2 * nv * 17);
        }
        const unsigned gp = __match_any_sync(full, grp);
        const int gl = __ffs(gp) - 1;
        unsigned long long tot = 0;
        for (unsigned m = gp; m; m &= m - 1) tot += __shfl_sync(gp, io, __ffs(m) - 1);
        if (grp >= 0 && lane == gl) atomicAdd(&s_io[grp], tot);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < N_KEY; i += blockDim.x) if (s_key[i]) atomicAdd(&cnt->n_key[i], (int)s_key[i]);
    for (int i = threadIdx.x; i < N_GROUP; i += blockDim.x) if (s_io[i]) atomicAdd(&cnt->io_grp[i], s_io[i]);
    if (threadIdx.x == 0 && s_cells) atomicAdd(&cnt->cells, s_cells);
}

// group ranges in the rank-sorted order array
__global__ void small_base_kernel(PlanCounters *cnt) {
    if (threadIdx.x != 0) return;
    int o = 0;
    for (int g = 0; g < N_GROUP; g++) cnt->grp_count[g] = 0;
    for (int r = 0; r < N_KEY; r++) {
        const int g = group_of_key(r);
        if (cnt->grp_count[g] == 0) cnt->grp_first[g] = o;
        o += cnt->n_key[r];
        cnt->grp_count[g] += cnt->n_key[r];
    }
}

// ---- fused small kernel ---------------------------------------------------------------------
// per-supercluster shared area (bytes): 4 x hap {str TL, flg TL, ptr TL, ins TR},
// 2 x qmaps {rptr TR, rflg TR, toQ SW, toR SW}, rseq TR, hlen 4 x int16
template <int TL, int TR> struct SmallArea {
    typedef SmallDims<TL, TR> D;
    u8 *base;
    __device__ u8 *str(int h) const { return base + h * D::HAP; }
    __device__ u8 *flg(int h) const { return str(h) + TL; }
    __device__ int8_t *ptr(int h) const { return (int8_t *)(str(h) + 2 * TL); }
    __device__ u8 *ins(int h) const { return str(h) + 3 * TL; }
    __device__ u8 *qm(int qh) const { return base + 4 * D::HAP + qh * D::QM; }
    __device__ int8_t *rptr(int qh) const { return (int8_t *)qm(qh); }
    __device__ u8 *rflg(int qh) const { return qm(qh) + TR; }
    __device__ int8_t *toQ(int qh) const { return (int8_t *)(qm(qh) + 2 * TR); }
    __device__ int8_t *toR(int qh) const { return (int8_t *)(qm(qh) + 2 * TR + D::SW); }
    __device__ u8 *rseq() const { return base + 4 * D::HAP + 2 * D::QM; }
    __device__ short *hlen() const { return (short *)(rseq() + TR); }
};

template <int K> struct SmallMem {
    static constexpr int STRIDE = SmallCfg<K>::CAP + 4;               // odd number of words -> conflict-free
    static constexpr int SMEM = (SmallCfg<K>::TPB / 4) * SmallDims<SmallCfg<K>::TL, SmallCfg<K>::TR>::SC_BYTES + SmallCfg<K>::TPB * STRIDE;
    static_assert(((STRIDE / 4) & 1) == 1, "private slice stride must be an odd number of words");
};

template <int K>
__global__ void __launch_bounds__(SmallCfg<K>::TPB, SmallCfg<K>::MINB)
small_kernel(BatchDev in, OutDev out, const ScPlan *__restrict__ plan, const int *__restrict__ order, int count) {
    typedef SmallCfg<K> C;
    typedef SmallDims<C::TL, C::TR> D;
    VD_DYN_SHARED(smem);
    constexpr int SPB = C::TPB / 4;
    const int tid = threadIdx.x, quad = tid >> 2, h = tid & 3;
    const int slot = blockIdx.x * SPB + quad;
    const bool active = slot < count;
    const int sc = active ? order[slot] : 0;
    SmallArea<C::TL, C::TR> A{smem + quad * D::SC_BYTES};
    SMemIL mem{smem + SPB * D::SC_BYTES + tid * SmallMem<K>::STRIDE};

    int lr = 0;
    bool novar = false;                                  // my haplotype carries no variant
    if (active) {
        lr = plan[sc].lr;
        // thread h expands haplotype h (generate_ptrs_strs); query haps also keep the ref side
        const bool isq = h < 2;
        const int len = expand_hap<int8_t>(in, sc, h, A.str(h), A.flg(h), A.ptr(h),
                                           isq ? A.rptr(h) : nullptr, isq ? A.rflg(h) : nullptr,
                                           A.ins(h), C::TL);
        bool ok = len == plan[sc].len[h];
        if (ok && isq) {
            ok = build_swsrc<int8_t>(A.ptr(h), A.flg(h), len, A.toR(h), lr) &&
                 build_swsrc<int8_t>(A.rptr(h), A.rflg(h), lr, A.toQ(h), len);
        }
        A.hlen()[h] = (short)(ok ? len : -1);
        novar = in.var_off[4 * (int64_t)sc + h + 1] == in.var_off[4 * (int64_t)sc + h];
        if (h == 2) {
            const u8 *rs = in.rplane_seq + in.ref_off[sc];
            for (int k = 0; k < lr; k++) A.rseq()[k] = rs[k];
        }
    }
    __syncwarp();
    // Which of the block's alignments there is something to compute for.  An alignment whose query and truth haplotypes
    // both carry no variant compares the window with itself (score 0, both path ends on the QUERY plane, nothing to
    // credit: the Q2T2 of every heterozygous site; see wsc_kernel) and a supercluster with a malformed haplotype is
    // flagged: both are written here.  The rest is compacted over the block, so that the sweeps run on full warps.
    const unsigned fullm = 0xffffffffu;
    const int lane = tid & 31, qbase = lane & ~3;
    const unsigned nv4 = (__ballot_sync(fullm, novar) >> qbase) & 15u;                        // bit h: haplotype h has no variant
    const short *hl = A.hlen();
    const bool hap_bad = active && (hl[0] < 0 || hl[1] < 0 || hl[2] < 0 || hl[3] < 0);
    bool todo = active && !hap_bad;
    if (active) {
        const int ai = h;
        if (hap_bad) {
            out.status[4 * (int64_t)sc + ai] = ST_BAD;
            out.aln_score[4 * (int64_t)sc + ai] = -1;
        } else if (in.rplane_seq == in.ref_seq && ((nv4 >> (ai >> 1)) & 1) && ((nv4 >> (2 + (ai & 1))) & 1)) {
            out.aln_score[4 * (int64_t)sc + ai] = 0;
            out.aln_end_plane[4 * (int64_t)sc + ai] = 0;
            out.aln_beg_plane[4 * (int64_t)sc + ai] = 0;
            out.status[4 * (int64_t)sc + ai] = 0;
            todo = false;
        }
    }
    __shared__ int s_wcnt[C::TPB / 32];
    __shared__ unsigned char s_task[C::TPB];
    const unsigned tm = __ballot_sync(fullm, todo);
    if (lane == 0) s_wcnt[tid >> 5] = __popc(tm);
    __syncthreads();
    int before = 0, ntask = 0;
    for (int w = 0; w < C::TPB / 32; w++) { if (w < (tid >> 5)) before += s_wcnt[w]; ntask += s_wcnt[w]; }
    if (todo) s_task[before + __popc(tm & ((1u << lane) - 1))] = (unsigned char)tid;
    __syncthreads();
    if (tid >= ntask) return;
    // ---- my task: alignment ai of the block's supercluster tq ----
    const int tq = s_task[tid] >> 2, ai = s_task[tid] & 3;
    const int tsc = order[blockIdx.x * SPB + tq];
    SmallArea<C::TL, C::TR> B{smem + tq * D::SC_BYTES};
    const int tlr = plan[tsc].lr;
    const int qh = ai >> 1, th = 2 + (ai & 1);
    u32 status = 0;
    const short *thl = B.hlen();
    Hap<int8_t> q{thl[qh], B.str(qh), B.flg(qh), B.ptr(qh), B.ins(qh)};
    Hap<int8_t> t{thl[th], B.str(th), B.flg(th), B.ptr(th), B.ins(th)};
    QMaps<int8_t> qm{B.rptr(qh), B.rflg(qh), B.toQ(qh), B.toR(qh)};
    const AlnLayout<int> L = make_layout<int, 2, true>(q.len + tlr, t.len, tlr);

    int score, end_plane;
    forward_scalar<SMemIL, 2, int8_t>(mem, L, q, qm, t, B.rseq(), tlr, score, end_plane);
    const int beg_plane = backward_scalar<SMemIL, 2, int8_t>(mem, L, q, qm, t, B.rseq(), tlr, end_plane, status);
    PFScalar<SMemIL> pfr{&mem, L.oPF, q.len + tlr, q.len};
    walk_credit<SMemIL, 2, int8_t>(mem, L, pfr, q, qm, t, B.rseq(), tlr, beg_plane, end_plane,
                                   in, out, tsc, ai, status);
    out.aln_score[4 * (int64_t)tsc + ai] = score;
    out.aln_end_plane[4 * (int64_t)tsc + ai] = (u8)end_plane;
    out.aln_beg_plane[4 * (int64_t)tsc + ai] = (u8)beg_plane;
    out.status[4 * (int64_t)tsc + ai] = status;
}


// Homozygous superclusters (replicate_hom): one THREAD per supercluster runs Q1T1 and copies the
// records; shared memory holds two haplotypes and one query map per supercluster.
template <int TL, int TR> struct SmallHomDims {
    static constexpr int SW = SmallDims<TL, TR>::SW;
    static constexpr int RAW = 2 * (3 * TL + TR) + (2 * TR + 2 * SW) + TR;
    static constexpr int SC_BYTES = ((RAW / 4) & 1) ? RAW : RAW + 4;      // odd number of words
};
template <int K> struct SmallHomMem {
    static constexpr int SMEM = SmallCfg<K>::TPB * (SmallHomDims<SmallCfg<K>::TL, SmallCfg<K>::TR>::SC_BYTES + SmallMem<K>::STRIDE);
};
template <int K>
__global__ void __launch_bounds__(SmallCfg<K>::TPB)
small_hom_kernel(BatchDev in, OutDev out, const ScPlan *__restrict__ plan, const int *__restrict__ order, int count) {
    typedef SmallCfg<K> C;
    typedef SmallHomDims<C::TL, C::TR> D;
    VD_DYN_SHARED(smem);
    const int tid = threadIdx.x;
    const int slot = blockIdx.x * C::TPB + tid;
    if (slot >= count) return;
    const int sc = order[slot];
    u8 *A = smem + tid * D::SC_BYTES;
    u8 *qstr = A, *qflg = A + C::TL, *qins = A + 3 * C::TL;
    int8_t *qptr = (int8_t *)(A + 2 * C::TL);
    u8 *T = A + 3 * C::TL + C::TR;
    u8 *tstr = T, *tflg = T + C::TL, *tins = T + 3 * C::TL;
    int8_t *tptr = (int8_t *)(T + 2 * C::TL);
    u8 *Q = T + 3 * C::TL + C::TR;
    int8_t *rptr = (int8_t *)Q, *toQ = (int8_t *)(Q + 2 * C::TR), *toR = (int8_t *)(Q + 2 * C::TR + D::SW);
    u8 *rflg = Q + C::TR;
    u8 *rseq = Q + 2 * C::TR + 2 * D::SW;
    SMemIL mem{smem + C::TPB * D::SC_BYTES + tid * SmallMem<K>::STRIDE};

    const int lr = plan[sc].lr;
    const int lq = expand_hap<int8_t>(in, sc, 0, qstr, qflg, qptr, rptr, rflg, qins, C::TL);
    bool ok = lq == plan[sc].len[0];
    if (ok) ok = build_swsrc<int8_t>(qptr, qflg, lq, toR, lr) && build_swsrc<int8_t>(rptr, rflg, lr, toQ, lq);
    const int lt = expand_hap<int8_t>(in, sc, 2, tstr, tflg, tptr, nullptr, nullptr, tins, C::TL);
    ok = ok && lt == plan[sc].len[2];
    {
        const u8 *rs = in.rplane_seq + in.ref_off[sc];
        for (int k = 0; k < lr; k++) rseq[k] = rs[k];
    }
    const int64_t oi = 4 * (int64_t)sc;
    if (!ok) {
        for (int k = 0; k < 4; k++) { out.status[oi + k] = ST_BAD; out.aln_score[oi + k] = -1; }
        return;
    }
    Hap<int8_t> q{lq, qstr, qflg, qptr, qins};
    Hap<int8_t> t{lt, tstr, tflg, tptr, tins};
    QMaps<int8_t> qm{rptr, rflg, toQ, toR};
    const AlnLayout<int> L = make_layout<int, 2, true>(lq + lr, lt, lr);
    u32 status = 0;
    int score, end_plane;
    forward_scalar<SMemIL, 2, int8_t>(mem, L, q, qm, t, rseq, lr, score, end_plane);
    const int beg_plane = backward_scalar<SMemIL, 2, int8_t>(mem, L, q, qm, t, rseq, lr, end_plane, status);
    PFScalar<SMemIL> pfr{&mem, L.oPF, lq + lr, lq};
    walk_credit<SMemIL, 2, int8_t>(mem, L, pfr, q, qm, t, rseq, lr, beg_plane, end_plane, in, out, sc, 0, status);
    for (int k = 0; k < 4; k++) {
        out.aln_score[oi + k] = score;
        out.aln_end_plane[oi + k] = (u8)end_plane;
        out.aln_beg_plane[oi + k] = (u8)beg_plane;
        out.status[oi + k] = status;
    }
    replicate_hom(in, out, sc);
}

template <int K> inline void small_configure_one() {
    cudaFuncSetAttribute(small_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, SmallMem<K>::SMEM);
    cudaFuncSetAttribute(small_hom_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, SmallHomMem<K>::SMEM);
}
inline void small_configure() { small_configure_one<0>(); small_configure_one<1>(); }
template <int K> inline void small_launch_one(cudaStream_t st, bool hom, const BatchDev &in, const OutDev &out, const ScPlan *plan,
                                              const int *order, int count) {
    constexpr int TPB = SmallCfg<K>::TPB, SPB = TPB / 4;
    if (hom) VD_LAUNCH(small_hom_kernel<K>, (count + TPB - 1) / TPB, TPB, SmallHomMem<K>::SMEM, st, in, out, plan, order, count);
    else VD_LAUNCH(small_kernel<K>, (count + SPB - 1) / SPB, TPB, SmallMem<K>::SMEM, st, in, out, plan, order, count);
}
inline void small_launch(cudaStream_t st, int k, bool hom, const BatchDev &in, const OutDev &out, const ScPlan *plan,
                         const int *order, int count) {
    if (count <= 0) return;
    switch (k) {
        case 0: small_launch_one<0>(st, hom, in, out, plan, order, count); break;
        case 1: small_launch_one<1>(st, hom, in, out, plan, order, count); break;
    }
}

// ---- HBM-slab path ---------------------------------------------------------------------------
// Expanded supercluster in HBM (32-bit pointers).  Offsets relative to the entry's slab.
struct SlabLayout {
    int64_t hap[4];      // str L | flg L | ins Lr | pad | ptr 4L
    int64_t qm[2];       // rflg Lr | pad | rptr 4Lr | toQ CSR 4(Lq+Lr+1) | toR CSR 4(Lq+Lr+1)
    int64_t rseq_unused;
    int64_t aln[4];      // scalar alignment scratch (make_layout<int64,4,false>) — slab path only
    int64_t total;
};

__host__ __device__ inline int64_t slab_hap_bytes(int L, int Lr) { return align_up(2 * (int64_t)L + Lr, 16) + 4 * (int64_t)align_up(L, 4); }
__host__ __device__ inline int64_t slab_qm_bytes(int Lq, int Lr) { return align_up(Lr, 16) + 4 * (int64_t)align_up(Lr, 4) + 8 * ((int64_t)Lq + Lr + 1); }

__host__ __device__ inline SlabLayout make_slab(const ScPlan &p, bool with_scalar_aln) {
    SlabLayout s;
    int64_t o = 0;
    for (int h = 0; h < 4; h++) { s.hap[h] = o; o = align_up(o + slab_hap_bytes(p.len[h], p.lr), 16); }
    for (int k = 0; k < 2; k++) { s.qm[k] = o; o = align_up(o + slab_qm_bytes(p.len[k], p.lr), 16); }
    s.rseq_unused = o;
    for (int ai = 0; ai < 4; ai++) {
        s.aln[ai] = o;
        if (with_scalar_aln)
            o = align_up(o + make_layout<int64_t, 4, false>(p.len[ai >> 1] + p.lr, p.len[2 + (ai & 1)], p.lr).total, 16);
    }
    s.total = o;
    return s;
}

struct SlabHap {
    u8 *str, *flg, *ins; int *ptr;
    __device__ SlabHap(u8 *base, int L, int Lr) {
        str = base; flg = base + L; ins = base + 2 * (int64_t)L;
        ptr = (int *)(base + align_up(2 * (int64_t)L + Lr, 16));
    }
};
struct SlabQm {
    u8 *rflg; int *rptr, *toQ, *toR;
    __device__ SlabQm(u8 *base, int Lq, int Lr) {
        rflg = base;
        rptr = (int *)(base + align_up(Lr, 16));
        toQ = rptr + align_up(Lr, 4);
        toR = toQ + ((int64_t)Lq + Lr + 1);
    }
};

// bytes of HBM slab each list entry needs (then exclusive-summed)
__global__ void slab_size_kernel(const ScPlan *plan, const int *list, int n, int64_t *bytes, int scalar_cls) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const ScPlan p = plan[list[i]];
    bytes[i] = make_slab(p, p.cls == scalar_cls).total;
}

// one thread per (entry, hap): expansion into the slab; query haps also build the swap tables
__global__ void slab_setup_kernel(BatchDev in, ScPlan *plan, const int *list, int i0, int i1,
                                  const int64_t *offs, u8 *slab, int *hap_ok) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = i0 + (g >> 2), h = g & 3;
    if (i >= i1) return;
    const int sc = list[i];
    const ScPlan p = plan[sc];
    if (p.cls == CLS_WAVE) return;                       // long path: long_setup_kernel (vd_setup.cuh)
    const SlabLayout S = make_slab(p, p.cls == CLS_SCALAR);
    u8 *base = slab + (offs[i] - offs[i0]);
    SlabHap H(base + S.hap[h], p.len[h], p.lr);
    int len;
    bool ok;
    if (h < 2) {
        SlabQm M(base + S.qm[h], p.len[h], p.lr);
        len = expand_hap<int>(in, sc, h, H.str, H.flg, H.ptr, M.rptr, M.rflg, H.ins, p.len[h]);
        ok = len == p.len[h];
        if (ok) ok = build_swsrc<int>(H.ptr, H.flg, len, M.toR, p.lr) &&
                     build_swsrc<int>(M.rptr, M.rflg, p.lr, M.toQ, len);
    } else {
        len = expand_hap<int>(in, sc, h, H.str, H.flg, H.ptr, nullptr, nullptr, H.ins, p.len[h]);
        ok = len == p.len[h];
    }
    hap_ok[4 * (int64_t)(i - i0) + h] = ok ? 1 : 0;
}

// one thread per (entry, alignment): everything in the HBM slab
__global__ void slab_align_kernel(BatchDev in, OutDev out, const ScPlan *plan, const int *list,
                                  int i0, int i1, const int64_t *offs, u8 *slab, const int *hap_ok,
                                  int only_cls) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = i0 + (g >> 2), ai = g & 3;
    if (i >= i1) return;
    const int sc = list[i];
    const ScPlan p = plan[sc];
    if (p.cls != only_cls) return;
    const int *okp = hap_ok + 4 * (int64_t)(i - i0);
    if (!(okp[0] && okp[1] && okp[2] && okp[3])) {
        out.status[4 * (int64_t)sc + ai] = ST_BAD;
        out.aln_score[4 * (int64_t)sc + ai] = -1;
        return;
    }
    const SlabLayout S = make_slab(p, true);
    u8 *base = slab + (offs[i] - offs[i0]);
    const int qh = ai >> 1, th = 2 + (ai & 1);
    SlabHap HQ(base + S.hap[qh], p.len[qh], p.lr), HT(base + S.hap[th], p.len[th], p.lr);
    SlabQm M(base + S.qm[qh], p.len[qh], p.lr);
    Hap<int> q{p.len[qh], HQ.str, HQ.flg, HQ.ptr, HQ.ins};
    Hap<int> t{p.len[th], HT.str, HT.flg, HT.ptr, HT.ins};
    QMaps<int> qm{M.rptr, M.rflg, M.toQ, M.toR};
    const u8 *rseq = in.rplane_seq + in.ref_off[sc];
    GMem mem{base + S.aln[ai]};
    const AlnLayout<int64_t> L = make_layout<int64_t, 4, false>(q.len + p.lr, t.len, p.lr);
    u32 status = 0;
    int score, end_plane;
    forward_scalar<GMem, 4, int>(mem, L, q, qm, t, rseq, p.lr, score, end_plane);
    const int beg_plane = backward_scalar<GMem, 4, int>(mem, L, q, qm, t, rseq, p.lr, end_plane, status);
    PFScalar<GMem> pfr{&mem, L.oPF, q.len + p.lr, q.len};
    walk_credit<GMem, 4, int>(mem, L, pfr, q, qm, t, rseq, p.lr, beg_plane, end_plane, in, out, sc, ai, status);
    out.aln_score[4 * (int64_t)sc + ai] = score;
    out.aln_end_plane[4 * (int64_t)sc + ai] = (u8)end_plane;
    out.aln_beg_plane[4 * (int64_t)sc + ai] = (u8)beg_plane;
    out.status[4 * (int64_t)sc + ai] = status;
}

}  // namespace vd
