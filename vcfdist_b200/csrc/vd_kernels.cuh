// Kernels of the precision/recall path (sm_100a):
//   plan_kernel        sizes + class of every supercluster, work list of the non-tiny ones
//   tiny_kernel        fused: one thread per alignment, one quad per supercluster, all
//                      matrices in shared memory (the 93 % class of SURVEY.md 8d)
//   slab_size_kernel / slab_setup_kernel / slab_align_kernel
//                      thread-per-alignment path with matrices in an HBM slab
// The wavefront kernels for long superclusters are in vd_wave.cuh.
#pragma once
#include "vd_scalar.cuh"
#include "vd_midlayout.cuh"

namespace vd {

// ---- tiny-class limits --------------------------------------------------------------------
constexpr int TINY_TPB = 128;          // threads per block = 32 superclusters
constexpr int TINY_TL = 16;            // max haplotype length
constexpr int TINY_TR = 12;            // max window (REF plane) length
constexpr int TINY_CAP = 320;          // private shared-memory bytes per alignment
constexpr int TINY_STRIDE = TINY_CAP + 4;   // private slice stride: odd number of words -> conflict-free
static_assert(((TINY_STRIDE / 4) & 1) == 1, "private slice stride must be an odd number of words");
constexpr int TINY_SW = (TINY_TL + TINY_TR + 1 + 3) & ~3;      // one CSR swap table
constexpr int TINY_SC_BYTES = 4 * (3 * TINY_TL + TINY_TR) + 2 * (2 * TINY_TR + 2 * TINY_SW) + TINY_TR + 8;   // per-supercluster shared area

constexpr u32 ST_BAD = 0x0800u;        // VD_ST_ERR_BADINPUT

struct PlanCounters {
    int n_list;                        // superclusters not handled by tiny_kernel
    int n_bad;
    unsigned long long cells;
    unsigned long long cells_list;
    unsigned status_or;
    int n_mid[N_MCLS];                 // superclusters of the fused mid-size kernel, per (K, smem bin) class
    int mid_cursor[N_MCLS];
};

// OR of all status words (so that the host only scans them when an error bit is set)
__global__ void status_or_kernel(const u32 *status, int64_t n, unsigned *dst) {
    unsigned v = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) v |= status[i];
    v = __reduce_or_sync(0xffffffffu, v);
    if ((threadIdx.x & 31) == 0 && v) atomicOr(dst, v);
}

__device__ __forceinline__ int tiny_need(int Lq, int Lr, int Lt) {
    return make_layout<int, 2, true>(Lq + Lr, Lt, Lr).total;
}

// One thread per supercluster.
__global__ void plan_kernel(BatchDev in, ScPlan *plan, int *list, int *mlist, PlanCounters *cnt, int force_class, int big_class) {
    const int sc0 = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = sc0 < in.n_sc;
    const int sc = live ? sc0 : in.n_sc - 1;      // dead lanes recompute the last one and discard it
    ScPlan p;
    const int lr = (int)(in.ref_off[sc + 1] - in.ref_off[sc]);
    p.lr = lr;
    bool bad = lr < 1;
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
        if (len < 1) bad = true;
    }
    unsigned long long cells = 0;
    int cls = CLS_SCALAR;
    if (bad) cls = CLS_BAD;
    else {
        bool tiny = lr <= TINY_TR;
        for (int h = 0; h < 4; h++) tiny = tiny && p.len[h] <= TINY_TL;
        for (int ai = 0; ai < 4; ai++) {
            const int lq = p.len[ai >> 1], lt = p.len[2 + (ai & 1)];
            cells += (unsigned long long)(lq + lr) * lt;
            if (tiny) tiny = tiny_need(lq, lr, lt) <= TINY_CAP;
        }
        if (tiny && force_class < 0) cls = CLS_TINY;
        else {
            const int kc = mid_kclass(p);
            const int need = kc >= 0 ? mid_layout(p, 1 << kc).total : (1 << 30);
            if (force_class == CLS_MID && need <= MID_SMEM_MAX) {      // measured slower than the slab path: opt-in only
                const int mc = kc * N_MBIN + mid_bin(need);
                cls = CLS_MID | (mc << 8);              // the class id rides in the upper bits
                if (live) atomicAdd(&cnt->n_mid[mc], 1);
            } else cls = (force_class > CLS_TINY && force_class != CLS_MID) ? force_class : big_class;
        }
    }
    p.cls = cls;
    if (live) plan[sc] = p;
    // warp-aggregated counters: one atomic per warp and counter instead of one per thread
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    (void)mlist;
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
        if (c_all) atomicAdd(&cnt->cells, c_all);
        if (m_list) { base = atomicAdd(&cnt->n_list, __popc(m_list)); atomicAdd(&cnt->cells_list, c_list); }
        if (m_bad) atomicAdd(&cnt->n_bad, __popc(m_bad));
    }
    base = __shfl_sync(full, base, 0);
    if (is_list) list[base + __popc(m_list & ((1u << lane) - 1))] = sc;
}

// class-sorted work list of the fused mid-size kernel
struct MidBase { int b[N_MCLS]; };
__global__ void mid_fill_kernel(const ScPlan *plan, int n_sc, PlanCounters *cnt, MidBase mb, int *mlist) {
    const int sc = blockIdx.x * blockDim.x + threadIdx.x;
    if (sc >= n_sc) return;
    const int cls = plan[sc].cls;
    if ((cls & 0xff) != CLS_MID) return;
    const int mc = cls >> 8;
    mlist[mb.b[mc] + atomicAdd(&cnt->mid_cursor[mc], 1)] = sc;
}

// ---- fused tiny kernel --------------------------------------------------------------------
// per-supercluster shared area (bytes): 4 x hap {str TL, flg TL, ptr TL, ins TR},
// 2 x qmaps {rptr TR, rflg TR, toQ SW, toR SW}, rseq TR, hlen 4 x int16
struct TinyArea {
    u8 *base;
    __device__ u8 *str(int h) const { return base + h * (3 * TINY_TL + TINY_TR); }
    __device__ u8 *flg(int h) const { return str(h) + TINY_TL; }
    __device__ int8_t *ptr(int h) const { return (int8_t *)(str(h) + 2 * TINY_TL); }
    __device__ u8 *ins(int h) const { return str(h) + 3 * TINY_TL; }
    __device__ u8 *qm(int qh) const { return base + 4 * (3 * TINY_TL + TINY_TR) + qh * (2 * TINY_TR + 2 * TINY_SW); }
    __device__ int8_t *rptr(int qh) const { return (int8_t *)qm(qh); }
    __device__ u8 *rflg(int qh) const { return qm(qh) + TINY_TR; }
    __device__ int8_t *toQ(int qh) const { return (int8_t *)(qm(qh) + 2 * TINY_TR); }
    __device__ int8_t *toR(int qh) const { return (int8_t *)(qm(qh) + 2 * TINY_TR + TINY_SW); }
    __device__ u8 *rseq() const { return base + 4 * (3 * TINY_TL + TINY_TR) + 2 * (2 * TINY_TR + 2 * TINY_SW); }
    __device__ short *hlen() const { return (short *)(rseq() + TINY_TR); }
};
static_assert((TINY_TR % 2) == 0 && (TINY_SC_BYTES % 4) == 0, "alignment of the tiny shared area");

__global__ void __launch_bounds__(TINY_TPB)
tiny_kernel(BatchDev in, OutDev out, const ScPlan *__restrict__ plan) {
    extern __shared__ __align__(16) u8 smem[];
    constexpr int SPB = TINY_TPB / 4;
    const int tid = threadIdx.x, quad = tid >> 2, h = tid & 3;
    const int sc = blockIdx.x * SPB + quad;
    const bool active = sc < in.n_sc && plan[sc].cls == CLS_TINY;
    TinyArea A{smem + quad * TINY_SC_BYTES};
    SMemIL mem{smem + SPB * TINY_SC_BYTES + tid * TINY_STRIDE};

    int lr = 0;
    if (active) {
        lr = plan[sc].lr;
        // thread h expands haplotype h (generate_ptrs_strs); query haps also keep the ref side
        const bool isq = h < 2;
        const int len = expand_hap<int8_t>(in, sc, h, A.str(h), A.flg(h), A.ptr(h),
                                           isq ? A.rptr(h) : nullptr, isq ? A.rflg(h) : nullptr,
                                           A.ins(h), TINY_TL);
        bool ok = len == plan[sc].len[h];
        if (ok && isq) {
            ok = build_swsrc<int8_t>(A.ptr(h), A.flg(h), len, A.toR(h), lr) &&
                 build_swsrc<int8_t>(A.rptr(h), A.rflg(h), lr, A.toQ(h), len);
        }
        A.hlen()[h] = (short)(ok ? len : -1);
        if (h == 2) {
            const u8 *rs = in.rplane_seq + in.ref_off[sc];
            for (int k = 0; k < lr; k++) A.rseq()[k] = rs[k];
        }
    }
    __syncwarp();
    if (!active) return;
    const int ai = h;
    const int qh = ai >> 1, th = 2 + (ai & 1);
    u32 status = 0;
    const short *hl = A.hlen();
    if (hl[0] < 0 || hl[1] < 0 || hl[2] < 0 || hl[3] < 0) {
        out.status[4 * (int64_t)sc + ai] = ST_BAD;
        out.aln_score[4 * (int64_t)sc + ai] = -1;
        return;
    }
    Hap<int8_t> q{hl[qh], A.str(qh), A.flg(qh), A.ptr(qh), A.ins(qh)};
    Hap<int8_t> t{hl[th], A.str(th), A.flg(th), A.ptr(th), A.ins(th)};
    QMaps<int8_t> qm{A.rptr(qh), A.rflg(qh), A.toQ(qh), A.toR(qh)};
    const AlnLayout<int> L = make_layout<int, 2, true>(q.len + lr, t.len, lr);

    int score, end_plane;
    forward_scalar<SMemIL, 2, int8_t>(mem, L, q, qm, t, A.rseq(), lr, score, end_plane);
    const int beg_plane = backward_scalar<SMemIL, 2, int8_t>(mem, L, q, qm, t, A.rseq(), lr, end_plane, status);
    PFScalar<SMemIL> pfr{&mem, L.oPF, q.len + lr, q.len};
    walk_credit<SMemIL, 2, int8_t>(mem, L, pfr, q, qm, t, A.rseq(), lr, beg_plane, end_plane,
                                   in, out, sc, ai, status);
    out.aln_score[4 * (int64_t)sc + ai] = score;
    out.aln_end_plane[4 * (int64_t)sc + ai] = (u8)end_plane;
    out.aln_beg_plane[4 * (int64_t)sc + ai] = (u8)beg_plane;
    out.status[4 * (int64_t)sc + ai] = status;
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
