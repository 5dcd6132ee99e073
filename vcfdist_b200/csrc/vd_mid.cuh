// Fused mid-size kernel: superclusters too big for tiny_kernel whose four alignments still fit
// in one SM's shared memory (matrix sides up to 256 rows).  One 128-thread block per
// supercluster, one warp per alignment: haplotype expansion, the wavefront forward and backward
// sweeps (wave_*_body with a 32-lane group), walk and credit all run on shared memory — no flag
// matrix in HBM, no slab, no separate walk launch.  HBM traffic is the compact batch in and the
// result records out, as for tiny_kernel.
#pragma once
#include "vd_wave.cuh"

namespace vd {

template <int K>
__global__ void __launch_bounds__(MID_TPB) mid_kernel(BatchDev in, OutDev out, const ScPlan *__restrict__ plan,
                                                     const int *__restrict__ items) {
    extern __shared__ __align__(16) u8 smem[];
    __shared__ int sEnd[4][2];
    __shared__ int sOk[4];
    const int sc = items[blockIdx.x];
    const ScPlan p = plan[sc];
    const MidLayout M = mid_layout(p, K);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int Lr = p.lr;
    auto hstr = [&](int h) { return smem + M.hap[h]; };
    auto hflg = [&](int h) { return smem + M.hap[h] + p.len[h]; };
    auto hins = [&](int h) { return smem + M.hap[h] + 2 * p.len[h]; };
    auto hptr = [&](int h) { return (short *)(smem + M.hap[h] + a2(2 * p.len[h] + Lr)); };
    auto qrflg = [&](int k) { return smem + M.qm[k]; };
    auto qrptr = [&](int k) { return (short *)(smem + M.qm[k] + a2(Lr)); };
    auto qtoQ = [&](int k) { return qrptr(k) + Lr; };
    auto qtoR = [&](int k) { return qtoQ(k) + (p.len[k] + Lr + 1); };
    auto wsrcQ = [&](int k) { return (int *)(smem + M.wq[k]); };
    auto wsrcR = [&](int k) { return wsrcQ(k) + p.len[k]; };
    auto wtpb = [&](int k) { return (u8 *)(wsrcR(k) + Lr); };
    u8 *rseq = smem + M.rseq;

    // ---- phase 1: warp h expands haplotype h and builds its tables (lane 0; O(L)) ----
    if (lane == 0) {
        const int h = warp;
        const bool isq = h < 2;
        const int len = expand_hap<short>(in, sc, h, hstr(h), hflg(h), hptr(h), isq ? qrptr(h) : nullptr,
                                          isq ? qrflg(h) : nullptr, hins(h), p.len[h]);
        bool ok = len == p.len[h];
        if (ok && isq) {
            ok = build_swsrc<short>(hptr(h), hflg(h), len, qtoR(h), Lr) &&
                 build_swsrc<short>(qrptr(h), qrflg(h), Lr, qtoQ(h), len);
            if (ok) {
                build_srcinfo<short>(hptr(h), hflg(h), len, qtoR(h), Lr, qrflg(h), nullptr, wsrcQ(h));
                build_srcinfo<short>(qrptr(h), qrflg(h), Lr, qtoQ(h), len, hflg(h), hptr(h), wsrcR(h));
                const short *pp = hptr(h);
                const u8 *ff = hflg(h);
                u8 *tb = wtpb(h);
                for (int a = 0; a < len; a++)
                    tb[a] = (a > 0 && (((int)pp[a] != (int)pp[a - 1] + 1) || (ff[a] & P_VAR_BEG))) ? 1 : 0;
            }
        }
        if (ok && !isq) {
            u8 *ti = smem + M.wt[h - 2];
            const u8 *ff = hflg(h), *ss = hstr(h);
            for (int c = 0; c < len; c++) {
                const bool tok = c > 0 && (!(ff[c - 1] & P_VARIANT) || (ff[c - 1] & P_VAR_END));
                ti[c] = (u8)((ss[c] & 0x7f) | (tok ? 0x80 : 0));
            }
        }
        sOk[h] = ok ? 1 : 0;
    }
    if (warp == 3) {
        const u8 *rs = in.rplane_seq + in.ref_off[sc];
        for (int k = lane; k < Lr; k += 32) rseq[k] = rs[k];
    }
    __syncthreads();
    const int ai = warp;
    const int64_t oi = 4 * (int64_t)sc + ai;
    if (!(sOk[0] && sOk[1] && sOk[2] && sOk[3])) {
        if (lane == 0) { out.status[oi] = ST_BAD; out.aln_score[oi] = -1; }
        return;
    }

    // ---- phase 2: warp ai runs its alignment ----
    const int qh = ai >> 1, th = 2 + (ai & 1);
    WaveCtxT<short> X;
    X.sc = sc; X.ai = ai; X.Lq = p.len[qh]; X.Lr = Lr; X.Lt = p.len[th];
    X.padQ = (X.Lq + K - 1) / K * K;
    X.NP = mid_np(X.Lq, Lr, K);
    X.qstr = hstr(qh); X.rseq = rseq; X.tinfo = smem + M.wt[ai & 1];
    X.qflg = hflg(qh); X.rflg = qrflg(qh); X.tpb = wtpb(qh);
    X.toQ = qtoQ(qh); X.toR = qtoR(qh); X.srcQ = wsrcQ(qh); X.srcR = wsrcR(qh);
    X.qptr = hptr(qh); X.rptr = qrptr(qh);
    X.F = smem + M.aln[ai];
    u8 *scratch = smem + M.scr[ai];

    int score, end_plane, beg_plane;
    u32 status;
    wave_fwd_body<32, K, short>(X, lane, scratch, sEnd[warp], score, end_plane);
    __syncwarp();
    wave_bwd_body<32, K, short>(X, lane, scratch, end_plane, beg_plane, status);
    status = __reduce_or_sync(0xffffffffu, status);
    __syncwarp();
    if (lane == 0) {
        Hap<short> q{X.Lq, hstr(qh), hflg(qh), hptr(qh), hins(qh)};
        Hap<short> t{X.Lt, hstr(th), hflg(th), hptr(th), hins(th)};
        QMaps<short> qm{qrptr(qh), qrflg(qh), qtoQ(qh), qtoR(qh)};
        SMemIL mem{scratch};
        AlnLayout<int> L;
        const int np = X.Lq + Lr + X.Lt + 4;
        L.oPF = L.oF = L.oD0 = L.oD1 = L.oT0 = L.oT1 = 0;
        L.oPQ = 0; L.oPT = a4(2 * np); L.oPS = 2 * a4(2 * np); L.oLev = L.oPS + a4(np); L.total = 0;
        PFWave pfr{X.F, X.NP, X.padQ};
        walk_credit<SMemIL, 2, short>(mem, L, pfr, q, qm, t, rseq, Lr, beg_plane, end_plane, in, out, sc, ai, status);
        out.aln_score[oi] = score;
        out.aln_end_plane[oi] = (u8)end_plane;
        out.aln_beg_plane[oi] = (u8)beg_plane;
        out.status[oi] = status;
    }
}

inline void mid_configure() {
    cudaFuncSetAttribute(mid_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, MID_SMEM_MAX);
    cudaFuncSetAttribute(mid_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, MID_SMEM_MAX);
    cudaFuncSetAttribute(mid_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, MID_SMEM_MAX);
    cudaFuncSetAttribute(mid_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, MID_SMEM_MAX);
}
inline void mid_launch(cudaStream_t st, int kcls, int n, int smem_bytes, const BatchDev &in, const OutDev &out,
                       const ScPlan *plan, const int *items) {
    if (n <= 0) return;
    switch (kcls) {
        case 0: mid_kernel<1><<<n, MID_TPB, smem_bytes, st>>>(in, out, plan, items); break;
        case 1: mid_kernel<2><<<n, MID_TPB, smem_bytes, st>>>(in, out, plan, items); break;
        case 2: mid_kernel<4><<<n, MID_TPB, smem_bytes, st>>>(in, out, plan, items); break;
        case 3: mid_kernel<8><<<n, MID_TPB, smem_bytes, st>>>(in, out, plan, items); break;
    }
}

}  // namespace vd
