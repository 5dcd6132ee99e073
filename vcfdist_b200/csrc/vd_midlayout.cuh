// Shared-memory layout of the fused mid-size kernel (vd_mid.cuh); also used by plan_kernel to
// decide which superclusters fit.
#pragma once
#include "vd_common.cuh"

namespace vd {

constexpr int MID_TPB = 128;
constexpr int MID_SMEM_MAX = 56 * 1024;      // 4 resident blocks per SM
constexpr int N_MK = 4;                      // rows per lane K = 1, 2, 4, 8
constexpr int N_MBIN = 4;                    // shared-memory bins (occupancy follows the bin, not the worst case)
constexpr int N_MCLS = N_MK * N_MBIN;
__host__ __device__ inline int mid_bin_cap(int b) { const int v[N_MBIN] = {6 * 1024, 12 * 1024, 24 * 1024, MID_SMEM_MAX}; return v[b]; }
__host__ __device__ inline int mid_bin(int need) { for (int b = 0; b < N_MBIN; b++) if (need <= mid_bin_cap(b)) return b; return -1; }

struct MidLayout {
    int hap[4], qm[2], wq[2], wt[2], rseq, aln[4], scr[4];
    int total;
};

__host__ __device__ inline int mid_np(int Lq, int Lr, int K) { return (Lq + K - 1) / K * K + (Lr + K - 1) / K * K; }

// smallest K in {1,2,4,8} whose 32*K rows hold every alignment of the supercluster; -1 if none
__host__ __device__ inline int mid_kclass(const ScPlan &p) {
    for (int c = 0; c < N_MK; c++) {
        const int K = 1 << c;
        bool ok = true;
        for (int q = 0; q < 2; q++) ok = ok && mid_np(p.len[q], p.lr, K) <= 32 * K;
        if (ok) return c;
    }
    return -1;
}

__host__ __device__ inline int a2(int x) { return (x + 1) & ~1; }
__host__ __device__ inline int a4(int x) { return (x + 3) & ~3; }
__host__ __device__ inline int a16(int x) { return (x + 15) & ~15; }

__host__ __device__ inline MidLayout mid_layout(const ScPlan &p, int K) {
    MidLayout m;
    int o = 0;
    const int Lr = p.lr;
    for (int h = 0; h < 4; h++) { m.hap[h] = o; o = a4(o + a2(2 * p.len[h] + Lr) + 2 * p.len[h]); }
    for (int k = 0; k < 2; k++) { m.qm[k] = o; o = a4(o + a2(Lr) + 2 * Lr + 4 * (p.len[k] + Lr + 1)); }
    for (int k = 0; k < 2; k++) { m.wq[k] = o; o = a4(o + 4 * (p.len[k] + Lr) + p.len[k]); }
    for (int k = 0; k < 2; k++) { m.wt[k] = o; o = a4(o + p.len[2 + k]); }
    m.rseq = o; o = a16(o + Lr);
    for (int ai = 0; ai < 4; ai++) {
        const int Lq = p.len[ai >> 1], Lt = p.len[2 + (ai & 1)];
        const int NP = mid_np(Lq, Lr, K);
        m.aln[ai] = o;
        o = a16(o + NP * Lt);
        m.scr[ai] = o;
        const int np = Lq + Lr + Lt + 4, mn = (Lr < Lt ? Lr : Lt) + 1;
        int need = 6 * 32 * K;                                         // backward sweep buffers (forward needs 4*32K)
        const int walk = 2 * a4(2 * np) + a4(np) + a4(2 * mn);          // path q/t (int16), flags, Levenshtein row
        if (walk > need) need = walk;
        o = a16(o + need);
    }
    m.total = o;
    return m;
}

}  // namespace vd
