// Wavefront kernels for long superclusters (block per alignment, column sweep with a
// min-plus prefix scan, flag matrices spilled to HBM).  -- placeholder until written --
#pragma once
#include "vd_kernels.cuh"
struct DevBuf;
namespace vd {
constexpr int kBigClass = CLS_SCALAR;
inline void wave_configure() {}
inline int wave_run(cudaStream_t, cudaEvent_t *, const BatchDev &, const OutDev &, const ScPlan *, const int *,
                    int, int, const int64_t *, const int64_t *, u8 *, const int *, DevBuf *, int, vd_stats *,
                    float *, float *, float *) { return VD_OK; }
}  // namespace vd
