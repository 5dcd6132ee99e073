// Shared definitions of the sm_100a precision/recall kernels.
#pragma once
#include <cstdint>
#ifdef VD_EMU
#define VD_NOINLINE __attribute__((noinline))
// kernel-logic debugging on the CPU: tests/simt/simt_emu.h (force-included by tests/simt/Makefile)
// provides the CUDA subset these sources use; never part of the product build
#else
#include <cuda_runtime.h>
#define VD_NOINLINE __noinline__
#define VD_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define VD_DYN_SHARED(name) extern __shared__ __align__(16) unsigned char name[]
#endif

// ---- TMA bulk copies (cp.async.bulk, the 1-D form of the tensor memory accelerator) + mbarrier ----------
// Contiguous, 16-byte aligned ranges of HBM are pulled into shared memory by the copy engine: one lane
// arms an mbarrier with the byte count and issues the copy, the consumers wait on the barrier's phase.
#ifdef VD_EMU
#define VD_MBAR_INIT(bar, count) ((void)0)
#define VD_MBAR_INIT_FENCE() ((void)0)
#define VD_BULK_G2S(dst, src, bytes, bar) memcpy((void *)(dst), (const void *)(src), (size_t)(bytes))
#define VD_MBAR_WAIT(bar, parity) __syncwarp()
#else
namespace vd {
__device__ __forceinline__ unsigned smem_addr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, void *bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(void *bar, unsigned parity) {
    asm volatile("{\n.reg .pred P1;\nVD_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra VD_DONE;\nbra VD_WAIT;\nVD_DONE:\n}"
                 ::"r"(smem_addr(bar)), "r"(parity) : "memory");
}
}  // namespace vd
#define VD_MBAR_INIT(bar, count) vd::mbar_init((bar), (count))
#define VD_MBAR_INIT_FENCE() vd::mbar_init_fence()
#define VD_BULK_G2S(dst, src, bytes, bar) vd::bulk_g2s((dst), (src), (bytes), (bar))
#define VD_MBAR_WAIT(bar, parity) vd::mbar_wait((bar), (parity))
#endif

#include "vcfdist_b200.h"

namespace vd {

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;

// move flags of the backward pass / walk (path_ptrs), src/defs.h:110-120
constexpr int PTR_INS = 1, PTR_DEL = 2, PTR_MAT = 4, PTR_SUB = 8, PTR_SWP = 16;
// forward flag byte (aln_ptrs).  MAT and SUB share one bit: which one it is follows from
// comparing the two bases again.  The freed bits hold the chosen swap source.
constexpr int F_INS = 1, F_DEL = 2, F_DIAG = 4, F_SWP = 8;
constexpr int F_TIE = 16;      // several sources had the winning score (reference order-dependent)
constexpr int F_K_SHIFT = 5;   // bits 5-7: index of the chosen source in the row's source list
// pointer flags, src/defs.h:122-129
constexpr int P_VARIANT = 1, P_VAR_BEG = 2, P_VAR_END = 4, P_INS_LOC = 8;

constexpr int INF = 1 << 29;
constexpr int NEG = -(1 << 28);

// device view of vd_batch_in (all pointers in HBM)
struct BatchDev {
    int n_sc;
    const int64_t *ref_off;
    const u8 *ref_seq;
    const u8 *rplane_seq;     // never null on the device: aliases ref_seq when absent
    const int64_t *var_off;
    const int32_t *var_pos;
    const int32_t *var_rlen;
    const u8 *var_type;
    const int64_t *alt_off;
    const u8 *alt_seq;
    const float *var_qual;
    float max_qual;
    int64_t n_var;
};

struct OutDev {
    int32_t *aln_score;
    u8 *aln_end_plane;
    u8 *aln_beg_plane;
    u32 *status;
    u8 *assigned;
    int32_t *sync_group;
    int32_t *ref_ed;
    int32_t *query_ed;
    float *callq;
};

// classes decided by the plan kernel
// CLS_TINY: one of the fused shared-memory kernels (small_kernel<K>); bits 8.. of ScPlan::cls then
// hold its (class, cost bin) slot
enum : int { CLS_TINY = 0, CLS_WAVE = 1, CLS_SCALAR = 2, CLS_BAD = 3 };

// per-supercluster plan record
struct ScPlan {
    int32_t lr;          // window / REF-plane length
    int32_t len[4];      // haplotype string lengths q1,q2,t1,t2
    int32_t cls;
    int32_t hom;         // query haplotypes identical and truth haplotypes identical: one alignment, records replicated
};

static __host__ __device__ inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

}  // namespace vd
