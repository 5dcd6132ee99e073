// simt_emu.h - a small SIMT emulator for the CPU.  TEST INFRASTRUCTURE ONLY.
//
// Compiles the .cu / .cuh sources of vcfdist_b200/csrc unchanged with g++ (-DVD_EMU -include simt_emu.h)
// into tests/simt/_build/libvcfdist_emu.so, so that kernel LOGIC (indexing, warp collectives, barriers,
// the scheduler of vd_api.cu) can be debugged against the oracle in this GPU-less container before a
// GPU box is spent on it.  Nothing under vcfdist_b200/ loads that library: the product path is the
// sm_100a build, and it fails loudly without a GPU.  The emulator says nothing about performance,
// memory-model races or anything else that needs real hardware - the -m gpu tests do.
//
// Model: one fiber per CUDA thread, blocks run one after the other, fibers of a block are scheduled
// round-robin and switch only inside warp collectives (__shfl*_sync, __ballot_sync, __syncwarp, ...) and
// __syncthreads(), which are two-phase rendezvous over the lanes named in the mask.  Exited threads
// count as arrived.  A pass of the scheduler in which no fiber makes progress is reported as a deadlock
// (divergent collective).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <functional>
#include <numeric>
#include <type_traits>
#include <vector>
#include <sys/mman.h>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __grid_constant__
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct short4 { short x, y, z, w; };
struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };
struct int2 { int x, y; };
struct int4 { int x, y, z, w; };
static inline short4 make_short4(short a, short b, short c, short d) { return short4{a, b, c, d}; }
static inline uint2 make_uint2(unsigned a, unsigned b) { return uint2{a, b}; }
static inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { return uint4{a, b, c, d}; }
static inline int2 make_int2(int a, int b) { return int2{a, b}; }
static inline int4 make_int4(int a, int b, int c, int d) { return int4{a, b, c, d}; }

template <class A, class B> constexpr typename std::common_type<A, B>::type min(A a, B b) {
    typedef typename std::common_type<A, B>::type C;
    return (C)b < (C)a ? (C)b : (C)a;
}
template <class A, class B> constexpr typename std::common_type<A, B>::type max(A a, B b) {
    typedef typename std::common_type<A, B>::type C;
    return (C)a < (C)b ? (C)b : (C)a;
}

namespace simt {

struct Fiber {
    void *sp = nullptr;
    dim3 tid;
    int lin = 0;
    bool done = false;
};

struct WarpState {
    uint64_t val[32];
    unsigned arrived = 0, left = 0, exited = 0, released = 0;
};

struct Ctx {
    std::vector<Fiber> fib;
    std::vector<WarpState> warp;
    void *sched_sp = nullptr;
    Fiber *cur = nullptr;
    dim3 blockIdx, blockDim, gridDim;
    std::function<void()> body;
    unsigned char *stacks = nullptr;
    size_t stack_bytes = 512 * 1024;
    int max_fibers = 1024;
    unsigned char *dyn = nullptr;
    size_t dyn_cap = 256 * 1024;
    // block barrier
    int bar_arrived = 0, bar_gen = 0, alive = 0;
    int bar_or[3] = {0, 0, 0};
    unsigned long progress = 0;
    unsigned long launches = 0;
};
inline Ctx &ctx() { static Ctx c; return c; }

extern "C" void simt_switch(void **save_sp, void *load_sp);
#ifdef SIMT_EMU_IMPL
asm(".text\n.weak simt_switch\n.type simt_switch,@function\nsimt_switch:\n"
    "pushq %rbp\npushq %rbx\npushq %r12\npushq %r13\npushq %r14\npushq %r15\n"
    "movq %rsp,(%rdi)\nmovq %rsi,%rsp\n"
    "popq %r15\npopq %r14\npopq %r13\npopq %r12\npopq %rbx\npopq %rbp\nret\n");
#endif

inline void yield() {
    Ctx &c = ctx();
    simt_switch(&c.cur->sp, c.sched_sp);
}

inline void fiber_exit_bookkeeping(Ctx &c, Fiber *f) {
    f->done = true;
    c.progress++;
    WarpState &w = c.warp[f->lin >> 5];
    w.exited |= 1u << (f->lin & 31);
    c.alive--;
    if (c.alive > 0 && c.bar_arrived >= c.alive && c.bar_arrived > 0) { c.bar_arrived = 0; c.bar_gen++; }
}

inline void fiber_main() {
    Ctx &c = ctx();
    c.body();
    fiber_exit_bookkeeping(c, c.cur);
    simt_switch(&c.cur->sp, c.sched_sp);
    abort();
}

inline void launch(dim3 grid, dim3 block, size_t smem, std::function<void()> body) {
    Ctx &c = ctx();
    const int nt = (int)(block.x * block.y * block.z);
    if (nt > c.max_fibers || nt <= 0) { fprintf(stderr, "simt: bad block size %d\n", nt); abort(); }
    if (smem > c.dyn_cap) { fprintf(stderr, "simt: %zu bytes of dynamic shared memory\n", smem); abort(); }
    if (!c.stacks) {
        c.stacks = (unsigned char *)mmap(nullptr, c.stack_bytes * c.max_fibers, PROT_READ | PROT_WRITE,
                                         MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        c.dyn = (unsigned char *)aligned_alloc(128, c.dyn_cap);
    }
    c.launches++;
    c.body = std::move(body);
    c.blockDim = block; c.gridDim = grid;
    c.fib.resize(nt);
    c.warp.resize((nt + 31) / 32);
    for (unsigned bz = 0; bz < grid.z; bz++) for (unsigned by = 0; by < grid.y; by++) for (unsigned bx = 0; bx < grid.x; bx++) {
        c.blockIdx = dim3(bx, by, bz);
        memset(c.dyn, 0xCD, smem);                      // CUDA does not zero shared memory either
        for (auto &w : c.warp) w = WarpState();
        for (int t = 0; t < nt; t++) {
            Fiber &f = c.fib[t];
            f.lin = t; f.done = false;
            f.tid = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
            // initial frame: six callee-saved registers, then the entry address as the return address
            uintptr_t top = (uintptr_t)(c.stacks + c.stack_bytes * (size_t)(t + 1));
            top &= ~(uintptr_t)15;
            void **sp = (void **)top;
            *--sp = nullptr;                                // alignment slot: rsp % 16 == 8 at fiber_main's entry
            *--sp = (void *)(void (*)())fiber_main;
            for (int k = 0; k < 6; k++) *--sp = nullptr;
            f.sp = sp;
        }
        if (nt & 31) c.warp.back().exited = ~0u << (nt & 31);   // lanes that do not exist
        c.alive = nt; c.bar_arrived = 0;
        int remaining = nt;
        while (remaining > 0) {
            const unsigned long p0 = c.progress;
            remaining = 0;
            for (int t = 0; t < nt; t++) {
                Fiber &f = c.fib[t];
                if (f.done) continue;
                c.cur = &f;
                simt_switch(&c.sched_sp, f.sp);
                if (!f.done) remaining++;
            }
            if (remaining > 0 && c.progress == p0) {
                fprintf(stderr, "simt: deadlock in block (%u,%u,%u): %d threads wait at a collective that cannot complete\n",
                        bx, by, bz, remaining);
                abort();
            }
        }
    }
    c.cur = nullptr;
}

// two-phase rendezvous of the lanes in `mask`; returns the values of all lanes in out[32]
inline unsigned collect(unsigned mask, uint64_t v, uint64_t *out) {
    Ctx &c = ctx();
    Fiber *f = c.cur;
    WarpState &w = c.warp[f->lin >> 5];
    const unsigned bit = 1u << (f->lin & 31);
    while (w.arrived & bit) yield();                        // my slot of the previous collective is still being read
    w.val[f->lin & 31] = v;
    w.arrived |= bit;
    c.progress++;
    while (((w.arrived | w.exited) & mask) != mask) yield();
    memcpy(out, w.val, sizeof(w.val));
    const unsigned participants = w.arrived & mask;         // lanes that took part (exited lanes did not)
    w.left |= bit;
    if (((w.left | w.exited) & mask) == mask) {
        w.arrived &= ~mask; w.left &= ~mask;
        w.released |= mask & ~bit & ~w.exited;
        c.progress++;
    } else {
        while (!(w.released & bit)) yield();
        w.released &= ~bit;
        c.progress++;
    }
    return participants;
}

template <class T> inline uint64_t pack(T v) { uint64_t u = 0; static_assert(sizeof(T) <= 8, "shuffle type"); memcpy(&u, &v, sizeof(T)); return u; }
template <class T> inline T unpack(uint64_t u) { T v; memcpy(&v, &u, sizeof(T)); return v; }
inline int lane_id() { return ctx().cur->lin & 31; }

}  // namespace simt

#define threadIdx (simt::ctx().cur->tid)
#define blockIdx (simt::ctx().blockIdx)
#define blockDim (simt::ctx().blockDim)
#define gridDim (simt::ctx().gridDim)

template <class T> inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
    uint64_t a[32]; simt::collect(mask, simt::pack(v), a);
    (void)width;
    return simt::unpack<T>(a[src & 31]);
}
template <class T> inline T __shfl_up_sync(unsigned mask, T v, unsigned d, int width = 32) {
    uint64_t a[32]; simt::collect(mask, simt::pack(v), a);
    (void)width;
    const int l = simt::lane_id();
    return l >= (int)d ? simt::unpack<T>(a[l - d]) : v;
}
template <class T> inline T __shfl_down_sync(unsigned mask, T v, unsigned d, int width = 32) {
    uint64_t a[32]; simt::collect(mask, simt::pack(v), a);
    (void)width;
    const int l = simt::lane_id();
    return l + (int)d < 32 ? simt::unpack<T>(a[l + d]) : v;
}
template <class T> inline T __shfl_xor_sync(unsigned mask, T v, int x, int width = 32) {
    uint64_t a[32]; simt::collect(mask, simt::pack(v), a);
    (void)width;
    return simt::unpack<T>(a[(simt::lane_id() ^ x) & 31]);
}
inline unsigned simt_active(unsigned mask) {      // lanes of the mask that still exist
    return mask & ~simt::ctx().warp[simt::ctx().cur->lin >> 5].exited;
}
inline unsigned __ballot_sync(unsigned mask, int pred) {
    uint64_t a[32]; const unsigned act = simt::collect(mask, (uint64_t)(pred != 0), a);
    unsigned r = 0;
    for (int i = 0; i < 32; i++) if (((act >> i) & 1) && a[i]) r |= 1u << i;
    return r;
}
inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
inline int __all_sync(unsigned mask, int pred) {
    uint64_t a[32]; const unsigned act = simt::collect(mask, (uint64_t)(pred != 0), a);
    for (int i = 0; i < 32; i++) if (((act >> i) & 1) && !a[i]) return 0;
    return 1;
}
inline int __reduce_max_sync(unsigned mask, int v) {
    uint64_t a[32]; const unsigned act = simt::collect(mask, simt::pack(v), a);
    int r = INT32_MIN;
    for (int i = 0; i < 32; i++) if ((act >> i) & 1) r = std::max(r, simt::unpack<int>(a[i]));
    return r;
}
inline int __reduce_min_sync(unsigned mask, int v) {
    uint64_t a[32]; const unsigned act = simt::collect(mask, simt::pack(v), a);
    int r = INT32_MAX;
    for (int i = 0; i < 32; i++) if ((act >> i) & 1) r = std::min(r, simt::unpack<int>(a[i]));
    return r;
}
inline unsigned __reduce_or_sync(unsigned mask, unsigned v) {
    uint64_t a[32]; const unsigned act = simt::collect(mask, simt::pack(v), a);
    unsigned r = 0;
    for (int i = 0; i < 32; i++) if ((act >> i) & 1) r |= simt::unpack<unsigned>(a[i]);
    return r;
}
inline unsigned __reduce_add_sync(unsigned mask, unsigned v) {
    uint64_t a[32]; const unsigned act = simt::collect(mask, simt::pack(v), a);
    unsigned r = 0;
    for (int i = 0; i < 32; i++) if ((act >> i) & 1) r += simt::unpack<unsigned>(a[i]);
    return r;
}
template <class T> inline unsigned __match_any_sync(unsigned mask, T v) {
    uint64_t a[32]; const unsigned act = simt::collect(mask, simt::pack(v), a);
    unsigned r = 0;
    const uint64_t mine = simt::pack(v);
    for (int i = 0; i < 32; i++) if (((act >> i) & 1) && a[i] == mine) r |= 1u << i;
    return r;
}
inline void __syncwarp(unsigned mask = 0xffffffffu) { uint64_t a[32]; simt::collect(mask, 0, a); }
inline unsigned __activemask() { return 0xffffffffu; }
inline void __syncthreads() {
    simt::Ctx &c = simt::ctx();
    c.progress++;
    const int g = c.bar_gen;
    c.bar_arrived++;
    if (c.bar_arrived >= c.alive) { c.bar_arrived = 0; c.bar_gen++; }
    else while (c.bar_gen == g) simt::yield();
}
inline int __syncthreads_or(int pred) {
    // a block-wide flag between two barriers.  Every thread enters with the same barrier generation; the flag of generation
    // g mod 3 cannot be set again before everybody has cleared it (that takes three more barriers).
    simt::Ctx &c = simt::ctx();
    const int idx = c.bar_gen % 3;
    if (pred) c.bar_or[idx] = 1;
    __syncthreads();
    const int r = c.bar_or[idx];
    __syncthreads();
    c.bar_or[idx] = 0;
    return r;
}
inline void __threadfence() {}
inline void __threadfence_block() {}
inline void __threadfence_system() {}
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
inline int __clzll(long long v) { return v ? __builtin_clzll((unsigned long long)v) : 64; }
inline unsigned __brev(unsigned v) { unsigned r = 0; for (int i = 0; i < 32; i++) if ((v >> i) & 1) r |= 1u << (31 - i); return r; }
inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned s) { const uint64_t v = ((uint64_t)hi << 32) | lo; return (unsigned)(v >> (s & 31)); }

template <class T, class U> inline T atomicAdd(T *p, U v) { T o = *p; *p = (T)(o + (T)v); return o; }
template <class T, class U> inline T atomicMin(T *p, U v) { T o = *p; if ((T)v < o) *p = (T)v; return o; }
template <class T, class U> inline T atomicMax(T *p, U v) { T o = *p; if ((T)v > o) *p = (T)v; return o; }
template <class T, class U> inline T atomicOr(T *p, U v) { T o = *p; *p = (T)(o | (T)v); return o; }
template <class T, class U> inline T atomicAnd(T *p, U v) { T o = *p; *p = (T)(o & (T)v); return o; }
template <class T, class U> inline T atomicExch(T *p, U v) { T o = *p; *p = (T)v; return o; }
template <class T, class U, class V> inline T atomicCAS(T *p, U cmp, V v) { T o = *p; if (o == (T)cmp) *p = (T)v; return o; }

// ---------------------------------------------------------------------------------------------
// the slice of the CUDA runtime API that vd_api.cu uses; everything executes synchronously
// ---------------------------------------------------------------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2 };
typedef struct simt_stream_s { int id; } *cudaStream_t;
typedef struct simt_event_s { double t; } *cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaHostAllocMapped = 2, cudaHostAllocDefault = 0 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaFuncAttributePreferredSharedMemoryCarveout = 9 };
struct cudaDeviceProp { int major = 10, minor = 0, multiProcessorCount = 148; char name[64] = "SIMT emulator"; };
inline const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated error"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) { *p = cudaDeviceProp(); return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaMemGetInfo(size_t *fr, size_t *tot) { *fr = (size_t)6 << 30; *tot = (size_t)8 << 30; return cudaSuccess; }
inline cudaError_t cudaMalloc(void **p, size_t n) { *p = aligned_alloc(256, (n + 255) / 256 * 256 + 256); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
template <class T> inline cudaError_t cudaMalloc(T **p, size_t n) { return cudaMalloc((void **)p, n); }
inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
inline cudaError_t cudaHostAlloc(void **p, size_t n, unsigned) { *p = calloc(1, n + 64); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
inline cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t = nullptr) { if (n) memset(p, v, n); return cudaSuccess; }
inline cudaError_t cudaMemset(void *p, int v, size_t n) { if (n) memset(p, v, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { if (n) memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { if (n) memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = new simt_stream_s{0}; return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t *s, unsigned, int) { *s = new simt_stream_s{0}; return cudaSuccess; }
inline cudaError_t cudaDeviceGetStreamPriorityRange(int *lo, int *hi) { *lo = 0; *hi = -1; return cudaSuccess; }
inline cudaError_t cudaStreamCreate(cudaStream_t *s) { *s = new simt_stream_s{0}; return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t s) { delete s; return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new simt_event_s{0}; return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = new simt_event_s{0}; return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr) {
    timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); e->t = ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; return cudaSuccess;
}
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventQuery(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)(b->t - a->t); return cudaSuccess; }
template <class F> inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
inline cudaError_t cudaHostRegister(void *, size_t, unsigned) { return cudaSuccess; }
inline cudaError_t cudaHostUnregister(void *) { return cudaSuccess; }

// the two CUB device primitives the scheduler uses
namespace cub {
struct DeviceRadixSort {
    template <class K, class V>
    static cudaError_t SortPairs(void *tmp, size_t &bytes, const K *kin, K *kout, const V *vin, V *vout, int n, int = 0, int = 8 * sizeof(K),
                                 cudaStream_t = nullptr) {
        if (!tmp) { bytes = 16; return cudaSuccess; }
        std::vector<int> idx(n);
        std::iota(idx.begin(), idx.end(), 0);
        std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return kin[a] < kin[b]; });
        for (int i = 0; i < n; i++) { kout[i] = kin[idx[i]]; vout[i] = vin[idx[i]]; }
        return cudaSuccess;
    }
};
struct DeviceScan {
    template <class In, class Out>
    static cudaError_t ExclusiveSum(void *tmp, size_t &bytes, const In *in, Out *out, int n, cudaStream_t = nullptr) {
        if (!tmp) { bytes = 16; return cudaSuccess; }
        Out acc = 0;
        for (int i = 0; i < n; i++) { const Out v = (Out)in[i]; out[i] = acc; acc += v; }
        return cudaSuccess;
    }
};
}  // namespace cub

#define VD_LAUNCH(kernel, grid, block, smem, stream, ...) \
    simt::launch(dim3(grid), dim3(block), (size_t)(smem), [&]() { kernel(__VA_ARGS__); })
#define VD_DYN_SHARED(name) unsigned char *name = simt::ctx().dyn
