// C-ABI of the precision/recall hot path (include/vcfdist_b200.h): handle management,
// host<->device staging, the batch scheduler and the kernel launches.
//
// Scheduler (the B200 counterpart of precision_recall_threads_wrapper's thread/RAM ladder,
// src/dist.cpp:1656-1727): one plan kernel classifies every supercluster; the fused
// shared-memory kernel takes the short ones in a single launch over the whole batch; the
// rest are expanded into an HBM slab and run through the wavefront (or scalar) kernels in
// chunks bounded by the handle's scratch budget.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <ctime>
#include <string>
#include <chrono>
#include <vector>

#ifndef VD_EMU
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#endif

#include "vd_kernels.cuh"
#include "vd_wave.cuh"
#include "vd_band.cuh"
#include "vd_setup.cuh"
#include "vd_reach.cuh"

using namespace vd;

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = n + n / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

// Per-chunk state of the scheduler: plan, order array, counters and the events of one chunk.  Two sets,
// so that the plan pass of chunk i+1 runs while the kernels of chunk i are still executing.
struct Work {
    DevBuf plan, list, order, ranks, ranks_out, iota, counters, cubtmp;
    DevBuf wsc_slab, wsc_hdr, wsc_work;     // split warp path: one slot per supercluster of the wsc groups, its header, the walk kernel's work list
    PlanCounters *h_counters = nullptr;     // pinned + mapped: written by publish_kernel, never by a copy engine
    cudaEvent_t ev[4] = {};                 // plan start, plan end, short kernels issued, chunk end
    cudaEvent_t ev5 = nullptr;              // fork point of the short-kernel launch groups
    cudaEvent_t evL = nullptr;              // everything issued on the main stream for this chunk
    cudaEvent_t gev[vd::N_GROUP][2] = {};   // start / end of each short-kernel launch group
    cudaEvent_t wev[4] = {};                // split warp path: expansion start / end, walk start / end
    bool wsc_used = false;
    bool grp_used[vd::N_GROUP] = {};
    bool busy = false;                      // launched, not yet harvested
    BatchDev in; OutDev out;
    float ms_fwd = 0, ms_bwd = 0, ms_walk = 0;
    int n_bad = 0;
};

struct vd_handle {
    static constexpr int NST = 3;           // chunks in flight in vd_run (staging + work sets)
    Work work[NST];
    cudaStream_t s_plan = nullptr;          // plan pass of the next chunk
    cudaStream_t s_epi = nullptr;           // join of a chunk's launch groups + status reduction (the main stream moves on)
    cudaStream_t s_wsc = nullptr;           // expansion kernel of the warp path (beside the main stream's small kernels)
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[8] = {};
    cudaStream_t side[vd::N_WCLS] = {};     // one stream per wavefront class: classes run concurrently
    cudaEvent_t sev[vd::N_WCLS][4] = {};
    cudaEvent_t gev[vd::N_GROUP][2] = {};   // start / end of each short-kernel launch group
    int64_t scratch_budget = 0;
    int num_sms = 148;
    std::string err;
    vd_stats stats = {};
    unsigned stats_status_or = 0;   // OR of every status word of the last call
    bool banded_fwd = true;         // VD_DENSE_FWD=1 skips the banded forward sweep (testing)
    bool banded_bwd = true;         // VD_SPARSE_BWD=1: frontier kernel instead of the banded backward sweep (testing)
    bool sparse_bwd = true;         // VD_DENSE_BWD=1 selects the dense backward sweep (testing)
    int sbwd_min_class = 4;         // VD_SBWD_MIN_CLASS: wave classes below it use the dense backward sweep
    int force_class = -1;           // VD_FORCE_CLASS env: testing hook (1 wave, 2 scalar slab)
    int small_lo = 0, small_hi = vd::N_SMALL - 1;   // VD_SMALL_MIN / VD_SMALL_MAX: small-kernel classes in use (testing)
    int serial = 0;                 // VD_SERIAL=1: every launch group on the main stream (clean per-kernel event times)
    int use_hom = 1;                // VD_HOM=0: homozygous superclusters run all four alignments (testing)
    int use_band = 1;               // VD_BAND=0: no banded warp kernels, every long alignment goes to the dense block kernels (testing)
    int use_wsc = 1;                // VD_WSC=0: mid-size superclusters go to the HBM-slab path instead of the warp kernel
    int wf_block_min = 384;         // VD_WF_BLOCK_MIN: wavefront width from which a clustering / --distance problem gets a block, not a warp
    cudaEvent_t wf_ev[4] = {};
    int64_t wf_cluster_min = 6144;  // VD_WF_CLUSTER_MIN: ... and from which it gets a cluster of blocks
    int walk_wpw = 0;               // VD_WALK_WPW: alignments per warp in the long path's walk kernels (0 = by their number)
    int wsc_split = 1;              // VD_WSC_SPLIT=0: the fused warp kernels instead of expansion / sweeps / walk as separate launches
    // staged input / output (vd_run)
    struct Stage {                  // one of the staging sets of the host-buffer pipeline
        DevBuf in_ref_off, in_ref_seq, in_rplane, in_var_off, in_var_pos, in_var_rlen, in_var_type,
               in_alt_off, in_alt_seq, in_var_qual;
        DevBuf o_score, o_endp, o_begp, o_status, o_assigned, o_sg, o_red, o_qed, o_callq;
        DevBuf p_score, p_planes, p_status, p_sg, p_red, p_qed;       // 16-bit records (vd_run_packed)
        DevBuf c_ref_len, c_hap_nvar, c_pos16, c_rlen16, c_alen16, c_tmp;   // compact input (vd_run_compact)
        cudaEvent_t in_done = nullptr, out_done = nullptr;
        bool out_pending = false;
    } stage[NST];
    cudaStream_t s_in = nullptr, s_out = nullptr;
    int ramp = 1;                   // VD_RAMP=0: uniform chunks
    int64_t chunk_sc = 786432;       // superclusters per pipeline chunk (VD_CHUNK_SC; a multiple of VD_COMPACT_BLOCK)
    // work
    DevBuf need_dense, bytes, offs, cubtmp, slab, hap_ok, wave_desc, band_state, band_lb, dense_bytes, dense_off, dense;
    DevBuf wf_in, wf_scratch;       // vd_wf_batch: staged problems, wavefront rings
    unsigned *h_range = nullptr;            // pinned + mapped: set by pack_out_kernel when a value does not fit 16 bits
    WaveItems *h_witems = nullptr;          // pinned + mapped (a small D2H memcpy would queue behind the bulk result copies of vd_run)
};

static int fail(vd_handle *h, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (h) h->err = buf;
    return code;
}

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
    return fail(h, e_ == cudaErrorMemoryAllocation ? VD_E_NOMEM : VD_E_CUDA, "%s: %s (%s:%d)", #call, \
                cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)

extern "C" int vd_abi_version(void) { return VD_ABI_VERSION; }

extern "C" int vd_create(int device, int64_t scratch_bytes, vd_handle **out) {
    if (!out) return VD_E_BADINPUT;
    *out = nullptr;
    int n = 0;
    const bool times = getenv("VD_CREATE_TIMES") != nullptr;                 // where start-up goes: context, streams and events, kernel attributes
    const auto t_0 = std::chrono::steady_clock::now();
    auto since = [&](std::chrono::steady_clock::time_point t) { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t).count(); };
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) return VD_E_NODEVICE;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return VD_E_NODEVICE;
    if (prop.major != 10) return VD_E_NODEVICE;          // built for sm_100a only
    if (cudaSetDevice(device) != cudaSuccess) return VD_E_NODEVICE;
    cudaFree(nullptr);                                                       // the context comes up here
    const double ms_ctx = since(t_0);
    vd_handle *h = new vd_handle();
    h->device = device;
    h->num_sms = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { delete h; return VD_E_CUDA; }
    bool cok = true;
    // VD_PRIO=1: higher stream priority for the side streams (warp sweeps ahead of the walk kernel, the long path's rungs)
    // and the epilogue stream.  Measured on the WGS step: 8.44 ms with, 8.35 ms without - the machine is full either way -
    // so it is off by default.
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    const int prio_mode = getenv("VD_PRIO") ? atoi(getenv("VD_PRIO")) : 0;      // 2: only the side streams of the wide rungs (few, long alignments)
    const int prio_top = prio_hi;
    if (prio_mode != 1) prio_hi = prio_lo;
    for (auto &e : h->ev) cok &= cudaEventCreate(&e) == cudaSuccess;
    for (auto &e : h->wf_ev) cok &= cudaEventCreateWithFlags(&e, cudaEventDisableTiming) == cudaSuccess;
    for (int c = 0; c < N_WCLS; c++) {
        cok &= cudaStreamCreateWithPriority(&h->side[c], cudaStreamNonBlocking, (prio_mode == 2 && c >= 2 && c < N_RUNG) ? prio_top : prio_hi) == cudaSuccess;
        for (auto &e : h->sev[c]) cok &= cudaEventCreate(&e) == cudaSuccess;
    }
    for (auto &w : h->work) {
        for (auto &e : w.ev) cok &= cudaEventCreate(&e) == cudaSuccess;
        cok &= cudaEventCreate(&w.ev5) == cudaSuccess;
        cok &= cudaEventCreate(&w.evL) == cudaSuccess;
        for (auto &g : w.gev) for (auto &e : g) cok &= cudaEventCreate(&e) == cudaSuccess;
        for (auto &e : w.wev) cok &= cudaEventCreate(&e) == cudaSuccess;
        cok &= cudaHostAlloc((void **)&w.h_counters, sizeof(PlanCounters), cudaHostAllocMapped) == cudaSuccess;
    }
    cok &= cudaStreamCreateWithFlags(&h->s_plan, cudaStreamNonBlocking) == cudaSuccess;
    cok &= cudaStreamCreateWithPriority(&h->s_epi, cudaStreamNonBlocking, prio_hi) == cudaSuccess;
    cok &= cudaStreamCreateWithPriority(&h->s_wsc, cudaStreamNonBlocking, prio_hi) == cudaSuccess;
    cok &= cudaHostAlloc((void **)&h->h_witems, sizeof(WaveItems), cudaHostAllocMapped) == cudaSuccess;
    cok &= cudaHostAlloc((void **)&h->h_range, 64, cudaHostAllocMapped) == cudaSuccess;
    if (scratch_bytes <= 0) {
        size_t fr = 0, tot = 0;
        cudaMemGetInfo(&fr, &tot);
        scratch_bytes = (int64_t)(fr * 0.6);
    }
    h->scratch_budget = scratch_bytes;
    if (const char *fc = getenv("VD_FORCE_CLASS")) h->force_class = atoi(fc);
    if (const char *v = getenv("VD_SMALL_MIN")) h->small_lo = atoi(v);
    if (const char *v = getenv("VD_SMALL_MAX")) h->small_hi = atoi(v);
    if (const char *v = getenv("VD_WSC")) h->use_wsc = atoi(v);
    if (const char *v = getenv("VD_BAND")) h->use_band = atoi(v);
    if (const char *v = getenv("VD_WSC_SPLIT")) h->wsc_split = atoi(v);
    if (const char *v = getenv("VD_WF_BLOCK_MIN")) h->wf_block_min = std::max(1, atoi(v));
    if (const char *v = getenv("VD_WF_CLUSTER_MIN")) h->wf_cluster_min = std::max(1, atoi(v));
#ifdef VD_EMU
    h->wf_cluster_min = INT64_MAX;                                            // no thread-block clusters under the emulator
#endif
    if (const char *v = getenv("VD_WALK_WPW")) { const int w = atoi(v); h->walk_wpw = (w == 1 || w == 2 || w == 4 || w == 8 || w == 16 || w == 32) ? w : 0; }
    if (const char *v = getenv("VD_SERIAL")) h->serial = atoi(v);
    if (const char *v = getenv("VD_HOM")) h->use_hom = atoi(v);
    if (const char *v = getenv("VD_RAMP")) h->ramp = atoi(v);
    if (const char *df = getenv("VD_DENSE_FWD")) h->banded_fwd = atoi(df) == 0;
    if (const char *db = getenv("VD_DENSE_BWD")) h->sparse_bwd = atoi(db) == 0;
    if (const char *v = getenv("VD_SPARSE_BWD")) h->banded_bwd = atoi(v) == 0;
    if (const char *sm = getenv("VD_SBWD_MIN_CLASS")) h->sbwd_min_class = atoi(sm);
    if (const char *cs = getenv("VD_CHUNK_SC")) h->chunk_sc = atoll(cs) > 0 ? atoll(cs) : h->chunk_sc;
    cok &= cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking) == cudaSuccess;
    cok &= cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking) == cudaSuccess;
    for (auto &sg : h->stage) {
        cok &= cudaEventCreateWithFlags(&sg.in_done, cudaEventDisableTiming) == cudaSuccess;
        cok &= cudaEventCreateWithFlags(&sg.out_done, cudaEventDisableTiming) == cudaSuccess;
    }
    if (!cok) { vd_destroy(h); return VD_E_CUDA; }          // a stream, event or pinned block could not be created
    const double ms_obj = since(t_0) - ms_ctx;
    small_configure();
    wsc_configure();
    wsc_split_configure();
    wave_configure();
    band_configure();
    if (times) fprintf(stderr, "vd_create: context %.1f ms, streams / events / pinned blocks %.1f ms, kernel attributes %.1f ms\n",
                       ms_ctx, ms_obj, since(t_0) - ms_ctx - ms_obj);
    *out = h;
    return VD_OK;
}

extern "C" void vd_destroy(vd_handle *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    for (auto &sg : h->stage) {
        DevBuf *sb[] = {&sg.in_ref_off, &sg.in_ref_seq, &sg.in_rplane, &sg.in_var_off, &sg.in_var_pos, &sg.in_var_rlen,
                        &sg.in_var_type, &sg.in_alt_off, &sg.in_alt_seq, &sg.in_var_qual, &sg.o_score, &sg.o_endp,
                        &sg.o_begp, &sg.o_status, &sg.o_assigned, &sg.o_sg, &sg.o_red, &sg.o_qed, &sg.o_callq,
                        &sg.p_score, &sg.p_planes, &sg.p_status, &sg.p_sg, &sg.p_red, &sg.p_qed,
                        &sg.c_ref_len, &sg.c_hap_nvar, &sg.c_pos16, &sg.c_rlen16, &sg.c_alen16, &sg.c_tmp};
        for (DevBuf *b : sb) b->release();
        if (sg.in_done) cudaEventDestroy(sg.in_done);
        if (sg.out_done) cudaEventDestroy(sg.out_done);
    }
    if (h->s_in) cudaStreamDestroy(h->s_in);
    if (h->s_out) cudaStreamDestroy(h->s_out);
    DevBuf *bufs[] = {&h->need_dense, &h->bytes, &h->offs, &h->cubtmp, &h->slab, &h->hap_ok, &h->wave_desc, &h->band_state,
                      &h->band_lb, &h->dense_bytes, &h->dense_off, &h->dense, &h->wf_in, &h->wf_scratch};
    for (DevBuf *b : bufs) b->release();
    for (auto &w : h->work) {
        DevBuf *wb[] = {&w.plan, &w.list, &w.order, &w.ranks, &w.ranks_out, &w.iota, &w.counters, &w.cubtmp, &w.wsc_slab, &w.wsc_hdr, &w.wsc_work};
        for (DevBuf *b : wb) b->release();
        if (w.h_counters) cudaFreeHost(w.h_counters);
        for (auto &e : w.ev) if (e) cudaEventDestroy(e);
        if (w.ev5) cudaEventDestroy(w.ev5);
        if (w.evL) cudaEventDestroy(w.evL);
        for (auto &g : w.gev) for (auto &e : g) if (e) cudaEventDestroy(e);
        for (auto &e : w.wev) if (e) cudaEventDestroy(e);
    }
    if (h->s_plan) cudaStreamDestroy(h->s_plan);
    if (h->s_epi) cudaStreamDestroy(h->s_epi);
    if (h->s_wsc) cudaStreamDestroy(h->s_wsc);
    if (h->h_witems) cudaFreeHost(h->h_witems);
    if (h->h_range) cudaFreeHost(h->h_range);
    for (auto &e : h->ev) if (e) cudaEventDestroy(e);
    for (auto &e : h->wf_ev) if (e) cudaEventDestroy(e);
    for (int c = 0; c < N_WCLS; c++) {
        for (auto &e : h->sev[c]) if (e) cudaEventDestroy(e);
        if (h->side[c]) cudaStreamDestroy(h->side[c]);
    }
    cudaStreamDestroy(h->stream);
    delete h;
}

extern "C" const char *vd_last_error(const vd_handle *h) { return h ? h->err.c_str() : "null handle"; }
extern "C" void *vd_stream(const vd_handle *h) { return h ? (void *)h->stream : nullptr; }
extern "C" int vd_get_stats(const vd_handle *h, vd_stats *out) {
    if (!h || !out) return VD_E_BADINPUT;
    *out = h->stats;
    return VD_OK;
}

// The whole path on device-resident buffers, in three steps per chunk so that chunks overlap:
//   chunk_plan     (stream sp)   memsets of the result buffers, plan pass, order sort, counters to the host
//   chunk_exec     (stream st)   reads the counters, issues every launch group (+ the long path, which still
//                                synchronises internally) and records the chunk's end event; no host sync
//   chunk_harvest                waits for the end event, folds event times and counters into the stats
// `out` may carry pointers shifted by the chunk's first variant (see vd_run); `base` has the true
// buffer starts for the memsets.  Stats accumulate across the chunks of one call.
static int chunk_plan(vd_handle *h, Work &W, cudaStream_t sp, const BatchDev &in, const OutDev &out, const OutDev &base) {
    const int n_sc = in.n_sc;
    const int64_t n_var = in.n_var;
    W.in = in; W.out = out;
    W.busy = true;
    W.ms_fwd = W.ms_bwd = W.ms_walk = 0;
    for (auto &u : W.grp_used) u = false;
    CK(cudaEventRecord(W.ev[0], sp));
    CK(cudaMemsetAsync(base.assigned, 0, 2 * n_var, sp));
    CK(cudaMemsetAsync(base.sync_group, 0, 8 * n_var, sp));
    CK(cudaMemsetAsync(base.ref_ed, 0, 8 * n_var, sp));
    CK(cudaMemsetAsync(base.query_ed, 0, 8 * n_var, sp));
    CK(cudaMemsetAsync(base.callq, 0, 8 * n_var, sp));
    CK(cudaMemsetAsync(base.status, 0, 16 * (size_t)n_sc, sp));

    CK(W.plan.ensure(sizeof(ScPlan) * (size_t)n_sc));
    CK(W.list.ensure(sizeof(int) * (size_t)n_sc));
    CK(W.order.ensure(sizeof(int) * (size_t)n_sc));
    CK(W.counters.ensure(sizeof(PlanCounters)));
    CK(W.ranks.ensure((size_t)n_sc)); CK(W.ranks_out.ensure((size_t)n_sc)); CK(W.iota.ensure(4 * (size_t)n_sc));
    CK(cudaMemsetAsync(W.counters.p, 0, sizeof(PlanCounters), sp));
    PlanCounters *dcnt = (PlanCounters *)W.counters.p;
    int *order = (int *)W.order.p;                   // class- and cost-sorted short superclusters
    VD_LAUNCH(plan_kernel, (n_sc + 255) / 256, 256, 0, sp, in, out, (ScPlan *)W.plan.p, (int *)W.list.p, (u8 *)W.ranks.p, (int *)W.iota.p, dcnt,
                                                    h->force_class, kBigClass, h->small_lo, h->small_hi, h->use_wsc, h->use_hom);
    VD_LAUNCH(small_base_kernel, 1, 32, 0, sp, dcnt);
    {   // order[] = supercluster indices stably sorted by rank (non-short superclusters sort to the end)
        size_t tmp = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, tmp, (const u8 *)W.ranks.p, (u8 *)W.ranks_out.p, (const int *)W.iota.p, order,
                                        n_sc, 0, 8, sp);
        CK(W.cubtmp.ensure(tmp));
        cub::DeviceRadixSort::SortPairs(W.cubtmp.p, tmp, (const u8 *)W.ranks.p, (u8 *)W.ranks_out.p, (const int *)W.iota.p, order,
                                        n_sc, 0, 8, sp);
    }
    h->stats.n_launches += 6;
    VD_LAUNCH(publish_kernel, 1, 64, 0, sp, (const u32 *)dcnt, (u32 *)W.h_counters, (int)(sizeof(PlanCounters) / 4));
    CK(cudaEventRecord(W.ev[1], sp));
    return VD_OK;
}

static int chunk_exec(vd_handle *h, Work &W) {
    cudaStream_t st = h->stream;
    const BatchDev &in = W.in;
    const OutDev &out = W.out;
    const int n_sc = in.n_sc;
    vd_stats &S = h->stats;
    const bool trace = getenv("VD_TRACE") != nullptr;
    ScPlan *plan = (ScPlan *)W.plan.p;
    int *list = (int *)W.list.p;
    int *order = (int *)W.order.p;
    CK(cudaEventSynchronize(W.ev[1]));       // counters are now on the host
    CK(cudaGetLastError());
    const PlanCounters pc = *W.h_counters;
    W.n_bad = pc.n_bad;
    CK(cudaStreamWaitEvent(st, W.ev[1], 0));

    // ---- short superclusters: one fused launch per group, most expensive group first: the warp
    //      kernel's (slots, shared-memory bin) groups on the side streams (they are latency-bound and
    //      overlap with each other and with everything else), the thread-per-alignment classes on the
    //      main stream ----
    int n_small = 0;
    // split warp path: slab slots (one per supercluster of a wsc group, of the group's bin size) and headers; the
    // expansion of all groups is one launch, ahead of everything else on the main stream
    int64_t slab_off[N_GROUP] = {}, hdr_off[N_GROUP] = {};
    WscGroups WG{};
    W.wsc_used = false;
    if (h->wsc_split) {
        int64_t so = 0, ho = 0;
        for (int g = 2 * N_SMALL; g < N_GROUP; g++) {
            const int w = (g >> 1) - N_SMALL, cap = wsc_bin_cap(N_WBIN - 1 - w % N_WBIN);
            slab_off[g] = so; hdr_off[g] = ho;
            if (pc.grp_count[g] > 0) {
                const int k = WG.n++;
                WG.first[k] = (int)ho; WG.order_first[k] = pc.grp_first[g]; WG.stride[k] = cap; WG.hom[k] = g & 1; WG.base[k] = so;
            }
            so += (int64_t)pc.grp_count[g] * cap;
            ho += pc.grp_count[g];
        }
        WG.first[WG.n] = (int)ho;
        if (ho > 0) {
            CK(W.wsc_slab.ensure((size_t)so + 256)); CK(W.wsc_hdr.ensure(sizeof(WscHdr) * (size_t)ho + 16));
            CK(W.wsc_work.ensure(4 * (size_t)(4 * ho + 4)));       // [0]: count, [4..]: items
            W.wsc_used = true;
            cudaStream_t sx = h->serial ? st : h->s_wsc;           // nothing on the main stream depends on the expansion
            if (sx != st) CK(cudaStreamWaitEvent(sx, W.ev[1], 0));
            CK(cudaMemsetAsync(W.wsc_work.p, 0, 16, sx));
            CK(cudaEventRecord(W.wev[0], sx));
            wsc_expand_launch(sx, in, plan, order, WG, (u8 *)W.wsc_slab.p, (WscHdr *)W.wsc_hdr.p);
            CK(cudaEventRecord(W.wev[1], sx));
            S.n_launches++;
        }
    }
    CK(cudaEventRecord(W.ev5, st));
    for (int g = N_GROUP - 1; g >= 0; g--) {
        const int cnt = pc.grp_count[g];
        if (cnt <= 0) continue;
        const int g0 = g >> 1;                               // group without the homozygous bit
        const bool hom = g & 1;
        cudaStream_t gs = (g0 < N_SMALL || h->serial) ? st : h->side[(g - 2 * N_SMALL) % N_WCLS];
        if (gs != st) CK(cudaStreamWaitEvent(gs, (g0 >= N_SMALL && W.wsc_used) ? W.wev[1] : W.ev5, 0));
        CK(cudaEventRecord(W.gev[g][0], gs));
        if (g0 < N_SMALL) small_launch(gs, g0, hom, in, out, plan, order + pc.grp_first[g], cnt);
        else {
            const int w = g0 - N_SMALL;                      // (slots - 1) * N_WBIN + (N_WBIN - 1 - bin)
            if (h->wsc_split) {
                wsc_split_launch(gs, w / N_WBIN + 1, N_WBIN - 1 - w % N_WBIN, hom, in, out, plan, order + pc.grp_first[g], cnt,
                                 (u8 *)W.wsc_slab.p + slab_off[g], (WscHdr *)W.wsc_hdr.p + hdr_off[g],
                                 WscWork{(int *)W.wsc_work.p, (int *)W.wsc_work.p + 4, (int)hdr_off[g]});
            } else {
                wsc_launch(gs, w / N_WBIN + 1, N_WBIN - 1 - w % N_WBIN, hom, in, out, plan, order + pc.grp_first[g], cnt);
            }
        }
        CK(cudaEventRecord(W.gev[g][1], gs));
        W.grp_used[g] = true;
        S.n_launches++;
        n_small += cnt;
        const int k = g0 < N_SMALL ? g0 : N_SMALL;
        S.n_small[k] += cnt;
        S.io_small[k] += (int64_t)pc.io_grp[g];
        if (hom) S.n_hom += cnt;
    }
    CK(cudaEventRecord(W.ev[2], st));
    S.cells += (int64_t)pc.cells;
    S.n_long += 4 * (int64_t)pc.n_list;
    S.n_short += 4 * (int64_t)n_small;

    // ---- the rest: HBM slab, wavefront / scalar kernels ----
    float ms_fwd = 0, ms_bwd = 0, ms_walk = 0;
    if (pc.n_list > 0) {
        const int n = pc.n_list;
        CK(h->bytes.ensure(8 * (size_t)(n + 1)));
        CK(h->offs.ensure(8 * (size_t)(n + 1)));
        int64_t *bytes = (int64_t *)h->bytes.p, *offs = (int64_t *)h->offs.p;
        CK(cudaMemsetAsync(bytes + n, 0, 8, st));
        VD_LAUNCH(wave_size_kernel, (n + 255) / 256, 256, 0, st, plan, list, n, bytes);
        S.n_launches++;
        size_t tmp = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tmp, bytes, offs, n + 1, st);
        CK(h->cubtmp.ensure(tmp));
        cub::DeviceScan::ExclusiveSum(h->cubtmp.p, tmp, bytes, offs, n + 1, st);
        S.n_launches++;
        std::vector<int64_t> hoffs((size_t)n + 1);
        CK(cudaMemcpyAsync(hoffs.data(), offs, 8 * (size_t)(n + 1), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        // chunks of consecutive list entries whose slabs fit the scratch budget
        int i0 = 0;
        while (i0 < n) {
            int i1 = i0 + 1;
            while (i1 < n && hoffs[i1 + 1] - hoffs[i0] <= h->scratch_budget) i1++;
            const int64_t need = hoffs[i1] - hoffs[i0];
            if (need > h->scratch_budget && i1 == i0 + 1 && need > (int64_t)150e9)
                return fail(h, VD_E_TOOLARGE, "supercluster needs %lld bytes of scratch", (long long)need);
            CK(h->slab.ensure((size_t)need));
            const int m = i1 - i0;
            CK(h->hap_ok.ensure(16 * (size_t)m));
            // expansion + tables: a warp per haplotype for the long path, a thread per haplotype for the scalar-slab class
            VD_LAUNCH(long_setup_kernel, m, 128, 0, st, in, (const ScPlan *)plan, (const int *)list, i0, i1, (const int64_t *)offs, (u8 *)h->slab.p,
                      (int *)h->hap_ok.p);
            S.n_launches++;
            if (h->force_class == CLS_SCALAR) {
                VD_LAUNCH(slab_setup_kernel, (4 * m + 127) / 128, 128, 0, st, in, plan, list, i0, i1, offs, (u8 *)h->slab.p,
                          (int *)h->hap_ok.p);
                VD_LAUNCH(slab_align_kernel, (4 * m + 63) / 64, 64, 0, st, in, out, plan, list, i0, i1, offs, (u8 *)h->slab.p,
                          (const int *)h->hap_ok.p, CLS_SCALAR);
                S.n_launches += 2;
            }
            // ---- long path: class-sorted items; banded warp kernels first, dense block kernels for the rest ----
            CK(h->wave_desc.ensure(sizeof(WaveItems) + 4 * (size_t)(4 * m + 4)));
            WaveItems *wi = (WaveItems *)h->wave_desc.p;
            int *items = (int *)((u8 *)h->wave_desc.p + sizeof(WaveItems));
            CK(cudaMemsetAsync(wi, 0, sizeof(WaveItems), st));
            VD_LAUNCH(wave_count_kernel, (4 * m + 127) / 128, 128, 0, st, plan, list, i0, i1, (const int *)h->hap_ok.p, wi);
            S.n_launches += 2;
            VD_LAUNCH(publish_kernel, 1, 64, 0, st, (const u32 *)wi, (u32 *)h->h_witems, (int)(sizeof(WaveItems) / 4));
            CK(cudaStreamSynchronize(st));
            WaveItems hwi = *h->h_witems;
            ClsBase cb;
            int total_items = 0;
            for (int c = 0; c < N_WLIST; c++) { cb.b[c] = total_items; total_items += hwi.count[c]; }
            CK(h->band_state.ensure(4 * (size_t)total_items + 16));
            CK(h->band_lb.ensure(8 * (size_t)total_items + 16));
            int *bstate = (int *)h->band_state.p, *blb = (int *)h->band_lb.p;
            VD_LAUNCH(wave_fill_kernel, (4 * m + 127) / 128, 128, 0, st, in, plan, list, i0, i1, (const int *)h->hap_ok.p,
                      wi, cb, items, out, bstate, blb, blb + total_items, h->use_band);
            S.n_launches++;
            if (total_items > 0) {
                CK(h->need_dense.ensure(4 * (size_t)total_items + 16));
                CK(h->dense_bytes.ensure(8 * (size_t)(total_items + 1)));
                CK(h->dense_off.ensure(8 * (size_t)(total_items + 1)));
                WaveArgs WA{in, out, plan, list, i0, offs, (u8 *)h->slab.p, items, bstate, nullptr, nullptr};
                // alignments per warp in the walk kernels (VD_WALK_WPW): one - measured on the SV workload, 41 k long alignments
                // per step, 2 .. 32 per warp change nothing: the walks' time is the longest walk, not their number
                const int wpw = h->walk_wpw > 0 ? h->walk_wpw : 1;
                CK(cudaEventRecord(h->ev[4], st));
                // ---- banded warp kernels (vd_band.cuh): rung K = 4, 8, 16; the backward sweep and the walk of a rung
                //      run beside the forward sweep of the next one ----
                if (h->use_band) {
                    // round 1: rung 0 takes every item; what it cannot solve leaves with a guess of the rung it needs,
                    // and rungs 1.. then run side by side, each on the items guessed for it.  round 2: rung after rung
                    // over whatever a guess was too low for (normally nothing)
                    int *bhint = blb + total_items;
                    for (int r = 0; r < N_RUNG; r++) {
                        cudaStream_t bs = h->serial ? st : h->side[r];
                        if (bs != st) CK(cudaStreamWaitEvent(bs, r ? h->sev[0][1] : h->ev[4], 0));
                        CK(cudaEventRecord(h->sev[r][0], bs));
                        band_launch_rung(bs, r, WA, total_items, bstate, blb, bhint, true, r > 0, wi);
                        CK(cudaEventRecord(h->sev[r][1], bs));
                        band_launch_rung(bs, r, WA, total_items, bstate, blb, bhint, false, 0, wi);
                        CK(cudaEventRecord(h->sev[r][2], bs));
                        VD_LAUNCH(band_walk_kernel, (total_items + 4 * wpw - 1) / (4 * wpw), 128, 0, bs, WA, total_items, bstate, band_rung_k(r), wpw);
                        CK(cudaEventRecord(h->sev[r][3], bs));
                        S.n_launches += 3;
                    }
                    for (int r = 0; r < N_RUNG; r++) if (!h->serial) CK(cudaStreamWaitEvent(st, h->sev[r][3], 0));
                    for (int r = 2; r < N_RUNG; r++) {
                        band_launch_rung(st, r, WA, total_items, bstate, blb, bhint, true, 0, wi);
                        band_launch_rung(st, r, WA, total_items, bstate, blb, bhint, false, 0, wi);
                        VD_LAUNCH(band_walk_kernel, (total_items + 4 * wpw - 1) / (4 * wpw), 128, 0, st, WA, total_items, bstate, band_rung_k(r), wpw);
                        S.n_launches += 3;
                    }
                }
                // ---- what is left: dense-phase scratch sized on the device, then the block kernels per shape class ----
                int64_t *dbytes = (int64_t *)h->dense_bytes.p, *doff = (int64_t *)h->dense_off.p;
                CK(cudaMemsetAsync(dbytes + total_items, 0, 8, st));
                VD_LAUNCH(wave_dense_size_kernel, (total_items + 255) / 256, 256, 0, st, plan, list, i0, (const int *)items, (const int *)bstate,
                          total_items, dbytes, wi);
                VD_LAUNCH(publish_kernel, 1, 64, 0, st, (const u32 *)wi, (u32 *)h->h_witems, (int)(sizeof(WaveItems) / 4));
                S.n_launches += 2;
                CK(cudaEventRecord(h->ev[5], st));
                CK(cudaStreamSynchronize(st));
                CK(cudaGetLastError());
                hwi = *h->h_witems;
                S.n_dense += hwi.n_dense;
                S.band_cells += (int64_t)hwi.band_cells; S.band_rows += (int64_t)hwi.band_rows; S.band_cols += (int64_t)hwi.band_cols;
                if (h->use_band) {
                    for (int r = 0; r < N_RUNG; r++) {
                        float a_ = 0, b_ = 0, w_ = 0;
                        cudaEventElapsedTime(&a_, h->sev[r][0], h->sev[r][1]);
                        cudaEventElapsedTime(&b_, h->sev[r][1], h->sev[r][2]);
                        cudaEventElapsedTime(&w_, h->sev[r][2], h->sev[r][3]);
                        ms_fwd += a_; ms_bwd += b_; ms_walk += w_;
                        if (trace) fprintf(stderr, "[run_resident] band rung K=%d tau=%d: fwd %.2f ms, bwd %.2f ms, walk %.2f ms\n", band_rung_k(r), band_rung_tau(r), a_, b_, w_);
                    }
                    float e_ = 0;
                    cudaEventElapsedTime(&e_, h->ev[4], h->ev[5]);
                    S.ms_band += e_;
                    if (trace) fprintf(stderr, "[run_resident] %d long alignments, banded phase %.2f ms, %d left to the dense kernels (%.1f MB)\n",
                                       total_items, e_, hwi.n_dense, hwi.dense_bytes / 1e6);
                }
                if (hwi.n_dense > 0) {
                    if ((int64_t)hwi.dense_bytes > h->scratch_budget)
                        return fail(h, VD_E_NOMEM, "dense phase needs %lld bytes of scratch (budget %lld)", (long long)hwi.dense_bytes,
                                    (long long)h->scratch_budget);
                    size_t tmp2 = 0;
                    cub::DeviceScan::ExclusiveSum(nullptr, tmp2, dbytes, doff, total_items + 1, st);
                    CK(h->cubtmp.ensure(tmp2));
                    cub::DeviceScan::ExclusiveSum(h->cubtmp.p, tmp2, dbytes, doff, total_items + 1, st);
                    CK(h->dense.ensure((size_t)hwi.dense_bytes + 256));
                    WA.dense = (u8 *)h->dense.p; WA.dense_off = doff;
                    S.n_launches++;
                    // forward then backward of each class on its own stream (an alignment's backward pass
                    // only depends on its own forward pass), all classes concurrently; then join
                    CK(cudaEventRecord(h->ev[6], st));
                    for (int c = N_WLIST - 1; c >= 0; c--) {            // biggest shapes first: they are the critical path
                        if (!hwi.count[c]) continue;
                        cudaStream_t ss = h->serial ? st : h->side[c % N_WCLS];
                        if (ss != st) CK(cudaStreamWaitEvent(ss, h->ev[6], 0));
                        CK(cudaEventRecord(h->sev[c % N_WCLS][0], ss));
                        if (c == N_WCLS) {                              // no block kernel takes the shape: thread per alignment
                            VD_LAUNCH(wave_oversize_kernel, (hwi.count[c] + 31) / 32, 32, 0, ss, WA, cb.b[c], hwi.count[c]);
                            CK(cudaEventRecord(h->sev[c % N_WCLS][3], ss));
                            if (ss != st) CK(cudaStreamWaitEvent(st, h->sev[c % N_WCLS][3], 0));
                            S.n_launches++;
                            continue;
                        }
                        wave_launch(ss, WA, c, cb.b[c], hwi.count[c], true, true, h->banded_fwd ? (int *)h->need_dense.p : nullptr);
                        CK(cudaEventRecord(h->sev[c][1], ss));
                        wave_launch(ss, WA, c, cb.b[c], hwi.count[c], false, h->sparse_bwd && c >= h->sbwd_min_class,
                                    h->banded_fwd ? (int *)h->need_dense.p : nullptr, h->banded_bwd);
                        CK(cudaEventRecord(h->sev[c][2], ss));
                        // walk + credit of this class right behind its backward sweep, on the same stream
                        VD_LAUNCH(wave_walk_kernel, (hwi.count[c] + 4 * wpw - 1) / (4 * wpw), 128, 0, ss, WA, cb.b[c], hwi.count[c], wpw);
                        CK(cudaEventRecord(h->sev[c][3], ss));
                        if (ss != st) CK(cudaStreamWaitEvent(st, h->sev[c][3], 0));
                        S.n_launches += 3;
                    }
                    CK(cudaEventRecord(h->ev[7], st));
                    CK(cudaStreamSynchronize(st));
                    CK(cudaGetLastError());
                    // kernel durations: summed over classes (they overlap in wall time)
                    for (int c = 0; c < N_WCLS; c++) {
                        if (!hwi.count[c]) continue;
                        float a_ = 0, b_ = 0, w2 = 0;
                        cudaEventElapsedTime(&a_, h->sev[c][0], h->sev[c][1]);
                        cudaEventElapsedTime(&b_, h->sev[c][1], h->sev[c][2]);
                        cudaEventElapsedTime(&w2, h->sev[c][2], h->sev[c][3]);
                        ms_fwd += a_; ms_bwd += b_; ms_walk += w2;
                        if (trace) fprintf(stderr, "[run_resident] dense phase, wave class %d: %d alignments in the class, fwd %.2f ms, bwd %.2f ms, walk %.2f ms\n", c, hwi.count[c], a_, b_, w2);
                    }
                    S.spill_bytes += 3 * (int64_t)hwi.spill_cells;
                }
                VD_LAUNCH(wave_hom_replicate_kernel, (m + 127) / 128, 128, 0, st, in, out, plan, list, i0, i1, (const int *)h->hap_ok.p);
                S.n_launches++;
                CK(cudaEventRecord(h->ev[7], st));
                CK(cudaStreamSynchronize(st));
                float w_ = 0;
                cudaEventElapsedTime(&w_, h->ev[4], h->ev[7]);
                S.ms_long_wall += w_;
            }
            i0 = i1;
        }
    }
    // join on the epilogue stream: the main stream is free for the next chunk while this chunk's
    // side-stream groups drain
    cudaStream_t se = h->serial ? st : h->s_epi;
    if (se != st) {
        CK(cudaEventRecord(W.evL, st));
        CK(cudaStreamWaitEvent(se, W.evL, 0));
        for (int g = 2 * N_SMALL; g < N_GROUP; g++) if (W.grp_used[g]) CK(cudaStreamWaitEvent(se, W.gev[g][1], 0));
    }
    if (W.wsc_used) {                          // walk + credit of all warp-path groups, behind their sweeps
        CK(cudaEventRecord(W.wev[2], se));
        wsc_walk_launch(se, in, out, plan, order, WG, (u8 *)W.wsc_slab.p, (const WscHdr *)W.wsc_hdr.p, (const int *)W.wsc_work.p,
                        (const int *)W.wsc_work.p + 4);
        CK(cudaEventRecord(W.wev[3], se));
        S.n_launches++;
    }
    VD_LAUNCH(status_or_kernel, 296, 256, 0, se, out.status, 4 * (int64_t)n_sc, &((PlanCounters *)W.counters.p)->status_or);
    S.n_launches++;
    VD_LAUNCH(publish_kernel, 1, 64, 0, se, (const u32 *)W.counters.p, (u32 *)W.h_counters, (int)(sizeof(PlanCounters) / 4));
    CK(cudaEventRecord(W.ev[3], se));
    W.ms_fwd = ms_fwd; W.ms_bwd = ms_bwd; W.ms_walk = ms_walk;
    return VD_OK;
}

static int chunk_harvest(vd_handle *h, Work &W) {
    if (!W.busy) return VD_OK;
    W.busy = false;
    vd_stats &S = h->stats;
    CK(cudaEventSynchronize(W.ev[3]));
    CK(cudaGetLastError());
    h->stats_status_or |= W.h_counters->status_or;
    float e_ = 0;
    cudaEventElapsedTime(&e_, W.ev[0], W.ev[1]); S.ms_plan += e_;
    cudaEventElapsedTime(&e_, W.ev5, W.ev[2]); S.ms_short += e_;
    cudaEventElapsedTime(&e_, W.ev[0], W.ev[3]); S.ms_total += e_;      // chunk spans overlap in vd_run
    S.ms_long_fwd += W.ms_fwd; S.ms_long_bwd += W.ms_bwd; S.ms_long_walk += W.ms_walk;
    for (int g = 0; g < N_GROUP; g++) {       // per-group durations of the short kernels (the warp kernel's overlap)
        if (!W.grp_used[g]) continue;
        cudaEventElapsedTime(&e_, W.gev[g][0], W.gev[g][1]);
        S.ms_small[(g >> 1) < N_SMALL ? (g >> 1) : N_SMALL] += e_;
    }
    if (W.wsc_used) {
        cudaEventElapsedTime(&e_, W.wev[0], W.wev[1]); S.ms_small[N_SMALL] += e_;
        cudaEventElapsedTime(&e_, W.wev[2], W.wev[3]); S.ms_small[N_SMALL] += e_;
        W.wsc_used = false;
    }
    if (W.n_bad > 0) return fail(h, VD_E_BADINPUT, "%d malformed superclusters", W.n_bad);
    return VD_OK;
}

static int64_t io_bytes_of(int64_t n_sc, int64_t n_var, int64_t ref_bytes, int64_t alt_bytes) {
    const int64_t inp = ref_bytes + 8 * (n_sc + 1) + 8 * (4 * n_sc + 1) + n_var * (4 + 4 + 1 + 4) + 8 * (n_var + 1) + alt_bytes;
    const int64_t outb = 4 * n_sc * (4 + 1 + 1 + 4) + 2 * n_var * (1 + 4 + 4 + 4 + 4);
    return inp + outb;
}

extern "C" int vd_run_device_slice(vd_handle *h, const vd_batch_in *in, vd_batch_out *out,
                                   int64_t first_var, int64_t n_var, int64_t ref_bytes, int64_t alt_bytes);

extern "C" int vd_run_device(vd_handle *h, const vd_batch_in *in, vd_batch_out *out,
                             int64_t n_var, int64_t ref_bytes, int64_t alt_bytes) {
    return vd_run_device_slice(h, in, out, 0, n_var, ref_bytes, alt_bytes);
}

extern "C" int vd_run_device_slice(vd_handle *h, const vd_batch_in *in, vd_batch_out *out,
                                   int64_t first_var, int64_t n_var, int64_t ref_bytes, int64_t alt_bytes) {
    if (!h || !in || !out) return VD_E_BADINPUT;
    CK(cudaSetDevice(h->device));
    BatchDev b{in->n_sc, in->ref_off, in->ref_seq, in->rplane_seq ? in->rplane_seq : in->ref_seq, in->var_off,
               in->var_pos, in->var_rlen, in->var_type, in->alt_off, in->alt_seq, in->var_qual, in->max_qual, n_var};
    OutDev o{out->aln_score, out->aln_end_plane, out->aln_beg_plane, out->status, out->assigned,
             out->sync_group, out->ref_ed, out->query_ed, out->callq};
    h->stats = vd_stats{};
    h->stats_status_or = 0;
    h->stats.n_sc = in->n_sc; h->stats.n_var = n_var;
    h->stats.io_bytes = io_bytes_of(in->n_sc, n_var, ref_bytes, alt_bytes);
    *h->h_range = 0;
    Work &W = h->work[0];
    const OutDev base = o;                 // per-variant arrays are indexed [slot*n_var + (v - first_var)]
    o.assigned -= first_var; o.sync_group -= first_var; o.ref_ed -= first_var; o.query_ed -= first_var; o.callq -= first_var;
    int rc = chunk_plan(h, W, h->stream, b, o, base);
    if (rc == VD_OK) rc = chunk_exec(h, W);
    const int rc2 = chunk_harvest(h, W);
    return rc != VD_OK ? rc : rc2;
}

// 16-bit records from device-resident wide ones, on the handle's stream (the exchange of a multi-GPU step moves these)
extern "C" int vd_pack_device(vd_handle *h, const vd_batch_out *wide, int64_t n_sc, int64_t n_var, const vd_packed_out *packed) {
    if (!h || !wide || !packed || n_sc < 0 || n_var < 0) return VD_E_BADINPUT;
    CK(cudaSetDevice(h->device));
    if (n_sc == 0) return VD_OK;
    OutDev o{wide->aln_score, wide->aln_end_plane, wide->aln_beg_plane, wide->status, wide->assigned,
             wide->sync_group, wide->ref_ed, wide->query_ed, wide->callq};
    const int64_t nel = 4 * n_sc > 2 * n_var ? 4 * n_sc : 2 * n_var;
    VD_LAUNCH(pack_out_kernel, (unsigned)((nel + 255) / 256), 256, 0, h->stream, o, 4 * n_sc, n_var, packed->aln_score, packed->aln_planes,
              packed->status, packed->sync_group, packed->ref_ed, packed->query_ed, h->h_range);
    if (packed->callq && packed->callq != wide->callq)
        CK(cudaMemcpyAsync(packed->callq, wide->callq, 8 * (size_t)n_var, cudaMemcpyDeviceToDevice, h->stream));
    CK(cudaGetLastError());
    return VD_OK;
}
#ifdef VD_PHASE_PROF
extern "C" void vd_debug_phases(unsigned long long *out48, int reset) {       // out48: 56 words
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out48, vd::g_wsc_phase, sizeof(unsigned long long) * 48);
    cudaMemcpyFromSymbol(out48 + 48, vd::g_walk_phase, sizeof(unsigned long long) * 8);
    if (reset) { unsigned long long z[48] = {}; cudaMemcpyToSymbol(vd::g_wsc_phase, z, sizeof(z)); cudaMemcpyToSymbol(vd::g_walk_phase, z, 64); }
}
#endif
#ifdef VD_EMU
// CPU build under the SIMT emulator only (tests/simt): the bit-parallel section edit distance on its own
extern "C" int vd_emu_lev_myers64(const uint8_t *a, int m, const uint8_t *b, int n) { return vd::lev_myers64(a, m, b, n); }
#endif
extern "C" int vd_packed_overflow(const vd_handle *h) { return h && h->h_range ? (int)*h->h_range : 0; }

// error path of vd_run: nothing of this call may still be in flight when the caller gets its buffers back, and
// the handle must be reusable (no stale chunk state, no stale malformed-input count)
static int quiesce(vd_handle *h, int rc) {
    cudaStream_t ss[] = {h->s_in, h->s_plan, h->stream, h->s_wsc, h->s_epi, h->s_out};
    for (cudaStream_t s_ : ss) if (s_) cudaStreamSynchronize(s_);
    for (int c = 0; c < N_WCLS; c++) if (h->side[c]) cudaStreamSynchronize(h->side[c]);
    for (auto &W : h->work) { W.busy = false; W.n_bad = 0; }
    for (auto &sg : h->stage) sg.out_pending = false;
    cudaGetLastError();
    return rc;
}

static int run_host(vd_handle *h, const vd_batch_in *in, const vd_compact_in *cin, vd_batch_out *out, vd_packed_out *pout);
extern "C" int vd_run(vd_handle *h, const vd_batch_in *in, vd_batch_out *out) {
    if (!h || !in || !out) return VD_E_BADINPUT;
    return run_host(h, in, nullptr, out, nullptr);
}
extern "C" int vd_run_packed(vd_handle *h, const vd_batch_in *in, vd_packed_out *pout) {
    if (!h || !in || !pout) return VD_E_BADINPUT;
    return run_host(h, in, nullptr, nullptr, pout);
}
extern "C" int vd_run_compact(vd_handle *h, const vd_compact_in *cin, vd_packed_out *pout) {
    if (!h || !cin || !pout) return VD_E_BADINPUT;
    return run_host(h, nullptr, cin, nullptr, pout);
}
extern "C" void *vd_host_alloc(int64_t bytes) {
    void *p = nullptr;
    if (bytes <= 0 || cudaHostAlloc(&p, (size_t)bytes, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    return p;
}
extern "C" void vd_host_free(void *p) { if (p) cudaFreeHost(p); }

// vd_run (out) / vd_run_packed (pout): exactly one of the two is given; the batch as vd_batch_in (in) or in compact form
// (cin, vd_run_compact): then the offsets are rebuilt per chunk on the device, 0-based
static int run_host(vd_handle *h, const vd_batch_in *in, const vd_compact_in *cin, vd_batch_out *out, vd_packed_out *pout) {
    CK(cudaSetDevice(h->device));
    const int64_t n_sc = in ? in->n_sc : cin->n_sc;
    if (n_sc < 0) return fail(h, VD_E_BADINPUT, "negative n_sc");
    h->stats = vd_stats{};
    h->stats_status_or = 0;
    *h->h_range = 0;
    if (n_sc == 0) return VD_OK;
    const int64_t n_var = in ? in->var_off[4 * n_sc] : cin->n_var;
    const int64_t ref_bytes = in ? in->ref_off[n_sc] : cin->ref_bytes;
    const int64_t alt_bytes = in ? (n_var ? in->alt_off[n_var] : 0) : cin->alt_bytes;
    const float max_qual = in ? in->max_qual : cin->max_qual;
    const bool has_rplane = in ? in->rplane_seq != nullptr : cin->rplane_seq != nullptr;
    if (n_var < 0 || ref_bytes < 0 || alt_bytes < 0) return fail(h, VD_E_BADINPUT, "negative sizes");
    h->stats.n_sc = n_sc; h->stats.n_var = n_var;
    h->stats.io_bytes = io_bytes_of(n_sc, n_var, ref_bytes, alt_bytes);

    // Chunked, double-buffered pipeline: while chunk i computes, chunk i+1's inputs are copied
    // in on s_in and chunk i-1's results are copied out on s_out.  Every chunk keeps the batch's
    // ABSOLUTE offsets (ref_off / var_off / alt_off values); the data pointers handed to the
    // kernels are shifted back by the chunk's first byte / variant instead.
    // chunk boundaries: small chunks first and last (the first H2D and the last D2H are not hidden
    // behind any compute), full-size chunks in between
    const int64_t CH = h->chunk_sc;
    std::vector<int64_t> cut{0};
    {
        int64_t s0 = 0;
        const int64_t ramp[2] = {CH / 4, CH / 2};
        for (int k = 0; k < 2 && h->ramp && n_sc - s0 > 2 * CH && ramp[k] > 0; k++) { s0 += ramp[k]; cut.push_back(s0); }
        while (n_sc - s0 > CH + CH / 2) { s0 += CH; cut.push_back(s0); }
        if (h->ramp && n_sc - s0 > CH / 2 && CH / 4 > 0) { s0 = n_sc - CH / 4; cut.push_back(s0); }
        cut.push_back(n_sc);
        if (cin) {                  // the compact form can only be cut where its block index has an entry
            std::vector<int64_t> c2{0};
            for (size_t k = 1; k + 1 < cut.size(); k++) {
                const int64_t a = cut[k] / VD_COMPACT_BLOCK * VD_COMPACT_BLOCK;
                if (a > c2.back()) c2.push_back(a);
            }
            c2.push_back(n_sc);
            cut.swap(c2);
        }
    }
    const int n_chunks = (int)cut.size() - 1;
    int64_t h2d = 0, d2h = 0;
    int rc_all = VD_OK;
    struct Range { int64_t s0, s1, v0, v1, r0, r1, a0, a1; };
    auto range_of = [&](int i) {
        Range r;
        r.s0 = cut[i]; r.s1 = cut[i + 1];
        if (cin) {
            auto at = [&](const int64_t *blk, int64_t s, int64_t total) { return s >= n_sc ? total : blk[s / VD_COMPACT_BLOCK]; };
            r.v0 = at(cin->blk_var, r.s0, n_var); r.v1 = at(cin->blk_var, r.s1, n_var);
            r.r0 = at(cin->blk_ref, r.s0, ref_bytes); r.r1 = at(cin->blk_ref, r.s1, ref_bytes);
            r.a0 = at(cin->blk_alt, r.s0, alt_bytes); r.a1 = at(cin->blk_alt, r.s1, alt_bytes);
            return r;
        }
        r.v0 = in->var_off[4 * r.s0]; r.v1 = in->var_off[4 * r.s1];
        r.r0 = in->ref_off[r.s0]; r.r1 = in->ref_off[r.s1];
        r.a0 = n_var ? in->alt_off[r.v0] : 0; r.a1 = n_var ? in->alt_off[r.v1] : 0;
        return r;
    };
    auto upload = [&](int i) -> int {
        vd_handle::Stage &sg = h->stage[i % vd_handle::NST];
        // the input buffers of this stage were read by chunk i-NST: copy in only behind its last kernel
        if (h->work[i % vd_handle::NST].busy) CK(cudaStreamWaitEvent(h->s_in, h->work[i % vd_handle::NST].ev[3], 0));
        const Range r = range_of(i);
        const int64_t ns = r.s1 - r.s0, nv = r.v1 - r.v0;
#define UP(buf, src, bytes) do { CK(sg.buf.ensure((size_t)(bytes) + 16)); \
        if ((bytes) > 0) CK(cudaMemcpyAsync(sg.buf.p, (src), (size_t)(bytes), cudaMemcpyHostToDevice, h->s_in)); \
        h2d += (bytes); } while (0)
        if (cin) {
            // compact arrays in; sizes widened and prefix-summed on the copy stream behind them (0-based offsets)
            UP(c_ref_len, cin->ref_len + r.s0, 2 * ns);
            UP(in_ref_seq, cin->ref_seq + r.r0, r.r1 - r.r0);
            if (cin->rplane_seq) UP(in_rplane, cin->rplane_seq + r.r0, r.r1 - r.r0);
            UP(c_hap_nvar, cin->hap_nvar + 4 * r.s0, 4 * ns);
            UP(c_pos16, cin->var_pos + r.v0, 2 * nv);
            UP(c_rlen16, cin->var_rlen + r.v0, 2 * nv);
            UP(c_alen16, cin->alt_len + r.v0, 2 * nv);
            UP(in_var_type, cin->var_type + r.v0, nv);
            UP(in_alt_seq, cin->alt_seq + r.a0, r.a1 - r.a0);
            UP(in_var_qual, cin->var_qual + r.v0, 4 * nv);
            CK(sg.in_ref_off.ensure(8 * (size_t)(ns + 1) + 16)); CK(sg.in_var_off.ensure(8 * (size_t)(4 * ns + 1) + 16));
            CK(sg.in_alt_off.ensure(8 * (size_t)(nv + 1) + 16));
            CK(sg.in_var_pos.ensure(4 * (size_t)nv + 16)); CK(sg.in_var_rlen.ensure(4 * (size_t)nv + 16));
            const int64_t nel = (4 * ns > nv ? 4 * ns : nv) + 1;
            VD_LAUNCH(unpack_sizes_kernel, (unsigned)((nel + 255) / 256), 256, 0, h->s_in, (const u16 *)sg.c_ref_len.p, ns, (const u8 *)sg.c_hap_nvar.p,
                      (const u16 *)sg.c_alen16.p, nv, (int64_t *)sg.in_ref_off.p, (int64_t *)sg.in_var_off.p, (int64_t *)sg.in_alt_off.p,
                      (const u16 *)sg.c_pos16.p, (const u16 *)sg.c_rlen16.p, (int32_t *)sg.in_var_pos.p, (int32_t *)sg.in_var_rlen.p);
            int64_t *offs3[3] = {(int64_t *)sg.in_ref_off.p, (int64_t *)sg.in_var_off.p, (int64_t *)sg.in_alt_off.p};
            const int64_t cnt3[3] = {ns + 1, 4 * ns + 1, nv + 1};
            for (int k = 0; k < 3; k++) {
                size_t tmp = 0;
                cub::DeviceScan::ExclusiveSum(nullptr, tmp, offs3[k], offs3[k], (int)cnt3[k], h->s_in);
                CK(sg.c_tmp.ensure(tmp));
                cub::DeviceScan::ExclusiveSum(sg.c_tmp.p, tmp, offs3[k], offs3[k], (int)cnt3[k], h->s_in);
            }
            h->stats.n_launches += 4;
            CK(cudaEventRecord(sg.in_done, h->s_in));
            return VD_OK;
        }
        UP(in_ref_off, in->ref_off + r.s0, 8 * (ns + 1));
        UP(in_ref_seq, in->ref_seq + r.r0, r.r1 - r.r0);
        if (in->rplane_seq) UP(in_rplane, in->rplane_seq + r.r0, r.r1 - r.r0);
        UP(in_var_off, in->var_off + 4 * r.s0, 8 * (4 * ns + 1));
        UP(in_var_pos, in->var_pos + r.v0, 4 * nv);
        UP(in_var_rlen, in->var_rlen + r.v0, 4 * nv);
        UP(in_var_type, in->var_type + r.v0, nv);
        if (n_var) UP(in_alt_off, in->alt_off + r.v0, 8 * (nv + 1));
        else { CK(sg.in_alt_off.ensure(16)); CK(cudaMemsetAsync(sg.in_alt_off.p, 0, 16, h->s_in)); }
        UP(in_alt_seq, in->alt_seq + r.a0, r.a1 - r.a0);
        UP(in_var_qual, in->var_qual + r.v0, 4 * nv);
#undef UP
        CK(cudaEventRecord(sg.in_done, h->s_in));
        return VD_OK;
    };

    const bool trace = getenv("VD_TRACE") != nullptr;
    auto now_ms = [] { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; };
    const double t_start = now_ms();
    // chunk i: [H2D on s_in] -> [plan pass on s_plan] -> [kernels on st + side streams] -> [D2H on s_out];
    // the host never waits for a chunk's kernels, only for its plan counters (and, NST chunks later, for
    // its end event when the work set is reused)
    auto plan_chunk = [&](int i) -> int {
        vd_handle::Stage &sg = h->stage[i % vd_handle::NST];
        Work &W = h->work[i % vd_handle::NST];
        int rcp = chunk_harvest(h, W);                       // chunk i-NST used this work set
        if (rcp != VD_OK && rcp != VD_E_BADINPUT) return rcp;
        if (rcp != VD_OK) rc_all = rcp;
        const Range r = range_of(i);
        const int64_t ns = r.s1 - r.s0, nv = r.v1 - r.v0;
        // result buffers of this stage are free once chunk i-NST's copy-out has finished
        if (sg.out_pending) { CK(cudaStreamWaitEvent(h->s_plan, sg.out_done, 0)); sg.out_pending = false; }
        CK(sg.o_score.ensure(16 * (size_t)ns)); CK(sg.o_endp.ensure(4 * (size_t)ns)); CK(sg.o_begp.ensure(4 * (size_t)ns));
        CK(sg.o_status.ensure(16 * (size_t)ns)); CK(sg.o_assigned.ensure(2 * (size_t)nv + 16));
        CK(sg.o_sg.ensure(8 * (size_t)nv + 16)); CK(sg.o_red.ensure(8 * (size_t)nv + 16));
        CK(sg.o_qed.ensure(8 * (size_t)nv + 16)); CK(sg.o_callq.ensure(8 * (size_t)nv + 16));
        CK(cudaStreamWaitEvent(h->s_plan, sg.in_done, 0));
        // offsets of a vd_batch_in chunk are the batch's own (absolute): the data pointers are shifted back instead;
        // a compact chunk's offsets were rebuilt from 0
        const int64_t sr = cin ? 0 : r.r0, sv = cin ? 0 : r.v0, sa = cin ? 0 : r.a0;
        const u8 *d_ref = (const u8 *)sg.in_ref_seq.p - sr;
        const u8 *d_rpl = has_rplane ? (const u8 *)sg.in_rplane.p - sr : d_ref;
        BatchDev b{(int)ns, (const int64_t *)sg.in_ref_off.p, d_ref, d_rpl, (const int64_t *)sg.in_var_off.p,
                   (const int32_t *)sg.in_var_pos.p - sv, (const int32_t *)sg.in_var_rlen.p - sv,
                   (const u8 *)sg.in_var_type.p - sv, (const int64_t *)sg.in_alt_off.p - sv,
                   (const u8 *)sg.in_alt_seq.p - sa, (const float *)sg.in_var_qual.p - sv, max_qual, nv};
        OutDev base{(int32_t *)sg.o_score.p, (u8 *)sg.o_endp.p, (u8 *)sg.o_begp.p, (u32 *)sg.o_status.p,
                    (u8 *)sg.o_assigned.p, (int32_t *)sg.o_sg.p, (int32_t *)sg.o_red.p, (int32_t *)sg.o_qed.p,
                    (float *)sg.o_callq.p};
        OutDev o = base;                       // per-variant arrays are indexed [slot*nv + (v - v0)]
        o.assigned -= sv; o.sync_group -= sv; o.ref_ed -= sv; o.query_ed -= sv; o.callq -= sv;
        return chunk_plan(h, W, h->s_plan, b, o, base);
    };

    int rc = upload(0);
    if (rc != VD_OK) return quiesce(h, rc);
    rc = plan_chunk(0);
    if (rc != VD_OK) return quiesce(h, rc);
    for (int i = 0; i < n_chunks; i++) {
        const double t_c0 = now_ms();
        vd_handle::Stage &sg = h->stage[i % vd_handle::NST];
        Work &W = h->work[i % vd_handle::NST];
        const Range r = range_of(i);
        const int64_t ns = r.s1 - r.s0, nv = r.v1 - r.v0;
        if (i + 1 < n_chunks) { rc = upload(i + 1); if (rc != VD_OK) return quiesce(h, rc); }
        const double t_c1 = now_ms();
        rc = chunk_exec(h, W);                 // returns with the chunk's kernels issued
        if (rc != VD_OK) return quiesce(h, rc);
        const double t_c2 = now_ms();
        // the plan pass of the next chunk runs beside this chunk's kernels
        if (i + 1 < n_chunks) { rc = plan_chunk(i + 1); if (rc != VD_OK) return quiesce(h, rc); }
        if (trace) fprintf(stderr, "[vd_run] chunk %d: start %.2f ms, upload %.2f ms, exec %.2f ms, plan next %.2f ms\n",
                           i, t_c0 - t_start, t_c1 - t_c0, t_c2 - t_c1, now_ms() - t_c2);

        CK(cudaStreamWaitEvent(h->s_out, W.ev[3], 0));      // copy-out behind the chunk's last kernel
#define DOWN(dst, buf, off, bytes) do { if ((bytes) > 0) CK(cudaMemcpyAsync((dst), (const u8 *)sg.buf.p + (off), \
        (size_t)(bytes), cudaMemcpyDeviceToHost, h->s_out)); d2h += (bytes); } while (0)
        if (out) {
            DOWN(out->aln_score + 4 * r.s0, o_score, 0, 16 * ns);
            DOWN(out->aln_end_plane + 4 * r.s0, o_endp, 0, 4 * ns);
            DOWN(out->aln_beg_plane + 4 * r.s0, o_begp, 0, 4 * ns);
            DOWN(out->status + 4 * r.s0, o_status, 0, 16 * ns);
            for (int slot = 0; slot < 2; slot++) {
                const int64_t ho = slot * n_var + r.v0, so = slot * nv;
                DOWN(out->assigned + ho, o_assigned, so, nv);
                DOWN(out->sync_group + ho, o_sg, 4 * so, 4 * nv);
                DOWN(out->ref_ed + ho, o_red, 4 * so, 4 * nv);
                DOWN(out->query_ed + ho, o_qed, 4 * so, 4 * nv);
                DOWN(out->callq + ho, o_callq, 4 * so, 4 * nv);
            }
        } else {
            // 16-bit records: narrowed on the device behind the chunk's last kernel, then copied out
            CK(sg.p_score.ensure(8 * (size_t)ns + 16)); CK(sg.p_planes.ensure(4 * (size_t)ns + 16)); CK(sg.p_status.ensure(8 * (size_t)ns + 16));
            CK(sg.p_sg.ensure(4 * (size_t)nv + 16)); CK(sg.p_red.ensure(4 * (size_t)nv + 16)); CK(sg.p_qed.ensure(4 * (size_t)nv + 16));
            OutDev wide{(int32_t *)sg.o_score.p, (u8 *)sg.o_endp.p, (u8 *)sg.o_begp.p, (u32 *)sg.o_status.p, (u8 *)sg.o_assigned.p,
                        (int32_t *)sg.o_sg.p, (int32_t *)sg.o_red.p, (int32_t *)sg.o_qed.p, (float *)sg.o_callq.p};
            const int64_t nel = 4 * ns > 2 * nv ? 4 * ns : 2 * nv;
            VD_LAUNCH(pack_out_kernel, (unsigned)((nel + 255) / 256), 256, 0, h->s_out, wide, 4 * ns, nv, (u16 *)sg.p_score.p, (u8 *)sg.p_planes.p,
                      (u16 *)sg.p_status.p, (u16 *)sg.p_sg.p, (u16 *)sg.p_red.p, (u16 *)sg.p_qed.p, h->h_range);
            h->stats.n_launches++;
            DOWN(pout->aln_score + 4 * r.s0, p_score, 0, 8 * ns);
            DOWN(pout->aln_planes + 4 * r.s0, p_planes, 0, 4 * ns);
            DOWN(pout->status + 4 * r.s0, p_status, 0, 8 * ns);
            for (int slot = 0; slot < 2; slot++) {
                const int64_t ho = slot * n_var + r.v0, so = slot * nv;
                DOWN(pout->sync_group + ho, p_sg, 2 * so, 2 * nv);
                DOWN(pout->ref_ed + ho, p_red, 2 * so, 2 * nv);
                DOWN(pout->query_ed + ho, p_qed, 2 * so, 2 * nv);
                DOWN(pout->callq + ho, o_callq, 4 * so, 4 * nv);
            }
        }
#undef DOWN
        CK(cudaEventRecord(sg.out_done, h->s_out));
        sg.out_pending = true;
    }
    for (auto &W : h->work) {
        const int rch = chunk_harvest(h, W);
        if (rch != VD_OK && rch != VD_E_BADINPUT) return quiesce(h, rch);
        if (rch != VD_OK) rc_all = rch;
    }
    const double t_e0 = now_ms();
    CK(cudaStreamSynchronize(h->s_out));
    if (trace) fprintf(stderr, "[vd_run] drain %.2f ms, total %.2f ms\n", now_ms() - t_e0, now_ms() - t_start);
    for (auto &sg : h->stage) sg.out_pending = false;
    h->stats.h2d_bytes = h2d;
    h->stats.d2h_bytes = d2h;
    if (rc_all != VD_OK) return rc_all;
    if (pout && *h->h_range) return fail(h, VD_E_RANGE, "a result does not fit the 16-bit records: use vd_run");
    // fatal reference conditions are reported; results stay available for inspection
    if (h->stats_status_or & VD_ST_ERR_MASK)
        for (int64_t i = 0; i < 4 * n_sc; i++) {
            const unsigned sti = out ? out->status[i] : pout->status[i];
            if (sti & VD_ST_ERR_MASK)
                return fail(h, (sti & VD_ST_ERR_BADINPUT) ? VD_E_BADINPUT : VD_E_ALIGN,
                            "alignment %lld of supercluster %lld: status 0x%x%s", (long long)(i & 3), (long long)(i >> 2), sti,
                            (sti & VD_ST_ERR_BADINPUT) ? " (input the kernels cannot process, e.g. more than 8 swap sources for one row)" : "");
        }
    return VD_OK;
}

// ---- cluster-growing stage: batches of affine-gap wavefront problems ------------------
extern "C" int vd_wf_batch(vd_handle *h, int mode, int n, const int64_t *q_off, const uint8_t *q_seq, const int64_t *t_off,
                           const uint8_t *t_seq, const int32_t *main_diag, const int32_t *main_diag_start, const int32_t *max_score,
                           const uint8_t *reverse, int sub, int open, int extend, int32_t *result) {
    if (!h || n < 0 || !q_off || !t_off || !result || (mode != WF_MODE_REACH && mode != WF_MODE_SCORE)) return VD_E_BADINPUT;
    if (mode == WF_MODE_REACH && (!main_diag || !main_diag_start || !max_score || !reverse)) return VD_E_BADINPUT;
    if (sub < 0 || open < 0 || extend < 1) return fail(h, VD_E_BADINPUT, "penalties: sub %d open %d extend %d", sub, open, extend);
    if (n == 0) return VD_OK;
    CK(cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    const int64_t qb = q_off[n], tb = t_off[n];
    std::vector<int64_t> soff((size_t)n + 1);
    soff[0] = 0;
    for (int i = 0; i < n; i++) {
        const int ql = (int)(q_off[i + 1] - q_off[i]), tl = (int)(t_off[i + 1] - t_off[i]);
        if (ql < 1 || tl < 1) return fail(h, VD_E_BADINPUT, "problem %d: empty string", i);
        soff[i + 1] = soff[i] + wf_scratch_ints(ql, tl, sub, open, extend);
    }
    if (4 * soff[n] > h->scratch_budget) return fail(h, VD_E_NOMEM, "wavefront rings need %lld bytes", (long long)(4 * soff[n]));
    // one staging block: offsets, strings, per-problem parameters, results
    size_t o_qoff = 0, o_toff = o_qoff + 8 * (size_t)(n + 1), o_soff = o_toff + 8 * (size_t)(n + 1), o_md = o_soff + 8 * (size_t)(n + 1);
    size_t o_mds = o_md + 4 * (size_t)n, o_ms = o_mds + 4 * (size_t)n, o_res = o_ms + 4 * (size_t)n, o_rev = o_res + 4 * (size_t)n;
    size_t o_q = (o_rev + n + 15) & ~(size_t)15, o_t = (o_q + qb + 15) & ~(size_t)15, o_sel = (o_t + tb + 15) & ~(size_t)15;
    size_t total = o_sel + 4 * (size_t)n + 16;
    // Problems whose wavefront can grow to wf_block_min diagonals or more get a whole block (WF_BLOCK threads, WF_WIDE for
    // the wider ones, a cluster of WF_CLUSTER such blocks from wf_cluster_min diagonals on), the others a warp: sel holds
    // the warp problems, then the block problems, the wide ones, the cluster ones.
    std::vector<int> sel((size_t)n);
    int n_form[4] = {0, 0, 0, 0};
    {
        std::vector<u8> form((size_t)n);
        for (int i = 0; i < n; i++) {
            int64_t width = (q_off[i + 1] - q_off[i]) + (t_off[i + 1] - t_off[i]) - 1;
            if (mode == WF_MODE_REACH) width = std::min<int64_t>(width, 2 * (int64_t)std::max(max_score[i], 0) + 1);
            form[i] = width >= h->wf_cluster_min ? 3 : width >= 4 * (int64_t)h->wf_block_min ? 2 : width >= h->wf_block_min ? 1 : 0;
            n_form[form[i]]++;
        }
        int at[4] = {0, n_form[0], n_form[0] + n_form[1], n_form[0] + n_form[1] + n_form[2]};
        for (int i = 0; i < n; i++) sel[(size_t)at[form[i]]++] = i;
    }
    const int n_warp = n_form[0], n_block = n_form[1], n_wide = n_form[2], n_cluster = n_form[3];
    CK(h->wf_in.ensure(total));
    CK(h->wf_scratch.ensure(4 * (size_t)soff[n] + 16));
    u8 *d = (u8 *)h->wf_in.p;
    CK(cudaMemcpyAsync(d + o_qoff, q_off, 8 * (size_t)(n + 1), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d + o_toff, t_off, 8 * (size_t)(n + 1), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d + o_soff, soff.data(), 8 * (size_t)(n + 1), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d + o_q, q_seq, (size_t)qb, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d + o_t, t_seq, (size_t)tb, cudaMemcpyHostToDevice, st));
    if (mode == WF_MODE_REACH) {
        CK(cudaMemcpyAsync(d + o_md, main_diag, 4 * (size_t)n, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(d + o_mds, main_diag_start, 4 * (size_t)n, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(d + o_ms, max_score, 4 * (size_t)n, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(d + o_rev, reverse, (size_t)n, cudaMemcpyHostToDevice, st));
    }
    WfBatch B{n, (const int64_t *)(d + o_qoff), (const int64_t *)(d + o_toff), d + o_q, d + o_t, (const int32_t *)(d + o_md),
              (const int32_t *)(d + o_mds), (const int32_t *)(d + o_ms), d + o_rev, (const int64_t *)(d + o_soff), (int32_t *)h->wf_scratch.p,
              (int32_t *)(d + o_res), sub, open, extend, mode};
    CK(cudaMemcpyAsync(d + o_sel, sel.data(), 4 * (size_t)n, cudaMemcpyHostToDevice, st));
    const int *d_sel = (const int *)(d + o_sel);
    // the few wide problems beside the many narrow ones: every form on its own stream
    const bool beside = !h->serial && (n_cluster > 0) + (n_wide > 0) + (n_block > 0) + (n_warp > 0) > 1;
    cudaStream_t s_cluster = beside ? h->side[2] : st, s_wide = beside ? h->side[0] : st, s_block = beside ? h->side[1] : st;
    if (beside) {
        CK(cudaEventRecord(h->wf_ev[0], st));
        CK(cudaStreamWaitEvent(s_cluster, h->wf_ev[0], 0));
        CK(cudaStreamWaitEvent(s_wide, h->wf_ev[0], 0));
        CK(cudaStreamWaitEvent(s_block, h->wf_ev[0], 0));
    }
#ifndef VD_EMU
    if (n_cluster > 0) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)n_cluster * WF_CLUSTER); cfg.blockDim = dim3(WF_WIDE); cfg.dynamicSmemBytes = 0; cfg.stream = s_cluster;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = WF_CLUSTER; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        CK(cudaLaunchKernelEx(&cfg, wf_kernel<WF_WIDE, WF_CLUSTER>, B, d_sel + n_warp + n_block + n_wide, n_cluster));
    }
#endif
    if (n_wide > 0) VD_LAUNCH(wf_kernel<WF_WIDE>, n_wide, WF_WIDE, 0, s_wide, B, d_sel + n_warp + n_block, n_wide);
    if (n_block > 0) VD_LAUNCH(wf_kernel<WF_BLOCK>, n_block, WF_BLOCK, 0, s_block, B, d_sel + n_warp, n_block);
    if (n_warp > 0) VD_LAUNCH(wf_kernel<32>, (n_warp + 3) / 4, 128, 0, st, B, d_sel, n_warp);
    if (beside) {
        CK(cudaEventRecord(h->wf_ev[1], s_wide));
        CK(cudaEventRecord(h->wf_ev[2], s_block));
        CK(cudaEventRecord(h->wf_ev[3], s_cluster));
        CK(cudaStreamWaitEvent(st, h->wf_ev[1], 0));
        CK(cudaStreamWaitEvent(st, h->wf_ev[2], 0));
        CK(cudaStreamWaitEvent(st, h->wf_ev[3], 0));
    }
    const auto t_sync = std::chrono::steady_clock::now();
    CK(cudaMemcpyAsync(result, d + o_res, 4 * (size_t)n, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    if (getenv("VD_WF_STATS")) {                                                      // one line per batch: what it held, how long it ran
        int64_t wmax = 0, lmax = 0, smax = 0;
        for (int i = 0; i < n; i++) {
            const int64_t nd = (q_off[i + 1] - q_off[i]) + (t_off[i + 1] - t_off[i]) - 1;
            lmax = std::max(lmax, nd);
            if (mode == WF_MODE_REACH) { smax = std::max<int64_t>(smax, max_score[i]); wmax = std::max(wmax, std::min<int64_t>(nd, 2 * (int64_t)max_score[i] + 1)); }
        }
        fprintf(stderr, "vd_wf_batch mode %d: %d problems (%d on a block, %d on a wide block, %d on a cluster), longest %lld diagonals, largest budget %lld (width %lld), %.1f ms\n", mode, n, n_block, n_wide, n_cluster,
                (long long)lmax, (long long)smax, (long long)wmax, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_sync).count());
    }
    return VD_OK;
}

// ---- `--distance` pass: batches of affine-gap alignments with their CIGARs -----------------
extern "C" int vd_swg_align_batch(vd_handle *h, int n, const int64_t *q_off, const uint8_t *q_seq, const int64_t *t_off,
                                  const uint8_t *t_seq, int sub, int open, int extend, int32_t *score, int32_t *cigar) {
    if (!h || n < 0 || !q_off || !t_off || !score || !cigar) return VD_E_BADINPUT;
    if (n == 0) return VD_OK;
    // pass 1: scores (they size the wavefront storage of pass 2)
    int rc = vd_wf_batch(h, WF_MODE_SCORE, n, q_off, q_seq, t_off, t_seq, nullptr, nullptr, nullptr, nullptr, sub, open, extend, score);
    if (rc != VD_OK) return rc;
    cudaStream_t st = h->stream;
    const int64_t qb = q_off[n], tb = t_off[n];
    // Every score of an alignment keeps its wavefronts and flags: a structural variant's alignment takes gigabytes.  The batch
    // runs in rounds of consecutive problems whose storage fits the budget together; rounds follow each other on the stream
    // and reuse the same storage.  boff[i] = offset of problem i inside its round.
    const int64_t cigar_bytes = ((4 * (qb + tb) + 15) / 16) * 16;
    const int64_t budget = h->scratch_budget - cigar_bytes;
    std::vector<int64_t> boff((size_t)n + 1, 0);
    std::vector<int> round_begin{0};
    int64_t at_byte = 0, round_max = 0;
    for (int i = 0; i < n; i++) {
        const int64_t need = wf_cigar_bytes((int)(q_off[i + 1] - q_off[i]), (int)(t_off[i + 1] - t_off[i]), score[i]);
        if (need > budget) return fail(h, VD_E_NOMEM, "alignment %d: its wavefronts need %lld bytes", i, (long long)need);
        if (at_byte + need > budget) { round_begin.push_back(i); at_byte = 0; }
        boff[i] = at_byte;
        at_byte += need;
        round_max = std::max(round_max, at_byte);
    }
    round_begin.push_back(n);
    // vd_wf_batch left offsets and strings staged in wf_in: lay the same block out again (same sizes)
    size_t o_qoff = 0, o_toff = o_qoff + 8 * (size_t)(n + 1), o_soff = o_toff + 8 * (size_t)(n + 1), o_md = o_soff + 8 * (size_t)(n + 1);
    size_t o_mds = o_md + 4 * (size_t)n, o_ms = o_mds + 4 * (size_t)n, o_res = o_ms + 4 * (size_t)n, o_rev = o_res + 4 * (size_t)n;
    size_t o_q = (o_rev + n + 15) & ~(size_t)15, o_t = (o_q + qb + 15) & ~(size_t)15;
    u8 *d = (u8 *)h->wf_in.p;
    CK(h->wf_scratch.ensure((size_t)(cigar_bytes + round_max) + 64));
    CK(cudaMemcpyAsync(d + o_soff, boff.data(), 8 * (size_t)(n + 1), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d + o_ms, score, 4 * (size_t)n, cudaMemcpyHostToDevice, st));
    int32_t *d_cigar = (int32_t *)h->wf_scratch.p;
    WfCigarBatch B{n, (const int64_t *)(d + o_qoff), (const int64_t *)(d + o_toff), d + o_q, d + o_t, (const int32_t *)(d + o_ms),
                   (const int64_t *)(d + o_soff), (u8 *)h->wf_scratch.p + cigar_bytes, d_cigar, (int32_t *)(d + o_res), sub, open, extend};
    // warp or block per problem by the width its wavefront reached (as vd_wf_batch; the score is known here); sel holds, round
    // after round, the warp problems, the block problems, the wide ones
    size_t o_sel = (o_t + tb + 15) & ~(size_t)15;
    std::vector<int> sel((size_t)n);
    std::vector<u8> form((size_t)n);
    for (int i = 0; i < n; i++) {
        const int64_t width = std::min<int64_t>((q_off[i + 1] - q_off[i]) + (t_off[i + 1] - t_off[i]) - 1, 2 * (int64_t)std::max(score[i], 0) + 1);
        form[i] = width >= 4 * (int64_t)h->wf_block_min ? 2 : width >= h->wf_block_min ? 1 : 0;
    }
    std::vector<int> cnt(3 * (round_begin.size() - 1), 0);
    for (size_t r = 0; r + 1 < round_begin.size(); r++) {
        int *c = &cnt[3 * r];
        for (int i = round_begin[r]; i < round_begin[r + 1]; i++) c[form[i]]++;
        int at[3] = {round_begin[r], round_begin[r] + c[0], round_begin[r] + c[0] + c[1]};
        for (int i = round_begin[r]; i < round_begin[r + 1]; i++) sel[(size_t)at[form[i]]++] = i;
    }
    CK(cudaMemcpyAsync(d + o_sel, sel.data(), 4 * (size_t)n, cudaMemcpyHostToDevice, st));
    const int *d_sel = (const int *)(d + o_sel);
    for (size_t r = 0; r + 1 < round_begin.size(); r++) {
        const int n_warp = cnt[3 * r], n_block = cnt[3 * r + 1], n_wide = cnt[3 * r + 2];
        const int *rs = d_sel + round_begin[r];
        if (n_wide > 0) VD_LAUNCH(wf_cigar_kernel<WF_WIDE>, n_wide, WF_WIDE, 0, st, B, rs + n_warp + n_block, n_wide);
        if (n_block > 0) VD_LAUNCH(wf_cigar_kernel<WF_BLOCK>, n_block, WF_BLOCK, 0, st, B, rs + n_warp, n_block);
        if (n_warp > 0) VD_LAUNCH(wf_cigar_kernel<32>, (n_warp + 3) / 4, 128, 0, st, B, rs, n_warp);
    }
    if (getenv("VD_WF_STATS"))
        fprintf(stderr, "vd_swg_align_batch: %d alignments in %d round(s), %.1f MB of wavefronts per round at most\n", n, (int)round_begin.size() - 1, round_max / 1e6);
    CK(cudaMemcpyAsync(score, d + o_res, 4 * (size_t)n, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(cigar, d_cigar, 4 * (size_t)(qb + tb), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    return VD_OK;
}
