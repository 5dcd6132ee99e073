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
#include <string>
#include <vector>

#include <cub/device/device_scan.cuh>

#include "vd_kernels.cuh"
#include "vd_wave.cuh"

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

struct vd_handle {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[8] = {};
    cudaStream_t side[vd::N_WCLS] = {};     // one stream per wavefront class: classes run concurrently
    cudaEvent_t sev[vd::N_WCLS][3] = {};
    int64_t scratch_budget = 0;
    int num_sms = 148;
    std::string err;
    vd_stats stats = {};
    int force_class = -1;           // VD_FORCE_CLASS env: testing hook (1 wave, 2 scalar slab)
    // staged input / output (vd_run)
    DevBuf in_ref_off, in_ref_seq, in_rplane, in_var_off, in_var_pos, in_var_rlen, in_var_type,
           in_alt_off, in_alt_seq, in_var_qual;
    DevBuf o_score, o_endp, o_begp, o_status, o_assigned, o_sg, o_red, o_qed, o_callq;
    // work
    DevBuf plan, list, counters, bytes, offs, cubtmp, slab, hap_ok, wave_desc;
    PlanCounters *h_counters = nullptr;     // pinned
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
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) return VD_E_NODEVICE;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return VD_E_NODEVICE;
    if (prop.major != 10) return VD_E_NODEVICE;          // built for sm_100a only
    if (cudaSetDevice(device) != cudaSuccess) return VD_E_NODEVICE;
    vd_handle *h = new vd_handle();
    h->device = device;
    h->num_sms = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { delete h; return VD_E_CUDA; }
    for (auto &e : h->ev) cudaEventCreate(&e);
    for (int c = 0; c < N_WCLS; c++) {
        cudaStreamCreateWithFlags(&h->side[c], cudaStreamNonBlocking);
        for (auto &e : h->sev[c]) cudaEventCreate(&e);
    }
    cudaMallocHost((void **)&h->h_counters, sizeof(PlanCounters));
    if (scratch_bytes <= 0) {
        size_t fr = 0, tot = 0;
        cudaMemGetInfo(&fr, &tot);
        scratch_bytes = (int64_t)(fr * 0.6);
    }
    h->scratch_budget = scratch_bytes;
    if (const char *fc = getenv("VD_FORCE_CLASS")) h->force_class = atoi(fc);
    cudaFuncSetAttribute(tiny_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (TINY_TPB / 4) * TINY_SC_BYTES + TINY_TPB * TINY_CAP);
    wave_configure();
    *out = h;
    return VD_OK;
}

extern "C" void vd_destroy(vd_handle *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    DevBuf *bufs[] = {&h->in_ref_off, &h->in_ref_seq, &h->in_rplane, &h->in_var_off, &h->in_var_pos,
                      &h->in_var_rlen, &h->in_var_type, &h->in_alt_off, &h->in_alt_seq, &h->in_var_qual,
                      &h->o_score, &h->o_endp, &h->o_begp, &h->o_status, &h->o_assigned, &h->o_sg, &h->o_red,
                      &h->o_qed, &h->o_callq, &h->plan, &h->list, &h->counters, &h->bytes, &h->offs,
                      &h->cubtmp, &h->slab, &h->hap_ok, &h->wave_desc};
    for (DevBuf *b : bufs) b->release();
    if (h->h_counters) cudaFreeHost(h->h_counters);
    for (auto &e : h->ev) if (e) cudaEventDestroy(e);
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

// The whole path on device-resident buffers.
static int run_resident(vd_handle *h, const BatchDev &in, const OutDev &out) {
    cudaStream_t st = h->stream;
    const int n_sc = in.n_sc;
    const int64_t n_var = in.n_var;
    vd_stats &S = h->stats;
    S.n_sc = n_sc; S.n_var = n_var;
    S.n_launches = 0; S.n_short = S.n_long = 0; S.spill_bytes = 0;
    S.ms_short = S.ms_long_fwd = S.ms_long_bwd = S.ms_long_walk = S.ms_plan = S.ms_long_wall = 0;
    if (n_sc == 0) { S.cells = 0; S.ms_total = 0; return VD_OK; }

    CK(cudaEventRecord(h->ev[0], st));
    CK(cudaMemsetAsync(out.assigned, 0, 2 * n_var, st));
    CK(cudaMemsetAsync(out.sync_group, 0, 8 * n_var, st));
    CK(cudaMemsetAsync(out.ref_ed, 0, 8 * n_var, st));
    CK(cudaMemsetAsync(out.query_ed, 0, 8 * n_var, st));
    CK(cudaMemsetAsync(out.callq, 0, 8 * n_var, st));
    CK(cudaMemsetAsync(out.status, 0, 16 * (size_t)n_sc, st));

    CK(h->plan.ensure(sizeof(ScPlan) * (size_t)n_sc));
    CK(h->list.ensure(sizeof(int) * (size_t)n_sc));
    CK(h->counters.ensure(sizeof(PlanCounters)));
    CK(cudaMemsetAsync(h->counters.p, 0, sizeof(PlanCounters), st));
    ScPlan *plan = (ScPlan *)h->plan.p;
    int *list = (int *)h->list.p;

    plan_kernel<<<(n_sc + 255) / 256, 256, 0, st>>>(in, plan, list, (PlanCounters *)h->counters.p,
                                                    h->force_class, kBigClass);
    S.n_launches++;
    CK(cudaMemcpyAsync(h->h_counters, h->counters.p, sizeof(PlanCounters), cudaMemcpyDeviceToHost, st));
    CK(cudaEventRecord(h->ev[1], st));

    // ---- short superclusters: one fused launch over the whole batch ----
    {
        constexpr int SPB = TINY_TPB / 4;
        const int smem = SPB * TINY_SC_BYTES + TINY_TPB * TINY_CAP;
        tiny_kernel<<<(n_sc + SPB - 1) / SPB, TINY_TPB, smem, st>>>(in, out, plan);
        S.n_launches++;
    }
    CK(cudaEventRecord(h->ev[2], st));
    CK(cudaStreamSynchronize(st));          // counters are now on the host
    CK(cudaGetLastError());
    const PlanCounters pc = *h->h_counters;
    S.cells = (int64_t)pc.cells;
    S.n_long = 4 * (int64_t)pc.n_list;
    S.n_short = 4 * (int64_t)(n_sc - pc.n_list - pc.n_bad);

    // ---- the rest: HBM slab, wavefront / scalar kernels ----
    float ms_fwd = 0, ms_bwd = 0, ms_walk = 0;
    if (pc.n_list > 0) {
        const int n = pc.n_list;
        CK(h->bytes.ensure(8 * (size_t)(n + 1)));
        CK(h->offs.ensure(8 * (size_t)(n + 1)));
        int64_t *bytes = (int64_t *)h->bytes.p, *offs = (int64_t *)h->offs.p;
        CK(cudaMemsetAsync(bytes + n, 0, 8, st));
        wave_size_kernel<<<(n + 255) / 256, 256, 0, st>>>(plan, list, n, bytes);
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
            slab_setup_kernel<<<(4 * m + 127) / 128, 128, 0, st>>>(in, plan, list, i0, i1, offs, (u8 *)h->slab.p,
                                                                   (int *)h->hap_ok.p);
            S.n_launches++;
            slab_align_kernel<<<(4 * m + 63) / 64, 64, 0, st>>>(in, out, plan, list, i0, i1, offs, (u8 *)h->slab.p,
                                                                (const int *)h->hap_ok.p, CLS_SCALAR);
            S.n_launches++;
            // ---- wavefront kernels: class-sorted items, then forward / backward / walk ----
            wave_tables_kernel<<<(4 * m + 127) / 128, 128, 0, st>>>(plan, list, i0, i1, offs, (u8 *)h->slab.p,
                                                                    (const int *)h->hap_ok.p);
            CK(h->wave_desc.ensure(sizeof(WaveItems) + 4 * (size_t)(4 * m + 4)));
            WaveItems *wi = (WaveItems *)h->wave_desc.p;
            int *items = (int *)((u8 *)h->wave_desc.p + sizeof(WaveItems));
            CK(cudaMemsetAsync(wi, 0, sizeof(WaveItems), st));
            wave_count_kernel<<<(4 * m + 127) / 128, 128, 0, st>>>(plan, list, i0, i1, (const int *)h->hap_ok.p, wi);
            S.n_launches += 2;
            WaveItems hwi;
            CK(cudaMemcpyAsync(&hwi, wi, sizeof(WaveItems), cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            ClsBase cb;
            int total_items = 0;
            for (int c = 0; c < N_WCLS; c++) { cb.b[c] = total_items; total_items += hwi.count[c]; }
            if (total_items > 0 || hwi.n_toolarge > 0) {
                wave_fill_kernel<<<(4 * m + 127) / 128, 128, 0, st>>>(plan, list, i0, i1, (const int *)h->hap_ok.p,
                                                                      wi, cb, items, out);
                S.n_launches++;
            }
            if (total_items > 0) {
                WaveArgs WA{in, out, plan, list, i0, offs, (u8 *)h->slab.p, items};
                // forward then backward of each class on its own stream (an alignment's backward pass
                // only depends on its own forward pass), all classes concurrently; then join
                CK(cudaEventRecord(h->ev[4], st));
                for (int c = 0; c < N_WCLS; c++) {
                    if (!hwi.count[c]) continue;
                    cudaStream_t ss = h->side[c];
                    CK(cudaStreamWaitEvent(ss, h->ev[4], 0));
                    CK(cudaEventRecord(h->sev[c][0], ss));
                    wave_launch(ss, WA, c, cb.b[c], hwi.count[c], true);
                    CK(cudaEventRecord(h->sev[c][1], ss));
                    wave_launch(ss, WA, c, cb.b[c], hwi.count[c], false);
                    CK(cudaEventRecord(h->sev[c][2], ss));
                    CK(cudaStreamWaitEvent(st, h->sev[c][2], 0));
                    S.n_launches += 2;
                }
                CK(cudaEventRecord(h->ev[6], st));
                wave_walk_kernel<<<(total_items + 63) / 64, 64, 0, st>>>(WA, total_items);
                S.n_launches++;
                CK(cudaEventRecord(h->ev[7], st));
                CK(cudaStreamSynchronize(st));
                CK(cudaGetLastError());
                // kernel durations: summed over classes (they overlap in wall time)
                for (int c = 0; c < N_WCLS; c++) {
                    if (!hwi.count[c]) continue;
                    float a_ = 0, b_ = 0;
                    cudaEventElapsedTime(&a_, h->sev[c][0], h->sev[c][1]);
                    cudaEventElapsedTime(&b_, h->sev[c][1], h->sev[c][2]);
                    ms_fwd += a_; ms_bwd += b_;
                }
                float c_ = 0, w_ = 0;
                cudaEventElapsedTime(&c_, h->ev[6], h->ev[7]);
                cudaEventElapsedTime(&w_, h->ev[4], h->ev[6]);
                ms_walk += c_;
                S.ms_long_wall += w_;
                S.spill_bytes += 3 * (int64_t)hwi.spill_cells;
            }
            if (hwi.n_toolarge > 0)
                return fail(h, VD_E_TOOLARGE, "%d alignments exceed the supported matrix side (32768 rows)", hwi.n_toolarge);
            i0 = i1;
        }
    }
    CK(cudaEventRecord(h->ev[3], st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    cudaEventElapsedTime(&S.ms_plan, h->ev[0], h->ev[1]);
    cudaEventElapsedTime(&S.ms_short, h->ev[1], h->ev[2]);
    cudaEventElapsedTime(&S.ms_total, h->ev[0], h->ev[3]);
    S.ms_long_fwd = ms_fwd; S.ms_long_bwd = ms_bwd; S.ms_long_walk = ms_walk;
    if (pc.n_bad > 0) return fail(h, VD_E_BADINPUT, "%d malformed superclusters", pc.n_bad);
    return VD_OK;
}

static int64_t io_bytes_of(int64_t n_sc, int64_t n_var, int64_t ref_bytes, int64_t alt_bytes) {
    const int64_t inp = ref_bytes + 8 * (n_sc + 1) + 8 * (4 * n_sc + 1) + n_var * (4 + 4 + 1 + 4) + 8 * (n_var + 1) + alt_bytes;
    const int64_t outb = 4 * n_sc * (4 + 1 + 1 + 4) + 2 * n_var * (1 + 4 + 4 + 4 + 4);
    return inp + outb;
}

extern "C" int vd_run_device(vd_handle *h, const vd_batch_in *in, vd_batch_out *out,
                             int64_t n_var, int64_t ref_bytes, int64_t alt_bytes) {
    if (!h || !in || !out) return VD_E_BADINPUT;
    CK(cudaSetDevice(h->device));
    BatchDev b{in->n_sc, in->ref_off, in->ref_seq, in->rplane_seq ? in->rplane_seq : in->ref_seq, in->var_off,
               in->var_pos, in->var_rlen, in->var_type, in->alt_off, in->alt_seq, in->var_qual, in->max_qual, n_var};
    OutDev o{out->aln_score, out->aln_end_plane, out->aln_beg_plane, out->status, out->assigned,
             out->sync_group, out->ref_ed, out->query_ed, out->callq};
    h->stats.h2d_bytes = h->stats.d2h_bytes = 0;
    h->stats.io_bytes = io_bytes_of(in->n_sc, n_var, ref_bytes, alt_bytes);
    int rc = run_resident(h, b, o);
    if (rc != VD_OK) return rc;
    return VD_OK;
}

extern "C" int vd_run(vd_handle *h, const vd_batch_in *in, vd_batch_out *out) {
    if (!h || !in || !out) return VD_E_BADINPUT;
    CK(cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    const int64_t n_sc = in->n_sc;
    if (n_sc < 0) return fail(h, VD_E_BADINPUT, "negative n_sc");
    if (n_sc == 0) { h->stats = vd_stats{}; return VD_OK; }
    const int64_t n_var = in->var_off[4 * n_sc];
    const int64_t ref_bytes = in->ref_off[n_sc];
    const int64_t alt_bytes = n_var ? in->alt_off[n_var] : 0;
    if (n_var < 0 || ref_bytes < 0 || alt_bytes < 0) return fail(h, VD_E_BADINPUT, "negative sizes");

    int64_t h2d = 0;
#define UP(buf, src, bytes) do { CK(h->buf.ensure((size_t)(bytes) + 16)); \
        if ((bytes) > 0) CK(cudaMemcpyAsync(h->buf.p, (src), (size_t)(bytes), cudaMemcpyHostToDevice, st)); \
        h2d += (bytes); } while (0)
    UP(in_ref_off, in->ref_off, 8 * (n_sc + 1));
    UP(in_ref_seq, in->ref_seq, ref_bytes);
    if (in->rplane_seq) UP(in_rplane, in->rplane_seq, ref_bytes);
    UP(in_var_off, in->var_off, 8 * (4 * n_sc + 1));
    UP(in_var_pos, in->var_pos, 4 * n_var);
    UP(in_var_rlen, in->var_rlen, 4 * n_var);
    UP(in_var_type, in->var_type, n_var);
    if (n_var) UP(in_alt_off, in->alt_off, 8 * (n_var + 1));
    else { CK(h->in_alt_off.ensure(16)); CK(cudaMemsetAsync(h->in_alt_off.p, 0, 16, st)); }
    UP(in_alt_seq, in->alt_seq, alt_bytes);
    UP(in_var_qual, in->var_qual, 4 * n_var);
#undef UP
    CK(h->o_score.ensure(16 * (size_t)n_sc)); CK(h->o_endp.ensure(4 * (size_t)n_sc)); CK(h->o_begp.ensure(4 * (size_t)n_sc));
    CK(h->o_status.ensure(16 * (size_t)n_sc)); CK(h->o_assigned.ensure(2 * (size_t)n_var + 16));
    CK(h->o_sg.ensure(8 * (size_t)n_var + 16)); CK(h->o_red.ensure(8 * (size_t)n_var + 16));
    CK(h->o_qed.ensure(8 * (size_t)n_var + 16)); CK(h->o_callq.ensure(8 * (size_t)n_var + 16));

    BatchDev b{(int)n_sc, (const int64_t *)h->in_ref_off.p, (const u8 *)h->in_ref_seq.p,
               (const u8 *)(in->rplane_seq ? h->in_rplane.p : h->in_ref_seq.p), (const int64_t *)h->in_var_off.p,
               (const int32_t *)h->in_var_pos.p, (const int32_t *)h->in_var_rlen.p, (const u8 *)h->in_var_type.p,
               (const int64_t *)h->in_alt_off.p, (const u8 *)h->in_alt_seq.p, (const float *)h->in_var_qual.p,
               in->max_qual, n_var};
    OutDev o{(int32_t *)h->o_score.p, (u8 *)h->o_endp.p, (u8 *)h->o_begp.p, (u32 *)h->o_status.p,
             (u8 *)h->o_assigned.p, (int32_t *)h->o_sg.p, (int32_t *)h->o_red.p, (int32_t *)h->o_qed.p,
             (float *)h->o_callq.p};
    h->stats.io_bytes = io_bytes_of(n_sc, n_var, ref_bytes, alt_bytes);
    int rc = run_resident(h, b, o);
    if (rc != VD_OK && rc != VD_E_BADINPUT) return rc;

    int64_t d2h = 0;
#define DOWN(dst, buf, bytes) do { if ((bytes) > 0) CK(cudaMemcpyAsync((dst), h->buf.p, (size_t)(bytes), cudaMemcpyDeviceToHost, st)); \
        d2h += (bytes); } while (0)
    DOWN(out->aln_score, o_score, 16 * n_sc);
    DOWN(out->aln_end_plane, o_endp, 4 * n_sc);
    DOWN(out->aln_beg_plane, o_begp, 4 * n_sc);
    DOWN(out->status, o_status, 16 * n_sc);
    DOWN(out->assigned, o_assigned, 2 * n_var);
    DOWN(out->sync_group, o_sg, 8 * n_var);
    DOWN(out->ref_ed, o_red, 8 * n_var);
    DOWN(out->query_ed, o_qed, 8 * n_var);
    DOWN(out->callq, o_callq, 8 * n_var);
#undef DOWN
    CK(cudaStreamSynchronize(st));
    h->stats.h2d_bytes = h2d;
    h->stats.d2h_bytes = d2h;
    if (rc != VD_OK) return rc;
    // fatal reference conditions are reported, results stay available for inspection
    for (int64_t i = 0; i < 4 * n_sc; i++)
        if (out->status[i] & VD_ST_ERR_MASK)
            return fail(h, VD_E_ALIGN, "alignment %lld of supercluster %lld: status 0x%x", (long long)(i & 3),
                        (long long)(i >> 2), out->status[i]);
    return VD_OK;
}
