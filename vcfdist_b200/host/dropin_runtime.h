// One GPU handle and one page-locked arena per process, shared by the drop-ins (pr_dropin.cpp, cluster_dropin.cpp): a
// background thread started when the first of them is loaded creates the handle, page-locks the staging arena and runs a
// small warm-up batch, so that CUDA start-up overlaps the reference's VCF parsing.
#pragma once
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <mutex>
#include <thread>
#include <vector>

#include "vcfdist_b200.h"

namespace vdhost {

// One GPU handle and one page-locked arena per process, set up by a background thread at load time.
struct Runtime {
    std::thread th;
    vd_handle *h = nullptr;
    int rc = VD_OK, device = 0;
    uint8_t *arena = nullptr;
    int64_t arena_cap = 0;
    double init_ms[3] = {0, 0, 0};              // vd_create, page-locking the arena, warm-up batch
    // start-up stages: 1 = the handle exists (or could not be created), 2 = arena page-locked and warm-up batch done.
    // The clustering stage only needs the first; the precision/recall stage, which uses the arena, waits for the second.
    std::mutex st_mu;
    std::condition_variable st_cv;
    int stage = 0;
    bool started = false;
    void set_stage(int s) { { std::lock_guard<std::mutex> lk(st_mu); stage = s; } st_cv.notify_all(); }
    void wait_stage(int s) { std::unique_lock<std::mutex> lk(st_mu); st_cv.wait(lk, [&] { return stage >= s; }); }
    bool reached(int s) { std::lock_guard<std::mutex> lk(st_mu); return stage >= s; }
    void init() {
        using clk = std::chrono::steady_clock;
        auto ms_since = [](clk::time_point t) { return std::chrono::duration<double, std::milli>(clk::now() - t).count(); };
        auto t0 = clk::now();
        if (const char *d = std::getenv("VD_DEVICE")) device = std::atoi(d);
        rc = vd_create(device, 0, &h);
        if (rc != VD_OK) { set_stage(2); return; }
        init_ms[0] = ms_since(t0); t0 = clk::now();
        { std::lock_guard<std::mutex> lk(pool_mu); main_busy = true; }      // the warm-up batch below runs on the first handle
        set_stage(1);
        int64_t mb = 1024;
        if (const char *m = std::getenv("VD_PIN_MB")) mb = std::atoll(m);
        if (mb > 0) { arena = (uint8_t *)vd_host_alloc(mb << 20); arena_cap = arena ? (mb << 20) : 0; }
        init_ms[1] = ms_since(t0); t0 = clk::now();
        if (!std::getenv("VD_NO_WARMUP_BATCH")) warm_up();
        init_ms[2] = ms_since(t0);
        { std::lock_guard<std::mutex> lk(pool_mu); main_busy = false; }
        pool_cv.notify_all();
        set_stage(2);
    }
    // A few synthetic superclusters of every size class through the whole path, so that the kernels' code is on the
    // GPU and the handle's work buffers exist before the real batch arrives (CUDA loads a kernel at its first launch).
    void warm_up() {
        const int lens[] = {5, 5, 5, 5, 14, 14, 40, 40, 100, 700, 3000};
        std::vector<int64_t> ref_off{0}, var_off{0}, alt_off{0};
        std::vector<uint8_t> ref, alt, type;
        std::vector<int32_t> pos, rlen;
        std::vector<float> qual;
        unsigned x = 12345u;
        auto rnd = [&x]() { x = x * 1664525u + 1013904223u; return x >> 16; };
        for (int rep = 0; rep < 8; rep++)
            for (int L : lens) {
                const size_t r0 = ref.size();
                for (int k = 0; k < L; k++) ref.push_back("ACGT"[rnd() & 3]);
                ref_off.push_back((int64_t)ref.size());
                for (int hap = 0; hap < 4; hap++) {
                    // haplotype `hap` carries a SNP at 1 + hap (+ a second one further on in long windows): heterozygous,
                    // so that every alignment of the supercluster is computed
                    for (int at : {1 + hap, L > 30 ? L / 2 + hap : -1}) {
                        if (at < 0 || at >= L - 1) continue;
                        pos.push_back(at); rlen.push_back(1); type.push_back(VD_TYPE_SUB);
                        alt.push_back(ref[r0 + at] == 'A' ? 'C' : 'A');
                        alt_off.push_back((int64_t)alt.size());
                        qual.push_back(30.f);
                    }
                    var_off.push_back((int64_t)pos.size());
                }
            }
        const int64_t n_sc = (int64_t)ref_off.size() - 1, n_var = (int64_t)pos.size();
        vd_batch_in in{};
        in.n_sc = (int32_t)n_sc; in.ref_off = ref_off.data(); in.ref_seq = ref.data(); in.rplane_seq = nullptr;
        in.var_off = var_off.data(); in.var_pos = pos.data(); in.var_rlen = rlen.data(); in.var_type = type.data();
        in.alt_off = alt_off.data(); in.alt_seq = alt.data(); in.var_qual = qual.data(); in.max_qual = 60.f;
        std::vector<uint16_t> a16(3 * 4 * n_sc), v16(3 * 2 * n_var);
        std::vector<uint8_t> pl(4 * n_sc);
        std::vector<float> cq(2 * n_var);
        vd_packed_out pk{a16.data(), pl.data(), a16.data() + 4 * n_sc, v16.data(), v16.data() + 2 * n_var, v16.data() + 4 * n_var, cq.data()};
        vd_run_packed(h, &in, &pk);          // result and return code are of no interest
    }
    void start() {
        std::lock_guard<std::mutex> lk(st_mu);
        if (!started) { started = true; th = std::thread([this] { init(); }); }
    }
    Runtime() { if (!std::getenv("VD_NO_WARM")) start(); }
    // A handle serves one host thread at a time.  The reference calls its clustering stage from several threads at once
    // (one per contig and haplotype): each takes a handle of its own from a small pool (the first one plus up to
    // VD_HANDLES - 1 more, created on demand), so that their batches run side by side on the GPU.
    std::mutex pool_mu;
    std::condition_variable pool_cv;
    std::vector<vd_handle *> idle_extra;
    bool main_busy = false;
    int n_extra = 0, max_extra = 7;
    vd_handle *acquire() {
        start();
        wait_stage(1);
        vd_handle *first = h;
        if (!first) return nullptr;
        if (const char *m = std::getenv("VD_HANDLES")) max_extra = std::atoi(m) - 1;
        std::unique_lock<std::mutex> lk(pool_mu);
        for (;;) {
            if (!main_busy) { main_busy = true; return first; }
            if (!idle_extra.empty()) { vd_handle *x = idle_extra.back(); idle_extra.pop_back(); return x; }
            if (n_extra < max_extra) {
                n_extra++;
                lk.unlock();
                vd_handle *x = nullptr;
                if (vd_create(device, 0, &x) == VD_OK) return x;
                lk.lock();
                n_extra--; max_extra = n_extra;         // no more handles to be had: share the ones there are
                continue;
            }
            pool_cv.wait(lk);
        }
    }
    void release(vd_handle *x) {
        { std::lock_guard<std::mutex> lk(pool_mu); if (x == h) main_busy = false; else idle_extra.push_back(x); }
        pool_cv.notify_one();
    }
    struct Lease {
        Runtime &rt; vd_handle *h;
        explicit Lease(Runtime &r) : rt(r), h(r.acquire()) {}
        ~Lease() { if (h) rt.release(h); }
        Lease(const Lease &) = delete;
        Lease &operator=(const Lease &) = delete;
    };
    ~Runtime() {
        if (th.joinable()) th.join();
        if (arena) vd_host_free(arena);
        for (vd_handle *x : idle_extra) vd_destroy(x);
        if (h) vd_destroy(h);
    }
};

inline Runtime &runtime() { static Runtime rt; return rt; }
// one per translation unit that includes this header: the runtime starts at load time, not at the first call
static const int runtime_started_at_load = (runtime(), 0);

}  // namespace vdhost
