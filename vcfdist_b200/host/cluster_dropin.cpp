// Drop-in replacement for the reference's cluster-growing stage
//
//     void wf_swg_cluster(variantData * vcf, int ctg_idx, int hap, int sub, int open, int extend);
//                                     (decl src/cluster.h:68-69, def src/cluster.cpp:954-1263)
//
// Same name, arguments and in-place result convention (ctgVariants::clusters / left_reaches / right_reaches).  The
// reference grows the clusters of one haplotype of one contig by running, per cluster and iteration, one affine-gap
// alignment (wf_swg_align, src/dist.cpp:1510-1652, for the score) and two reach searches with window doubling
// (wf_swg_max_reach, :2150-2333) one after the other on the calling thread.  Here every round is ONE batch of the
// hand-written wavefront kernel (vd_wf_batch, csrc/vd_reach.cuh: a warp per problem): the scores of all active clusters,
// then per doubling round the searches that still hit the far end of their window.  The merge passes and the doubling
// control (:1171-1238) stay on the host: O(#clusters) integer work.  vcfdist_b200/cluster.py is the same driver in
// Python (tests/test_gpu_cluster.py pins it against the reference's recorded answers).
//
// The reference calls this function from up to -t host threads at once (src/main.cpp:82, :119, :146, :193), one per
// (contig, haplotype): every caller takes a handle of its own from the runtime's pool.
//
// Clustering is the first stage after the VCFs are read, so on a small input it starts before CUDA is up (0.5-4 s on these
// boxes) and the first calls wait for it.  VD_GPU_CLUSTER_NOWAIT=1: a call that arrives while the runtime is still starting
// runs the reference's own definition (linked in under the name ref_wf_swg_cluster) instead of waiting - same results, that
// is what the parity tests check; VD_GPU_CLUSTER=0 keeps the stage on the CPU altogether.  Neither is a fallback of the
// precision/recall path, which has none.
#include <algorithm>
#include <chrono>
#include <climits>
#include <cstdint>
#include <cstdlib>
#include <memory>
#include <string>
#include <vector>

#include "globals.h"
#include "variant.h"
#include "cluster.h"

#include "vcfdist_b200.h"
#include "dropin_runtime.h"

namespace {

// generate_str (src/dist.cpp:81-136, min_qual 0): the window [beg_pos, end_pos) of the contig with variants
// beg_idx .. end_idx-1 applied
void apply_variants(const std::string &fa, const ctgVariants &V, int beg_idx, int end_idx, int beg_pos, int end_pos,
                    std::string &out) {
    int vi = beg_idx;
    while (vi < V.n && V.poss[vi] < beg_pos) vi++;
    for (int ref_pos = beg_pos; ref_pos < end_pos;) {
        if (vi < end_idx && ref_pos == V.poss[vi]) {
            switch (V.types[vi]) {
                case TYPE_INS: out += V.alts[vi]; break;
                case TYPE_DEL: ref_pos += (int)V.refs[vi].size(); break;
                case TYPE_SUB: out += V.alts[vi]; ref_pos++; break;
                case TYPE_CPX: out += V.alts[vi]; ref_pos += (int)V.refs[vi].size(); break;
            }
            vi++;
        } else {
            const int ref_end = vi < end_idx ? std::min(end_pos, V.poss[vi]) : end_pos;
            if (ref_end < ref_pos) ERROR("No variant, but ref_end < ref_pos (generate_str)");          // src/dist.cpp:125
            if (ref_pos > (int)fa.size()) ERROR("Contig position out of range (generate_str)");         // :130
            out.append(fa, (size_t)ref_pos, (size_t)(ref_end - ref_pos));
            ref_pos = ref_end;
        }
    }
}

struct Times { double gpu_ms = 0, wait_ms = 0; long batches = 0, problems = 0, bytes = 0; };

struct Batch {                       // problems of one vd_wf_batch call
    std::vector<int64_t> q_off{0}, t_off{0};
    std::string q, t;
    std::vector<int32_t> main_diag, main_diag_start, max_score, result;
    std::vector<uint8_t> reverse;
    int n() const { return (int)q_off.size() - 1; }
    void close() { q_off.push_back((int64_t)q.size()); t_off.push_back((int64_t)t.size()); }
    void run(int mode, int sub, int open, int extend, Times &tm) {
        using clk = std::chrono::steady_clock;
        result.assign((size_t)std::max(n(), 1), 0);
        if (!n()) return;
        const auto t0 = clk::now();
        vdhost::Runtime &rt = vdhost::runtime();
        vdhost::Runtime::Lease lease(rt);                  // this thread's handle for the call
        vd_handle *h = lease.h;
        if (!h) ERROR("vcfdist_b200: cannot initialise CUDA device %d (code %d); there is no CPU fallback for the clustering path",
                      rt.device, rt.rc);
        const auto t1 = clk::now();
        const bool reach = mode == 0;
        const int rc = vd_wf_batch(h, mode, n(), q_off.data(), (const uint8_t *)q.data(), t_off.data(), (const uint8_t *)t.data(),
                                   reach ? main_diag.data() : nullptr, reach ? main_diag_start.data() : nullptr,
                                   reach ? max_score.data() : nullptr, reach ? reverse.data() : nullptr, sub, open, extend, result.data());
        if (rc != VD_OK) ERROR("vcfdist_b200: vd_wf_batch failed (code %d): %s", rc, vd_last_error(h));
        tm.wait_ms += std::chrono::duration<double, std::milli>(t1 - t0).count();
        tm.gpu_ms += std::chrono::duration<double, std::milli>(clk::now() - t1).count();
        tm.batches++; tm.problems += n(); tm.bytes += (long)(q.size() + t.size());
    }
};

struct Search {                      // one reach search of one cluster (left or right) across its doubling rounds
    int c, first, last, beg_pos, end_pos, ref_len, main_diag, main_diag_start, score, reach;
    bool left;
};

}  // namespace

void ref_wf_swg_cluster(variantData *vcf, int ctg_idx, int hap, int sub, int open, int extend);   // the reference's, renamed at build time

void wf_swg_cluster(variantData *vcf, int ctg_idx, int hap, int sub, int open, int extend) {
    {
        const char *on = std::getenv("VD_GPU_CLUSTER"), *nowait = std::getenv("VD_GPU_CLUSTER_NOWAIT");
        if ((on && !std::atoi(on)) || (nowait && std::atoi(nowait) && !vdhost::runtime().reached(1)))
            return ref_wf_swg_cluster(vcf, ctg_idx, hap, sub, open, extend);
    }
    const std::string ctg = vcf->contigs[ctg_idx];
    std::shared_ptr<ctgVariants> vars = vcf->variants[hap][ctg];
    if (!vars->n) return;                                                                  // :964
    const ctgVariants &V = *vars;
    const std::string &fa = vcf->ref->fasta.at(ctg);
    const int L = vcf->lengths[ctg_idx];
    const int n = V.n;

    Times tm;
    const auto t_begin = std::chrono::steady_clock::now();
    std::vector<int> prev_clusters(n + 1);
    for (int i = 0; i <= n; i++) prev_clusters[i] = i;
    std::vector<char> prev_active(n + 1, 1);
    std::vector<int> right_reach(n + 1), left_reach(n + 1);
    int iter = 0;
    while (std::find(prev_active.begin(), prev_active.end(), (char)1) != prev_active.end()) {   // :979-981
        if (++iter > g.max_cluster_itrs) break;
        const std::vector<int> &cl = prev_clusters;
        const int nc = (int)cl.size();
        left_reach[nc - 1] = right_reach[nc - 1] = INT_MAX;                                 // sentinels, :995-996
        // ---- alignment score of every active cluster against the reference (:1037-1046): one batch ----
        std::vector<int> act;
        for (int c = 0; c < nc - 1; c++) if (prev_active[c]) act.push_back(c);              // :1002-1010
        Batch sb;
        for (int c : act) {
            const int first = cl[c], last = cl[c + 1] - 1;
            const int beg = std::max(0, V.poss[first] - 1);
            const int end = std::min(L, V.poss[last] + V.rlens[last] + 1);
            apply_variants(fa, V, first, last + 1, beg, end, sb.q);
            sb.t.append(fa, (size_t)beg, (size_t)(end - beg));
            sb.close();
        }
        sb.run(/*score*/1, sub, open, extend, tm);
        // ---- reaches: every (cluster, direction) search, one batch per doubling round (:1049-1158) ----
        std::vector<Search> searches;
        searches.reserve(2 * act.size());
        for (size_t k = 0; k < act.size(); k++) {
            const int c = act[k], first = cl[c], last = cl[c + 1] - 1;
            int main_diag = 0;
            for (int v = first; v <= last; v++) main_diag += (int)V.refs[v].size() - (int)V.alts[v].size();   // :1062-1064
            for (int left = 1; left >= 0; left--) {
                Search s;
                s.c = c; s.left = left != 0; s.first = first; s.last = last; s.score = sb.result[k]; s.main_diag = main_diag;
                s.beg_pos = V.poss[first] - 1;
                s.end_pos = V.poss[last] + V.rlens[last] + 1;
                s.main_diag_start = s.left ? s.end_pos - V.poss[first] : V.poss[last] + V.rlens[last] - s.beg_pos;
                s.ref_len = s.end_pos - s.beg_pos;
                s.reach = s.ref_len - 1;
                searches.push_back(s);
            }
        }
        std::vector<int> pending(searches.size());
        for (size_t i = 0; i < pending.size(); i++) pending[i] = (int)i;
        std::string q, r;
        while (!pending.empty()) {
            Batch rb;
            for (int i : pending) {                                                         // window doubling, :1066-1077 / :1127-1137
                Search &s = searches[i];
                s.ref_len *= 2;
                const int slack = std::abs(s.main_diag) + s.score / extend + 3;
                q.clear(); r.clear();
                if (s.left) {
                    s.beg_pos = std::max(0, s.end_pos - s.ref_len - slack);
                    apply_variants(fa, V, s.first, s.last + 1, s.beg_pos, s.end_pos, q);
                    const int start = std::max(0, s.end_pos - s.ref_len);
                    r.assign(fa, (size_t)start, (size_t)s.ref_len);                         // substr clips at the contig end
                    std::reverse(q.begin(), q.end());
                    std::reverse(r.begin(), r.end());
                } else {
                    s.end_pos = std::min(L, s.beg_pos + s.ref_len + slack);
                    apply_variants(fa, V, s.first, s.last + 1, s.beg_pos, s.end_pos, q);
                    r.assign(fa, (size_t)s.beg_pos, (size_t)std::min(s.ref_len, s.end_pos - s.beg_pos));
                }
                rb.q += q; rb.t += r; rb.close();
                rb.main_diag.push_back(s.main_diag); rb.main_diag_start.push_back(s.main_diag_start);
                rb.max_score.push_back(s.score); rb.reverse.push_back(s.left ? 1 : 0);
            }
            rb.run(/*reach*/0, sub, open, extend, tm);
            std::vector<int> next;
            for (size_t k = 0; k < pending.size(); k++) {
                Search &s = searches[pending[k]];
                s.reach = rb.result[k];
                const bool hit_edge = s.left ? s.beg_pos == 0 : s.end_pos == L;             // :1094, :1155
                if (s.reach == s.ref_len - 1 && !hit_edge) next.push_back(pending[k]);
            }
            pending.swap(next);
        }
        for (const Search &s : searches) {
            if (s.left) left_reach[s.c] = s.end_pos - s.reach;                              // :1097
            else right_reach[s.c] = s.beg_pos + s.reach + 1;                                // :1158
        }
        // ---- merge rightwards, :1171-1192 ----
        std::vector<int> tmp_clusters, tmp_left, tmp_right;
        std::vector<char> tmp_active;
        for (int c = 0; c < nc;) {
            int size = 1, max_r = right_reach[c], min_l = left_reach[c];
            while (c + size < nc && (long long)max_r + g.reach_min_gap >= left_reach[c + size]) {
                max_r = std::max(max_r, right_reach[c + size]);
                min_l = std::min(min_l, left_reach[c + size]);
                size++;
            }
            tmp_right.push_back(max_r); tmp_left.push_back(min_l);
            tmp_clusters.push_back(prev_clusters[c]); tmp_active.push_back(size > 1);
            c += size;
        }
        // ---- merge leftwards, :1208-1231 ----
        std::vector<int> next_clusters, nl, nr;
        std::vector<char> next_active;
        for (int c = (int)tmp_clusters.size() - 1; c >= 0; c--) {
            int min_l = tmp_left[c], max_r = tmp_right[c];
            char active = tmp_active[c];
            while (c > 0 && min_l <= (long long)tmp_right[c - 1] + g.reach_min_gap) {
                min_l = std::min(min_l, tmp_left[c - 1]);
                max_r = std::max(max_r, tmp_right[c - 1]);
                active = 1;
                c--;
            }
            nl.push_back(min_l); nr.push_back(max_r);
            next_clusters.push_back(tmp_clusters[c]); next_active.push_back(active);
        }
        std::reverse(next_clusters.begin(), next_clusters.end());
        std::reverse(next_active.begin(), next_active.end());
        std::reverse(nl.begin(), nl.end());
        std::reverse(nr.begin(), nr.end());
        left_reach.swap(nl);                                                                // one entry per cluster from here on (:1196-1237)
        right_reach.swap(nr);
        prev_clusters.swap(next_clusters);
        prev_active.swap(next_active);
    }
    if (g.verbosity >= 2 || std::getenv("VD_DROPIN_TIMES"))
        INFO("  GPU clustering %s hap %d: %d variants -> %d clusters, %d iterations, %ld batches, %ld problems, %.1f MB of strings; "
             "%.0f ms in all, %.0f ms in vd_wf_batch, %.0f ms waiting for the handle", ctg.data(), hap + 1, n, (int)prev_clusters.size() - 1,
             iter, tm.batches, tm.problems, tm.bytes / 1e6,
             std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count(), tm.gpu_ms, tm.wait_ms);
    vars->clusters = prev_clusters;                                                          // :1259-1262
    vars->left_reaches = left_reach;
    vars->right_reaches = right_reach;
}
