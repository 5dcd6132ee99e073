// Drop-in replacement for the reference's `--distance` pass
//
//     editData edits_wrapper(std::shared_ptr<superclusterData> clusterdata_ptr);
//                                     (decl src/dist.h:244, def src/dist.cpp:1908-2077)
//
// Same name, argument, return value and log lines.  The reference walks contigs, superclusters, haplotypes and quality
// thresholds on one thread and runs, for each, one affine-gap alignment (wf_swg_align, src/dist.cpp:1510-1652) and its walk
// back (wf_swg_backtrack, :2625-2757).  Here the strings of a contig are gathered in the reference's order (by the
// reference's own generate_ptrs_strs), aligned as batches of the hand-written wavefront kernels (vd_swg_align_batch,
// csrc/vd_reach.cuh: a warp per problem, score pass then alignment pass with the CIGAR written as the reference fills
// it), and the CIGARs go through the reference's own count_dist / editData::add_edits in the same order.
#include <algorithm>
#include <cstdint>
#include <memory>
#include <set>
#include <string>
#include <vector>

#include "globals.h"
#include "variant.h"
#include "cluster.h"
#include "dist.h"
#include "edit.h"

#include "vcfdist_b200.h"
#include "dropin_runtime.h"

namespace {

struct Problem { int sc_idx, beg, hap, prev_qual, qual; };

struct EditBatch {
    std::vector<int64_t> q_off{0}, t_off{0};
    std::string q, t;
    std::vector<Problem> prob;
    int64_t bytes() const { return (int64_t)(q.size() + t.size()); }
    void clear() { q_off.assign(1, 0); t_off.assign(1, 0); q.clear(); t.clear(); prob.clear(); }
};

// aligns the batch and feeds the CIGARs to the reference's bookkeeping, in order (src/dist.cpp:2042-2061)
void flush(EditBatch &b, const std::string &ctg, editData &edits, std::vector<int> &all_qual_dists, std::vector<int> &ctg_qual_dists) {
    const int n = (int)b.prob.size();
    if (!n) return;
    std::vector<int32_t> score((size_t)n), cigar((size_t)b.bytes() + 1);
    {
        vdhost::Runtime &rt = vdhost::runtime();
        vdhost::Runtime::Lease lease(rt);
        if (!lease.h) ERROR("vcfdist_b200: cannot initialise CUDA device %d (code %d); there is no CPU fallback for the --distance path",
                            rt.device, rt.rc);
        const int rc = vd_swg_align_batch(lease.h, n, b.q_off.data(), (const uint8_t *)b.q.data(), b.t_off.data(), (const uint8_t *)b.t.data(),
                                          g.eval_sub, g.eval_open, g.eval_extend, score.data(), cigar.data());
        if (rc != VD_OK) ERROR("vcfdist_b200: vd_swg_align_batch failed (code %d): %s", rc, vd_last_error(lease.h));
    }
    for (int i = 0; i < n; i++) {
        const Problem &p = b.prob[i];
        if (score[i] < 0) ERROR("Unexpected pointer during WFA backtrack (supercluster %d on contig '%s')", p.sc_idx, ctg.data());
        const int64_t o = b.q_off[i] + b.t_off[i], len = (b.q_off[i + 1] - b.q_off[i]) + (b.t_off[i + 1] - b.t_off[i]);
        std::vector<int> cig(cigar.begin() + o, cigar.begin() + o + len);
        std::reverse(cig.begin(), cig.end());                                              // :2050
        const int dist = count_dist(cig);                                                  // :2051
        for (int q = p.prev_qual; q < p.qual; q++) { all_qual_dists[q] += dist; ctg_qual_dists[q] += dist; }   // :2055-2059
        edits.add_edits(ctg, p.beg, (uint8_t)p.hap, cig, p.sc_idx, p.prev_qual, p.qual);    // :2060
    }
    b.clear();
}

}  // namespace

editData edits_wrapper(std::shared_ptr<superclusterData> clusterdata_ptr) {
    if (g.verbosity >= 1) INFO(" ");
    if (g.verbosity >= 1) INFO("%s[6/8] Calculating edit distance metrics%s", COLOR_PURPLE, COLOR_WHITE);
    std::vector<int> all_qual_dists(g.max_qual + 2, 0);                                     // :1915
    editData edits;
    int ctg_id = 0;
    if (g.verbosity >= 1) INFO("  Contigs:");
    int64_t flush_bytes = 64 << 20;
    if (const char *v = std::getenv("VD_EDITS_BATCH_MB")) flush_bytes = std::max<int64_t>(1, std::atoll(v)) << 20;
    for (std::string ctg : clusterdata_ptr->contigs) {
        std::vector<int> ctg_qual_dists(g.max_qual + 2, 0);
        std::shared_ptr<ctgSuperclusters> sc = clusterdata_ptr->superclusters[ctg];
        EditBatch batch;
        for (int sc_idx = 0; sc_idx < sc->n; sc_idx++) {
            // truth haplotype strings (:1962-1981)
            std::string truth1 = "", ref_t1 = "", truth2 = "", ref_t2 = "";
            std::vector<std::vector<int>> t1r, rt1, t2r, rt2;
            generate_ptrs_strs(truth1, ref_t1, t1r, rt1, sc->ctg_variants[TRUTH][HAP1], sc->superclusters[TRUTH][HAP1][sc_idx],
                               sc->superclusters[TRUTH][HAP1][sc_idx + 1], sc->begs[sc_idx], sc->ends[sc_idx], clusterdata_ptr->ref, ctg);
            generate_ptrs_strs(truth2, ref_t2, t2r, rt2, sc->ctg_variants[TRUTH][HAP2], sc->superclusters[TRUTH][HAP2][sc_idx],
                               sc->superclusters[TRUTH][HAP2][sc_idx + 1], sc->begs[sc_idx], sc->ends[sc_idx], clusterdata_ptr->ref, ctg);
            const int phase = sc->sc_phase[sc_idx];
            if (phase < 0) ERROR("Phase never set for supercluster %d on contig '%s'", sc_idx, ctg.data());   // :1992-1994
            std::vector<std::string> truth(2);
            if (phase == PHASE_SWAP) { truth[HAP1] = truth2; truth[HAP2] = truth1; }        // :1995-1999
            else { truth[HAP1] = truth1; truth[HAP2] = truth2; }
            for (int hap = 0; hap < HAPS; hap++) {
                std::shared_ptr<ctgVariants> qv = sc->ctg_variants[QUERY][hap];
                const int beg_idx = qv->clusters.size() ? qv->clusters[sc->superclusters[QUERY][hap][sc_idx]] : 0;        // :2007-2012
                const int end_idx = qv->clusters.size() ? qv->clusters[sc->superclusters[QUERY][hap][sc_idx + 1]] : 0;
                std::set<int> quals = {};                                                   // :2016-2020
                for (int v = beg_idx; v < end_idx; v++) quals.insert(qv->var_quals[v] + 1);
                quals.insert(g.max_qual + 2);
                std::string rtruth(truth[hap].rbegin(), truth[hap].rend());
                int prev_qual = 0;
                for (int qual : quals) {                                                    // :2024
                    std::string query, ref;
                    std::vector<std::vector<int>> qrp, rqp;
                    generate_ptrs_strs(query, ref, qrp, rqp, qv, sc->superclusters[QUERY][hap][sc_idx],
                                       sc->superclusters[QUERY][hap][sc_idx + 1], sc->begs[sc_idx], sc->ends[sc_idx],
                                       clusterdata_ptr->ref, ctg, prev_qual);
                    batch.q.append(query.rbegin(), query.rend());                           // both strings reversed, :2045-2046
                    batch.t += rtruth;
                    batch.q_off.push_back((int64_t)batch.q.size());
                    batch.t_off.push_back((int64_t)batch.t.size());
                    batch.prob.push_back({sc_idx, sc->begs[sc_idx], hap, prev_qual, qual});
                    prev_qual = qual;
                }
            }
            if (batch.bytes() >= flush_bytes) flush(batch, ctg, edits, all_qual_dists, ctg_qual_dists);
        }
        flush(batch, ctg, edits, all_qual_dists, ctg_qual_dists);
        if (g.verbosity >= 1)
            INFO("    [%2d] %s: %d", ctg_id, ctg.data(), *std::min_element(ctg_qual_dists.begin(), ctg_qual_dists.end()));
        ctg_id++;
    }
    INFO(" ");
    if (g.verbosity >= 1) INFO("  Total edit distance: %d", *std::min_element(all_qual_dists.begin(), all_qual_dists.end()));
    return edits;
}
