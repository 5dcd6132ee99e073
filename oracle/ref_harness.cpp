/* TEST INFRASTRUCTURE ONLY — function-level harness around the REFERENCE's own object
 * code.  Linked against dist.o / cluster.o / variant.o / ... compiled by oracle/Makefile
 * from the unmodified sources in /root/reference/src (outputs in oracle/_ref/ only).
 *
 * vdref_run() rebuilds the reference's `superclusterData` (src/cluster.h:45-62) from a
 * compact vd_batch_in, calls the reference's sort_superclusters (src/cluster.cpp:42-122)
 * and precision_recall_threads_wrapper (src/dist.cpp:1656-1727) — the exact seam the
 * product replaces — and copies the per-variant / per-supercluster results out.  It is
 * used (a) to pin oracle/vd_oracle.c, (b) to generate tests/golden/, and (c) as the
 * "reference" CPU baseline of bench.py, timed with the reference's own std::thread ladder.
 */
#include <chrono>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "globals.h"
#include "variant.h"
#include "cluster.h"
#include "dist.h"
#include "fasta.h"

#include "vcfdist_b200.h"

/* the reference defines these in src/main.cpp:13-26, which the harness replaces */
Globals g;
std::vector<std::string> type_strs = {"REF", "SNP", "INS", "DEL", "CPX"};
std::vector<std::string> type_strs2 = {"ALL", "SNP", "INS", "DEL", "INDEL"};
std::vector<std::string> vartype_strs = {"SNP", "INDEL", "SV", "ALL"};
std::vector<std::string> error_strs = {"TP", "FP", "FN", "PE", "GE", "??"};
std::vector<std::string> gt_strs = {"0", "1", "0|0", "0|1", "1|0", "1|1", "1|2", "2|1", ".|.", "M|N"};
std::vector<std::string> region_strs = {"OUTSIDE", "INSIDE ", "BORDER ", "OFF CTG"};
std::vector<std::string> aln_strs = {"QUERY1-TRUTH1", "QUERY1-TRUTH2", "QUERY2-TRUTH1", "QUERY2-TRUTH2"};
std::vector<std::string> callset_strs = {"QUERY", "TRUTH"};
std::vector<std::string> phase_strs = {"=", "X", "?"};
std::vector<std::string> switch_strs = {"FLIP", "SWITCH", "SWITCH+FLIP", "SWITCH_ERR", "FLIP_BEG", "FLIP_END", "NONE"};

extern "C" {

typedef struct vdref_out {
    uint8_t *errtypes;     /* [2*n_var]  ERRTYPE_* (src/defs.h:66-72), slot-major like vd_batch_out */
    int32_t *sync_group;   /* [2*n_var] */
    int32_t *ref_ed;       /* [2*n_var] */
    int32_t *query_ed;     /* [2*n_var] */
    float   *callq;        /* [2*n_var] */
    float   *credit;       /* [2*n_var] */
    int32_t *sc_phase;     /* [n_sc] */
    int32_t *orig_dist;    /* [n_sc] */
    int32_t *swap_dist;    /* [n_sc] */
} vdref_out;

/* Runs the reference's hot path on the batch.  Returns 0; *seconds = wall time of
 * precision_recall_threads_wrapper alone (what TIME_PR_ALN brackets, src/main.cpp:218-221). */
int vdref_run(const vd_batch_in *in, vdref_out *out, int threads, double max_ram,
              double phase_threshold, double credit_threshold, int max_qual,
              double *seconds) {
    const int n_sc = in->n_sc;
    const int64_t n_var = in->var_off[4 * (int64_t)n_sc];

    /* --- globals the path reads --- */
    g.verbosity = 0;
    g.max_qual = max_qual;
    g.phase_threshold = phase_threshold;
    g.credit_threshold = credit_threshold;
    g.max_threads = threads;
    g.max_ram = max_ram;
    g.thread_steps.clear();
    g.ram_steps.clear();
    g.thread_nsteps = 0;
    for (int t = threads; t > 0; t /= 2) {      /* halving ladder, as src/globals.cpp:486-494 */
        g.thread_steps.push_back(t);
        g.ram_steps.push_back(float(max_ram / t));
        g.thread_nsteps++;
    }

    /* --- rebuild superclusterData: one contig holding all windows back to back --- */
    FILE *devnull = fopen("/dev/null", "r");
    std::shared_ptr<fastaData> ref(new fastaData(devnull));      /* empty; closes the FILE */
    const std::string ctg = "vdref";
    const int64_t ref_bytes = in->ref_off[n_sc];
    ref->fasta[ctg] = std::string((const char *)in->ref_seq, (size_t)ref_bytes);
    ref->lengths[ctg] = (int)ref_bytes;

    std::shared_ptr<variantData> qv(new variantData()), tv(new variantData());
    std::shared_ptr<superclusterData> scd(new superclusterData(qv, tv, ref));
    scd->contigs.push_back(ctg);
    scd->lengths.push_back((int)ref_bytes);
    scd->ploidy.push_back(2);
    std::shared_ptr<ctgSuperclusters> cs(new ctgSuperclusters());
    scd->superclusters[ctg] = cs;
    for (int c = 0; c < CALLSETS; c++)
        for (int h = 0; h < HAPS; h++)
            cs->ctg_variants[c][h] = std::shared_ptr<ctgVariants>(new ctgVariants());

    /* batch hap order q1,q2,t1,t2 -> (callset, hap) */
    std::vector<std::vector<int64_t>> batch_idx(4);   /* per (callset,hap): batch variant index */
    for (int s = 0; s < n_sc; s++) {
        const int64_t beg = in->ref_off[s];
        const int64_t end = in->ref_off[s + 1] - 1;
        std::vector<int> brks(4);
        for (int k = 0; k < 4; k++) {
            std::shared_ptr<ctgVariants> cv = cs->ctg_variants[k >> 1][k & 1];
            brks[k] = cv->n;                          /* one cluster per variant */
            for (int64_t v = in->var_off[4 * s + k]; v < in->var_off[4 * s + k + 1]; v++) {
                const int pos = (int)(beg + in->var_pos[v]);
                const int rlen = in->var_rlen[v];
                /* REF allele text: query hap 1 defines the REF-plane string (src/dist.cpp:1784-1792) */
                const uint8_t *rsrc = (k == VD_HAP_Q1 && in->rplane_seq) ? in->rplane_seq : in->ref_seq;
                std::string refa((const char *)rsrc + pos, (size_t)rlen);
                std::string alta((const char *)in->alt_seq + in->alt_off[v],
                                 (size_t)(in->alt_off[v + 1] - in->alt_off[v]));
                cv->add_cluster(cv->n);
                cv->add_var(pos, rlen, k & 1, in->var_type[v], BED_INSIDE, refa, alta,
                            GT_ALT1_REF, 0, in->var_qual[v], 0);
                cv->var_quals.back() = in->var_qual[v];      /* bypass the min(vq, max_qual) clamp */
                batch_idx[k].push_back(v);
            }
        }
        cs->add_supercluster(brks, (int)beg, (int)end);
    }
    for (int k = 0; k < 4; k++) {                     /* closing cluster / supercluster indices */
        std::shared_ptr<ctgVariants> cv = cs->ctg_variants[k >> 1][k & 1];
        cv->add_cluster(cv->n);
        cs->superclusters[k >> 1][k & 1].push_back(cv->n);
    }

    auto sc_groups = sort_superclusters(scd);
    auto t0 = std::chrono::steady_clock::now();
    precision_recall_threads_wrapper(scd, sc_groups);
    auto t1 = std::chrono::steady_clock::now();
    if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();

    /* --- copy results out --- */
    for (int k = 0; k < 4; k++) {
        std::shared_ptr<ctgVariants> cv = cs->ctg_variants[k >> 1][k & 1];
        for (int j = 0; j < cv->n; j++) {
            const int64_t v = batch_idx[k][j];
            for (int p = 0; p < PHASES; p++) {
                const int64_t o = (int64_t)p * n_var + v;
                out->errtypes[o] = cv->errtypes[p][j];
                out->sync_group[o] = cv->sync_group[p][j];
                out->ref_ed[o] = cv->ref_ed[p][j];
                out->query_ed[o] = cv->query_ed[p][j];
                out->callq[o] = cv->callq[p][j];
                out->credit[o] = cv->credit[p][j];
            }
        }
    }
    for (int s = 0; s < n_sc; s++) {
        out->sc_phase[s] = cs->sc_phase[s];
        out->orig_dist[s] = cs->orig_phase_dist[s];
        out->swap_dist[s] = cs->swap_phase_dist[s];
    }
    return 0;
}

/* SURVEY 8f-1 groundwork: the reference's own wf_swg_max_reach (src/dist.cpp:2150-2333) on plain
 * strings, with the offsets buffer prepared as its callers do (src/cluster.cpp:1080-1091, :1141-1152:
 * MATS * (max(open+extend, sub)+1) * (|query|+|truth|-1) entries, all -2).                          */
int vdref_max_reach(const char *query, int query_len, const char *truth, int truth_len,
                    int main_diag, int main_diag_start, int max_score, int sub, int open, int extend,
                    int reverse) {
    const std::string q(query, query + query_len), t(truth, truth + truth_len);
    std::vector<int> offs((size_t)MATS * (std::max(open + extend, sub) + 1) * (q.size() + t.size() - 1), -2);
    return wf_swg_max_reach(q, t, offs, main_diag, main_diag_start, max_score, sub, open, extend,
                            false /* print */, reverse != 0);
}

/* the reference's wf_swg_align (src/dist.cpp:1510-1652): its score */
int vdref_swg_score(const char *query, int query_len, const char *truth, int truth_len, int sub, int open, int extend) {
    const std::string q(query, query + query_len), t(truth, truth + truth_len);
    std::vector<std::vector<std::vector<uint8_t>>> ptrs(MATS);
    std::vector<std::vector<std::vector<int>>> offs(MATS);
    int s = 0;
    wf_swg_align(q, t, ptrs, offs, s, sub, open, extend, false);
    return s;
}

/* the reference's wf_swg_cluster (src/cluster.cpp:954-1263) on one haplotype of one contig: variants
 * (sorted, 0-based positions, REF/ALT texts) in, cluster boundaries and reaches out.  `clusters` has
 * room for n+1 entries (variant index of each cluster start + the sentinel n); returns their number. */
int vdref_cluster(const char *fasta, int fasta_len, int n, const int *pos, const int *rlen, const uint8_t *type,
                  const int64_t *alt_off, const char *alt_seq, int sub, int open, int extend,
                  int max_iters, int reach_min_gap, int *clusters, int *left_reach, int *right_reach) {
    g.max_cluster_itrs = max_iters;
    g.reach_min_gap = reach_min_gap;
    FILE *devnull = fopen("/dev/null", "r");
    std::shared_ptr<fastaData> ref(new fastaData(devnull));
    const std::string ctg = "vdref";
    ref->fasta[ctg] = std::string(fasta, (size_t)fasta_len);
    ref->lengths[ctg] = fasta_len;
    variantData vd;
    vd.ref = ref;
    vd.contigs.push_back(ctg);
    vd.lengths.push_back(fasta_len);
    vd.ploidy.push_back(2);
    vd.variants.resize(HAPS);
    std::shared_ptr<ctgVariants> cv(new ctgVariants());
    for (int h = 0; h < HAPS; h++) vd.variants[h][ctg] = (h == 0) ? cv : std::shared_ptr<ctgVariants>(new ctgVariants());
    for (int v = 0; v < n; v++) {
        std::string refa(fasta + pos[v], (size_t)rlen[v]);
        std::string alta(alt_seq + alt_off[v], (size_t)(alt_off[v + 1] - alt_off[v]));
        cv->add_var(pos[v], rlen[v], 0, type[v], BED_INSIDE, refa, alta, GT_ALT1_REF, 0, 30, 0);
    }
    wf_swg_cluster(&vd, 0, 0, sub, open, extend);
    const int nc = (int)cv->clusters.size();
    for (int i = 0; i < nc; i++) {
        clusters[i] = cv->clusters[i];
        left_reach[i] = i < (int)cv->left_reaches.size() ? cv->left_reaches[i] : 0;
        right_reach[i] = i < (int)cv->right_reaches.size() ? cv->right_reaches[i] : 0;
    }
    return nc;
}

/* the reference's wf_swg_align + wf_swg_backtrack (src/dist.cpp:1510-1652, :2625-2757): score and CIGAR
 * (|query|+|truth| entries) */
int vdref_swg_cigar(const char *query, int query_len, const char *truth, int truth_len, int sub, int open, int extend,
                    int *cigar) {
    const std::string q(query, query + query_len), t(truth, truth + truth_len);
    std::vector<std::vector<std::vector<uint8_t>>> ptrs(MATS);
    std::vector<std::vector<std::vector<int>>> offs(MATS);
    int s = 0;
    wf_swg_align(q, t, ptrs, offs, s, sub, open, extend, false);
    const std::vector<int> c = wf_swg_backtrack(q, t, ptrs, offs, s, sub, open, extend, false);
    for (size_t i = 0; i < c.size(); i++) cigar[i] = c[i];
    return s;
}

}  /* extern "C" */
