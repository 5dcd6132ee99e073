// Drop-in replacement for the reference's hot-path entry point
//
//     void precision_recall_threads_wrapper(
//             std::shared_ptr<superclusterData> clusterdata_ptr,
//             std::vector<std::vector<std::vector<int>>> sc_groups);
//                                     (decl src/dist.h:245-247, def src/dist.cpp:1656-1727)
//
// Same name, arguments, in-place result convention and fatal-error behaviour.  It is
// compiled against the reference's own headers and linked into the reference's CLI in
// place of the original definition; everything else of vcfdist — VCF /
// FASTA / BED parsing, clustering, phasing, output writers — is the reference's code.
//
// Host work here is O(input): pack the superclusters into the compact vd_batch_in
// (include/vcfdist_b200.h), one vd_run() on the GPU, vd_finalize() for the float step,
// scatter into ctgVariants / ctgSuperclusters (src/variant.h:49-60, src/cluster.h:36-42).
//
// Build-time switch VD_DROPIN_WITH_REF (oracle/Makefile only, never the product build):
// adds the fixture-dump mode that runs the REFERENCE's renamed wrapper instead of the GPU
// and writes its results, used to generate tests/golden/.
#include <array>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "globals.h"
#include "variant.h"
#include "cluster.h"
#include "dist.h"

#include "vcfdist_b200.h"

#ifdef VD_DROPIN_WITH_REF
void ref_precision_recall_threads_wrapper(
        std::shared_ptr<superclusterData> clusterdata_ptr,
        std::vector< std::vector< std::vector<int> > > sc_groups);
#endif

namespace vdhost {

struct VarLoc { int ctg, callset, hap, idx; };
struct ScLoc { int ctg, sc; };

struct Packed {
    std::vector<int64_t> ref_off{0}, var_off{0}, alt_off{0};
    std::vector<uint8_t> ref_seq, rplane_seq, alt_seq, var_type;
    std::vector<int32_t> var_pos, var_rlen;
    std::vector<float> var_qual;
    bool need_rplane = false;
    std::vector<VarLoc> var_loc;
    std::vector<ScLoc> sc_loc;

    vd_batch_in view(float max_qual) const {
        vd_batch_in in;
        in.n_sc = (int32_t)sc_loc.size();
        in.ref_off = ref_off.data();
        in.ref_seq = ref_seq.data();
        in.rplane_seq = need_rplane ? rplane_seq.data() : nullptr;
        in.var_off = var_off.data();
        in.var_pos = var_pos.data();
        in.var_rlen = var_rlen.data();
        in.var_type = var_type.data();
        in.alt_off = alt_off.data();
        in.alt_seq = alt_seq.data();
        in.var_qual = var_qual.data();
        in.max_qual = max_qual;
        return in;
    }
};

// what precision_recall_wrapper reads per supercluster (src/dist.cpp:1786-1822)
static void pack_one(const superclusterData *scd, int ctg_idx, int sc_idx, Packed &p) {
    const std::string &ctg = scd->contigs[ctg_idx];
    const std::shared_ptr<ctgSuperclusters> &sc = scd->superclusters.at(ctg);
    const int beg = sc->begs[sc_idx], end = sc->ends[sc_idx];
    const std::string &fa = scd->ref->fasta.at(ctg);
    if (beg < 0 || end >= (int)fa.size() || end < beg)
        ERROR("Contig '%s' not present in reference FASTA", ctg.data());   // src/dist.cpp:237-239
    const size_t r0 = p.ref_seq.size();
    p.ref_seq.insert(p.ref_seq.end(), fa.begin() + beg, fa.begin() + end + 1);
    p.rplane_seq.insert(p.rplane_seq.end(), fa.begin() + beg, fa.begin() + end + 1);
    p.ref_off.push_back((int64_t)p.ref_seq.size());
    for (int k = 0; k < CALLSETS * HAPS; k++) {
        const std::shared_ptr<ctgVariants> &cv = sc->ctg_variants[k >> 1][k & 1];
        int vb = 0, ve = 0;
        if (cv->clusters.size()) {                                          // src/dist.cpp:159-162
            vb = cv->clusters[sc->superclusters[k >> 1][k & 1][sc_idx]];
            ve = cv->clusters[sc->superclusters[k >> 1][k & 1][sc_idx + 1]];
        }
        for (int v = vb; v < ve; v++) {
            p.var_pos.push_back(cv->poss[v] - beg);
            p.var_rlen.push_back((int32_t)cv->refs[v].size());
            p.var_type.push_back(cv->types[v]);
            p.alt_seq.insert(p.alt_seq.end(), cv->alts[v].begin(), cv->alts[v].end());
            p.alt_off.push_back((int64_t)p.alt_seq.size());
            p.var_qual.push_back(cv->var_quals[v]);
            p.var_loc.push_back({ctg_idx, k >> 1, k & 1, v});
            // the REF-plane string is ref_q1: FASTA with query-hap-1 REF alleles written in
            // (src/dist.cpp:187, :195, :1784-1792)
            if (k == 0 && (cv->types[v] == TYPE_DEL || cv->types[v] == TYPE_SUB)) {
                const int rel = cv->poss[v] - beg;
                for (size_t j = 0; j < cv->refs[v].size(); j++) {
                    const size_t at = r0 + rel + j;
                    if (rel >= 0 && at < p.rplane_seq.size() && p.rplane_seq[at] != (uint8_t)cv->refs[v][j]) {
                        p.rplane_seq[at] = (uint8_t)cv->refs[v][j];
                        p.need_rplane = true;
                    }
                }
            }
        }
        p.var_off.push_back((int64_t)p.var_pos.size());
    }
    p.sc_loc.push_back({ctg_idx, sc_idx});
}

static Packed pack(const superclusterData *scd,
                   const std::vector<std::vector<std::vector<int>>> &sc_groups) {
    Packed p;
    // largest-RAM bucket first, as the reference schedules them (src/dist.cpp:1670-1672)
    for (int step = (int)sc_groups.size() - 1; step >= 0; step--)
        for (size_t k = 0; k < sc_groups[step][SC_IDX].size(); k++)
            pack_one(scd, sc_groups[step][CTG_IDX][k], sc_groups[step][SC_IDX][k], p);
    return p;
}

static void scatter(superclusterData *scd, const Packed &p, const vd_final &fin) {
    const int64_t n_var = (int64_t)p.var_loc.size();
    for (int64_t v = 0; v < n_var; v++) {
        const VarLoc &l = p.var_loc[v];
        ctgVariants &cv = *scd->superclusters.at(scd->contigs[l.ctg])->ctg_variants[l.callset][l.hap];
        for (int slot = 0; slot < PHASES; slot++) {
            const int64_t o = slot * n_var + v;
            cv.errtypes[slot][l.idx] = fin.errtypes[o];
            cv.sync_group[slot][l.idx] = fin.sync_group[o];
            cv.credit[slot][l.idx] = fin.credit[o];
            cv.ref_ed[slot][l.idx] = fin.ref_ed[o];
            cv.query_ed[slot][l.idx] = fin.query_ed[o];
            cv.callq[slot][l.idx] = fin.callq[o];
        }
    }
    for (size_t s = 0; s < p.sc_loc.size(); s++)
        scd->superclusters.at(scd->contigs[p.sc_loc[s].ctg])->set_phase(     // src/cluster.cpp:31-38
                p.sc_loc[s].sc, fin.sc_phase[s], fin.orig_dist[s], fin.swap_dist[s]);
}

// ---- tiny self-describing array container, read by vcfdist_b200/fixtures.py ----
static void put_arr(FILE *f, const char *name, char dtype, const void *data, int64_t n, int itemsize) {
    char nm[24] = {0};
    std::strncpy(nm, name, 23);
    std::fwrite(nm, 1, 24, f);
    std::fwrite(&dtype, 1, 1, f);
    std::fwrite(&n, 8, 1, f);
    if (n) std::fwrite(data, (size_t)itemsize, (size_t)n, f);
}

static void dump_batch(const Packed &p, float max_qual, const char *path) {
    FILE *f = std::fopen(path, "wb");
    if (!f) ERROR("cannot write '%s'", path);
    std::fwrite("VDARR001", 1, 8, f);
    put_arr(f, "ref_off", 'q', p.ref_off.data(), (int64_t)p.ref_off.size(), 8);
    put_arr(f, "ref_seq", 'B', p.ref_seq.data(), (int64_t)p.ref_seq.size(), 1);
    if (p.need_rplane) put_arr(f, "rplane_seq", 'B', p.rplane_seq.data(), (int64_t)p.rplane_seq.size(), 1);
    put_arr(f, "var_off", 'q', p.var_off.data(), (int64_t)p.var_off.size(), 8);
    put_arr(f, "var_pos", 'i', p.var_pos.data(), (int64_t)p.var_pos.size(), 4);
    put_arr(f, "var_rlen", 'i', p.var_rlen.data(), (int64_t)p.var_rlen.size(), 4);
    put_arr(f, "var_type", 'B', p.var_type.data(), (int64_t)p.var_type.size(), 1);
    put_arr(f, "alt_off", 'q', p.alt_off.data(), (int64_t)p.alt_off.size(), 8);
    put_arr(f, "alt_seq", 'B', p.alt_seq.data(), (int64_t)p.alt_seq.size(), 1);
    put_arr(f, "var_qual", 'f', p.var_qual.data(), (int64_t)p.var_qual.size(), 4);
    put_arr(f, "max_qual", 'f', &max_qual, 1, 4);
    std::fclose(f);
}

// results as they stand in superclusterData, in batch order (whoever computed them)
static void dump_final(const superclusterData *scd, const Packed &p, const char *path) {
    const int64_t n_var = (int64_t)p.var_loc.size(), n_sc = (int64_t)p.sc_loc.size();
    std::vector<uint8_t> err(2 * n_var);
    std::vector<int32_t> sg(2 * n_var), red(2 * n_var), qed(2 * n_var), ph(n_sc), od(n_sc), sd(n_sc);
    std::vector<float> cq(2 * n_var), cr(2 * n_var);
    for (int64_t v = 0; v < n_var; v++) {
        const VarLoc &l = p.var_loc[v];
        const ctgVariants &cv = *scd->superclusters.at(scd->contigs[l.ctg])->ctg_variants[l.callset][l.hap];
        for (int slot = 0; slot < PHASES; slot++) {
            const int64_t o = slot * n_var + v;
            err[o] = cv.errtypes[slot][l.idx]; sg[o] = cv.sync_group[slot][l.idx];
            red[o] = cv.ref_ed[slot][l.idx]; qed[o] = cv.query_ed[slot][l.idx];
            cq[o] = cv.callq[slot][l.idx]; cr[o] = cv.credit[slot][l.idx];
        }
    }
    for (int64_t s = 0; s < n_sc; s++) {
        const ctgSuperclusters &cs = *scd->superclusters.at(scd->contigs[p.sc_loc[s].ctg]);
        ph[s] = cs.sc_phase[p.sc_loc[s].sc];
        od[s] = cs.orig_phase_dist[p.sc_loc[s].sc];
        sd[s] = cs.swap_phase_dist[p.sc_loc[s].sc];
    }
    FILE *f = std::fopen(path, "wb");
    if (!f) ERROR("cannot write '%s'", path);
    std::fwrite("VDARR001", 1, 8, f);
    put_arr(f, "errtypes", 'B', err.data(), 2 * n_var, 1);
    put_arr(f, "sync_group", 'i', sg.data(), 2 * n_var, 4);
    put_arr(f, "ref_ed", 'i', red.data(), 2 * n_var, 4);
    put_arr(f, "query_ed", 'i', qed.data(), 2 * n_var, 4);
    put_arr(f, "callq", 'f', cq.data(), 2 * n_var, 4);
    put_arr(f, "credit", 'f', cr.data(), 2 * n_var, 4);
    put_arr(f, "sc_phase", 'i', ph.data(), n_sc, 4);
    put_arr(f, "orig_dist", 'i', od.data(), n_sc, 4);
    put_arr(f, "swap_dist", 'i', sd.data(), n_sc, 4);
    std::fclose(f);
}

// the dump the reference prints to stdout next to a data WARN (src/dist.cpp:1225-1252)
static void print_supercluster(const superclusterData *scd, int ctg_idx, int sc_idx) {
    const std::string &ctg = scd->contigs[ctg_idx];
    std::printf("\n\nSupercluster: %d\n", sc_idx);
    std::shared_ptr<ctgSuperclusters> sc = scd->superclusters.at(ctg);
    for (int j = 0; j < CALLSETS * HAPS; j++) {
        int callset = j >> 1, hap = j % 2;
        int cluster_beg = sc->superclusters[callset][hap][sc_idx];
        int cluster_end = sc->superclusters[callset][hap][sc_idx + 1];
        std::printf("%s%d: %d clusters (%d-%d)\n", callset_strs[callset].data(), hap + 1,
                    cluster_end - cluster_beg, cluster_beg, cluster_end);
        for (int k = cluster_beg; k < cluster_end; k++) {
            std::shared_ptr<ctgVariants> vars = sc->ctg_variants[callset][hap];
            int variant_beg = vars->clusters[k], variant_end = vars->clusters[k + 1];
            std::printf("\tCluster %d: %d variants (%d-%d)\n", k, variant_end - variant_beg,
                        variant_beg, variant_end);
            for (int l = variant_beg; l < variant_end; l++)
                std::printf("\t\t%s %d\t%s\t%s\tQ=%f\n", ctg.data(), vars->poss[l],
                            vars->refs[l].size() ? vars->refs[l].data() : "_",
                            vars->alts[l].size() ? vars->alts[l].data() : "_", vars->var_quals[l]);
        }
    }
}

}  // namespace vdhost

void precision_recall_threads_wrapper(
        std::shared_ptr<superclusterData> clusterdata_ptr,
        std::vector< std::vector< std::vector<int> > > sc_groups) {
    // same log lines as src/dist.cpp:1659-1668
    if (g.verbosity >= 1) INFO(" ");
    if (g.verbosity >= 1) INFO("%s[5/8] Calculating precision and recall%s",
            COLOR_PURPLE, COLOR_WHITE);
    if (g.verbosity >= 1)
    for (int i = 0; i < g.thread_nsteps; i++) {
        INFO("  Superclusters using %7.3f to %7.3f GB RAM each (%3d threads): %8d",
                i == 0 ? 0 : g.ram_steps[i-1], g.ram_steps[i], g.thread_steps[i],
                int(sc_groups[i][CTG_IDX].size()));
    }

    vdhost::Packed p = vdhost::pack(clusterdata_ptr.get(), sc_groups);
    const float max_qual = float(g.max_qual);                       // src/dist.cpp:1284
    if (const char *path = std::getenv("VD_DUMP_BATCH")) vdhost::dump_batch(p, max_qual, path);

#ifdef VD_DROPIN_WITH_REF
    {   // fixture tool (oracle/_ref/vcfdist_dump): the REFERENCE computes, we only record
        ref_precision_recall_threads_wrapper(clusterdata_ptr, sc_groups);
        if (const char *path = std::getenv("VD_DUMP_FINAL"))
            vdhost::dump_final(clusterdata_ptr.get(), p, path);
        return;
    }
#else

    const int64_t n_sc = (int64_t)p.sc_loc.size(), n_var = (int64_t)p.var_loc.size();
    if (n_sc == 0) return;
    vd_batch_in in = p.view(max_qual);

    std::vector<int32_t> aln_score(4 * n_sc), sync_group(2 * n_var + 1), ref_ed(2 * n_var + 1),
            query_ed(2 * n_var + 1);
    std::vector<uint8_t> end_plane(4 * n_sc), beg_plane(4 * n_sc), assigned(2 * n_var + 1);
    std::vector<uint32_t> status(4 * n_sc);
    std::vector<float> callq(2 * n_var + 1);
    vd_batch_out out{aln_score.data(), end_plane.data(), beg_plane.data(), status.data(),
                     assigned.data(), sync_group.data(), ref_ed.data(), query_ed.data(), callq.data()};

    int device = 0;
    if (const char *d = std::getenv("VD_DEVICE")) device = std::atoi(d);
    vd_handle *h = nullptr;
    int rc = vd_create(device, 0, &h);
    if (rc != VD_OK) ERROR("vcfdist_b200: cannot initialise CUDA device %d (code %d); "
                           "there is no CPU fallback for the precision/recall path", device, rc);
    rc = vd_run(h, &in, &out);
    if (rc != VD_OK && rc != VD_E_ALIGN && rc != VD_E_BADINPUT)     // the last two are reported per supercluster below
        ERROR("vcfdist_b200: vd_run failed (code %d): %s", rc, vd_last_error(h));
    if (g.verbosity >= 2) {
        vd_stats st; vd_get_stats(h, &st);
        INFO("  GPU precision/recall: %lld superclusters, %lld cells, %.3f ms on device, %lld launches",
             (long long)st.n_sc, (long long)st.cells, st.ms_total, (long long)st.n_launches);
    }
    vd_destroy(h);

    // fatal conditions and data warnings of the reference, in batch order
    for (int64_t s = 0; s < n_sc; s++) {
        uint32_t st = status[4 * s] | status[4 * s + 1] | status[4 * s + 2] | status[4 * s + 3];
        if (!st) continue;
        const std::string &ctg = clusterdata_ptr->contigs[p.sc_loc[s].ctg];
        const int sc_idx = p.sc_loc[s].sc;
        if (st & VD_ST_ERR_BADINPUT)
            ERROR("vcfdist_b200: ctg %s supercluster %d cannot be processed on the GPU path (malformed variants, or more than "
                  "8 swap sources for one row, e.g. 8 adjacent deletions)", ctg.data(), sc_idx);
        if (st & VD_ST_ERR_MASK & ~(VD_ST_ERR_BADINPUT | VD_ST_ERR_UNFINISHED | VD_ST_ERR_NO_SWAP_PRED | VD_ST_ERR_NO_POINTER))
            ERROR("vcfdist_b200: unknown error status 0x%x at ctg %s supercluster %d", st, ctg.data(), sc_idx);
        if (st & VD_ST_ERR_UNFINISHED) ERROR("Alignment not finished in 'prec_recall_aln()'.");   // :440
        if (st & VD_ST_ERR_NO_SWAP_PRED) ERROR("No swap predecessor, but PTR_SWP_MAT set.");      // :606
        if (st & VD_ST_ERR_NO_POINTER) ERROR("No valid pointer at ctg %s supercluster %d", ctg.data(), sc_idx); // :937
        bool warn = false;
        if (st & VD_ST_WARN_REFED_NOTRUTH) { warn = true;                                          // :1204
            WARN("Nonzero reference edit distance with no truth variants at ctg %s supercluster %d", ctg.data(), sc_idx); }
        if (st & VD_ST_WARN_QED_NOQUERY) { warn = true;                                            // :1208
            WARN("Query edit distance changed with no query variants at ctg %s supercluster %d", ctg.data(), sc_idx); }
        if (st & VD_ST_WARN_QED_GT_REFED) { warn = true;                                           // :1212
            WARN("Query edit distance exceeds reference edit distance at ctg %s supercluster %d", ctg.data(), sc_idx); }
        if (st & VD_ST_WARN_ZERO_REFED) { warn = true;                                             // :1221
            WARN("Zero edit distance with truth variants at ctg %s supercluster %d", ctg.data(), sc_idx); }
        if (warn) vdhost::print_supercluster(clusterdata_ptr.get(), p.sc_loc[s].ctg, sc_idx);
    }

    std::vector<uint8_t> errtypes(2 * n_var + 1);
    std::vector<float> credit(2 * n_var + 1), fcallq(2 * n_var + 1);
    std::vector<int32_t> fsg(2 * n_var + 1), fred(2 * n_var + 1), fqed(2 * n_var + 1),
            phase(n_sc), od(n_sc), sd(n_sc);
    vd_final fin{errtypes.data(), credit.data(), fcallq.data(), fsg.data(), fred.data(), fqed.data(),
                 phase.data(), od.data(), sd.data()};
    vd_finalize(&in, &out, g.phase_threshold, g.credit_threshold, &fin);
    vdhost::scatter(clusterdata_ptr.get(), p, fin);
#endif  // VD_DROPIN_WITH_REF
}
