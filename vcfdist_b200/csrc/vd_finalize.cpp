// Host-only float step of the precision/recall path.
//
// The device returns integers only (scores, ref_ed, query_ed, sync groups) plus exact
// float minima (callq).  The two threshold decisions of the reference are evaluated here
// with the reference's own expression shapes, so that the stored `credit` floats and the
// TP/FP/FN and phase decisions are bit-identical:
//   store_phase        src/dist.cpp:449-475   1 - float(a)/b > double threshold
//   calc_prec_recall   src/dist.cpp:1293-1352 float credit = 1 - float(query_ed)/ref_ed;
//                                              credit >= double threshold
#include <cstdint>
#include <cstdlib>
#include <thread>
#include <vector>

#include "vcfdist_b200.h"

namespace {
// src/defs.h:66-72, :131-134
constexpr uint8_t kTP = 0, kFP = 1, kFN = 2, kUN = 5;
constexpr int kPhaseOrig = 0, kPhaseSwap = 1, kPhaseNone = 2;

// the device's records, wide (vd_batch_out) or 16-bit (vd_packed_out)
struct WideOut {
    const vd_batch_out *o;
    int score(int64_t i) const { return o->aln_score[i]; }
    int assigned(int64_t i) const { return o->assigned[i]; }
    int sync_group(int64_t i) const { return o->sync_group[i]; }
    int ref_ed(int64_t i) const { return o->ref_ed[i]; }
    int query_ed(int64_t i) const { return o->query_ed[i]; }
    float callq(int64_t i) const { return o->callq[i]; }
};
struct PackedOut {
    const vd_packed_out *o;
    int score(int64_t i) const { return o->aln_score[i] == 0xffff ? -1 : (int)o->aln_score[i]; }
    int assigned(int64_t i) const { return o->sync_group[i] >> 14; }
    int sync_group(int64_t i) const { return o->sync_group[i] & 0x3fff; }
    int ref_ed(int64_t i) const { return o->ref_ed[i]; }
    int query_ed(int64_t i) const { return o->query_ed[i]; }
    float callq(int64_t i) const { return o->callq[i]; }
};

// superclusters [s0, s1): independent units, disjoint variant ranges
template <class O>
void finalize_range(const vd_batch_in *in, const O &out, double phase_threshold, double credit_threshold, vd_final *fin,
                    int64_t s0, int64_t s1) {
    const int64_t n_var = in->var_off[4 * (int64_t)in->n_sc];
    for (int64_t s = s0; s < s1; s++) {
        // ---- store_phase, src/dist.cpp:449-475 ----
        int orig_phase_dist = out.score(4 * s + 0) + out.score(4 * s + 3);   // QUERY1_TRUTH1 + QUERY2_TRUTH2
        int swap_phase_dist = out.score(4 * s + 2) + out.score(4 * s + 1);   // QUERY2_TRUTH1 + QUERY1_TRUTH2
        int phase = kPhaseNone;
        if (orig_phase_dist == swap_phase_dist) {
            phase = kPhaseNone;
        } else if (orig_phase_dist == 0) {
            phase = kPhaseOrig;
        } else if (swap_phase_dist == 0) {
            phase = kPhaseSwap;
        } else if (1 - float(swap_phase_dist) / orig_phase_dist > phase_threshold) {
            phase = kPhaseSwap;
        } else if (1 - float(orig_phase_dist) / swap_phase_dist > phase_threshold) {
            phase = kPhaseOrig;
        }
        fin->sc_phase[s] = phase;
        fin->orig_dist[s] = orig_phase_dist;
        fin->swap_dist[s] = swap_phase_dist;
        // ---- per-variant credit, src/dist.cpp:1157-1168, :1291-1353 ----
        for (int h = 0; h < 4; h++) {
            const bool is_truth = h >= 2;
            for (int64_t v = in->var_off[4 * s + h]; v < in->var_off[4 * s + h + 1]; v++) {
                for (int slot = 0; slot < 2; slot++) {
                    const int64_t o = slot * n_var + v;
                    // defaults are the reference's initial values (src/variant.cpp:45-52)
                    fin->errtypes[o] = kUN; fin->credit[o] = 0; fin->callq[o] = 0;
                    fin->sync_group[o] = 0; fin->ref_ed[o] = 0; fin->query_ed[o] = 0;
                    const int a = out.assigned(o);
                    if (a == VD_ASSIGN_REF_FP) {              // :1157-1168
                        fin->errtypes[o] = kFP;
                        fin->sync_group[o] = out.sync_group(o);
                        fin->callq[o] = out.callq(o);
                    } else if (a == VD_ASSIGN_SYNC) {
                        const int ref_ed = out.ref_ed(o);
                        const int query_ed = out.query_ed(o);
                        float credit = 1 - float(query_ed) / ref_ed;      // :1293, :1327
                        fin->sync_group[o] = out.sync_group(o);
                        fin->credit[o] = credit;
                        fin->ref_ed[o] = ref_ed;
                        fin->query_ed[o] = query_ed;
                        if (credit >= credit_threshold) {                 // :1296, :1328
                            fin->errtypes[o] = kTP;
                            fin->callq[o] = out.callq(o);
                        } else if (is_truth) {                            // :1340-1346
                            fin->errtypes[o] = kFN;
                            fin->callq[o] = in->max_qual;
                        } else {                                          // :1308-1314
                            fin->errtypes[o] = kFP;
                            fin->callq[o] = out.callq(o);
                        }
                    }
                }
            }
        }
    }
}

// big batches are split over host threads (the reference runs this step inside its alignment threads)
template <class O>
int finalize_all(const vd_batch_in *in, const O &out, double phase_threshold, double credit_threshold, vd_final *fin) {
    const int64_t n_sc = in->n_sc;
    unsigned nt = std::thread::hardware_concurrency();
    if (const char *e = std::getenv("VD_HOST_THREADS")) nt = (unsigned)std::atoi(e);
    if (nt < 1) nt = 1;
    if (nt > 32) nt = 32;
    if (n_sc < 200000 || nt == 1) {
        finalize_range(in, out, phase_threshold, credit_threshold, fin, 0, n_sc);
        return VD_OK;
    }
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; t++)
        th.emplace_back([=, &out]() { finalize_range(in, out, phase_threshold, credit_threshold, fin, n_sc * t / nt, n_sc * (t + 1) / nt); });
    for (auto &x : th) x.join();
    return VD_OK;
}
}  // namespace

extern "C" int vd_finalize(const vd_batch_in *in, const vd_batch_out *out,
                           double phase_threshold, double credit_threshold, vd_final *fin) {
    if (!in || !out || !fin) return VD_E_BADINPUT;
    return finalize_all(in, WideOut{out}, phase_threshold, credit_threshold, fin);
}

extern "C" int vd_finalize_packed(const vd_batch_in *in, const vd_packed_out *out,
                                  double phase_threshold, double credit_threshold, vd_final *fin) {
    if (!in || !out || !fin) return VD_E_BADINPUT;
    return finalize_all(in, PackedOut{out}, phase_threshold, credit_threshold, fin);
}

// compact form of a batch (host only; include/vcfdist_b200.h: vd_compact_in)
extern "C" int vd_compact_pack(const vd_batch_in *in, vd_compact_in *out) {
    if (!in || !out) return VD_E_BADINPUT;
    const int64_t n_sc = in->n_sc;
    const int64_t n_var = in->var_off[4 * n_sc];
    uint16_t *ref_len = const_cast<uint16_t *>(out->ref_len), *pos = const_cast<uint16_t *>(out->var_pos),
             *rlen = const_cast<uint16_t *>(out->var_rlen), *alen = const_cast<uint16_t *>(out->alt_len);
    uint8_t *nvar = const_cast<uint8_t *>(out->hap_nvar);
    int64_t *bv = const_cast<int64_t *>(out->blk_var), *br = const_cast<int64_t *>(out->blk_ref), *ba = const_cast<int64_t *>(out->blk_alt);
    out->n_sc = in->n_sc; out->max_qual = in->max_qual;
    out->n_var = n_var; out->ref_bytes = in->ref_off[n_sc]; out->alt_bytes = n_var ? in->alt_off[n_var] : 0;
    out->ref_seq = in->ref_seq; out->rplane_seq = in->rplane_seq; out->var_type = in->var_type;
    out->alt_seq = in->alt_seq; out->var_qual = in->var_qual;
    bool fits = true;
    for (int64_t s = 0; s < n_sc; s++) {
        const int64_t len = in->ref_off[s + 1] - in->ref_off[s];
        fits &= len >= 0 && len < 65536;
        ref_len[s] = (uint16_t)len;
        for (int k = 0; k < 4; k++) {
            const int64_t n = in->var_off[4 * s + k + 1] - in->var_off[4 * s + k];
            fits &= n >= 0 && n < 256;
            nvar[4 * s + k] = (uint8_t)n;
        }
        if (s % VD_COMPACT_BLOCK == 0) {
            const int64_t b = s / VD_COMPACT_BLOCK, v0 = in->var_off[4 * s];
            bv[b] = v0; br[b] = in->ref_off[s]; ba[b] = n_var ? in->alt_off[v0] : 0;
        }
    }
    const int64_t n_blk = (n_sc + VD_COMPACT_BLOCK - 1) / VD_COMPACT_BLOCK;
    bv[n_blk] = n_var; br[n_blk] = out->ref_bytes; ba[n_blk] = out->alt_bytes;
    for (int64_t v = 0; v < n_var; v++) {
        const int64_t al = in->alt_off[v + 1] - in->alt_off[v];
        fits &= al >= 0 && al < 65536 && in->var_pos[v] >= 0 && in->var_pos[v] < 65536 && in->var_rlen[v] >= 0 && in->var_rlen[v] < 65536;
        pos[v] = (uint16_t)in->var_pos[v]; rlen[v] = (uint16_t)in->var_rlen[v]; alen[v] = (uint16_t)al;
    }
    return fits ? VD_OK : VD_E_RANGE;
}
