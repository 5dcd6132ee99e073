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

#include "vcfdist_b200.h"

namespace {
// src/defs.h:66-72, :131-134
constexpr uint8_t kTP = 0, kFP = 1, kFN = 2, kUN = 5;
constexpr int kPhaseOrig = 0, kPhaseSwap = 1, kPhaseNone = 2;
}  // namespace

extern "C" int vd_finalize(const vd_batch_in *in, const vd_batch_out *out,
                           double phase_threshold, double credit_threshold, vd_final *fin) {
    if (!in || !out || !fin) return VD_E_BADINPUT;
    const int64_t n_sc = in->n_sc;
    const int64_t n_var = in->var_off[4 * n_sc];

    // ---- store_phase, src/dist.cpp:449-475 ----
    for (int64_t s = 0; s < n_sc; s++) {
        const int32_t *sc = out->aln_score + 4 * s;
        int orig_phase_dist = sc[0] + sc[3];   // QUERY1_TRUTH1 + QUERY2_TRUTH2
        int swap_phase_dist = sc[2] + sc[1];   // QUERY2_TRUTH1 + QUERY1_TRUTH2
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
    }

    // ---- per-variant credit, src/dist.cpp:1157-1168, :1291-1353 ----
    // defaults are the reference's initial values (src/variant.cpp:45-52)
    for (int64_t i = 0; i < 2 * n_var; i++) {
        fin->errtypes[i] = kUN;
        fin->credit[i] = 0;
        fin->callq[i] = 0;
        fin->sync_group[i] = 0;
        fin->ref_ed[i] = 0;
        fin->query_ed[i] = 0;
    }
    for (int64_t s = 0; s < n_sc; s++) {
        for (int h = 0; h < 4; h++) {
            const bool is_truth = h >= 2;
            for (int64_t v = in->var_off[4 * s + h]; v < in->var_off[4 * s + h + 1]; v++) {
                for (int slot = 0; slot < 2; slot++) {
                    const int64_t o = slot * n_var + v;
                    const uint8_t a = out->assigned[o];
                    if (a == VD_ASSIGN_REF_FP) {              // :1157-1168
                        fin->errtypes[o] = kFP;
                        fin->sync_group[o] = out->sync_group[o];
                        fin->credit[o] = 0;
                        fin->ref_ed[o] = 0;
                        fin->query_ed[o] = 0;
                        fin->callq[o] = out->callq[o];
                    } else if (a == VD_ASSIGN_SYNC) {
                        const int ref_ed = out->ref_ed[o];
                        const int query_ed = out->query_ed[o];
                        float credit = 1 - float(query_ed) / ref_ed;      // :1293, :1327
                        fin->sync_group[o] = out->sync_group[o];
                        fin->credit[o] = credit;
                        fin->ref_ed[o] = ref_ed;
                        fin->query_ed[o] = query_ed;
                        if (credit >= credit_threshold) {                 // :1296, :1328
                            fin->errtypes[o] = kTP;
                            fin->callq[o] = out->callq[o];
                        } else if (is_truth) {                            // :1340-1346
                            fin->errtypes[o] = kFN;
                            fin->callq[o] = in->max_qual;
                        } else {                                          // :1308-1314
                            fin->errtypes[o] = kFP;
                            fin->callq[o] = out->callq[o];
                        }
                    }
                }
            }
        }
    }
    return VD_OK;
}
