/* TEST INFRASTRUCTURE - plain-C restatement of the reference's wf_swg_max_reach
 * (/root/reference/src/dist.cpp:2150-2333): affine-gap (Gotoh) wavefront search that returns the
 * farthest truth (reference) index reachable with a score <= max_score when every variant of the
 * cluster must be taken; the cluster-growing step of SURVEY.md 8f-1 (src/cluster.cpp:954-1263 calls
 * it with iterative doubling).  Groundwork for the next row of the scope table: only the oracle
 * exists so far, pinned against the reference's object code by tests/test_reach_oracle.py.
 *
 * Wavefront layout.  Diagonal index d in [0, nd), nd = |query| + |truth| - 1, stands for
 * k = d + 1 - |query| (truth index minus query index).  Three wavefront kinds (M: match/substitution,
 * I: insertion = query base consumed, D: deletion = truth base consumed), each holding for every
 * diagonal the furthest QUERY index reached, NONE (-2) when the diagonal is not reached; the M
 * wavefront of score 0 starts at query index -1 on the main diagonal k = 0.  Only the last
 * max(x, o+e) + 1 scores are kept (ring buffer).
 *
 * forward  (reverse = 0): opening a gap costs o+e, extending e, leaving a gap is free and happens
 *          at the start of the same score (I/D wavefronts are merged into M before extension);
 * reverse  (reverse = 1): entering a gap costs e, leaving it costs o (the string pair is reversed
 *          by the caller, so the gap-open penalty is paid on the other side).
 * main_diag / main_diag_start: on the diagonal k = main_diag the free extension stops one before
 * query index (main_diag_start - main_diag) - matching reference bases beyond the cluster is not a
 * "reach" (src/dist.cpp:2144-2147).                                                               */
#include <stdlib.h>

#define NONE (-2)
enum { WM = 0, WI = 1, WD = 2, NW = 3 };

typedef struct {
    int *v;              /* [NW][ring][nd] */
    int nd, ring;
} wf_t;

static inline int *wf(const wf_t *w, int kind, int slot, int d) {
    return w->v + ((size_t)kind * w->ring + slot) * w->nd + d;
}
static inline int slot_of(int slot, int back, int ring) {
    int s = slot - back;
    return s < 0 ? s + ring : s;
}

int vdo_max_reach(const char *query, int qlen, const char *truth, int tlen,
                  int main_diag, int main_diag_start, int max_score, int x, int o, int e, int reverse) {
    wf_t w;
    w.nd = qlen + tlen - 1;
    w.ring = (x > o + e ? x : o + e) + 1;
    const size_t total = (size_t)NW * w.ring * w.nd;
    w.v = (int *)malloc(sizeof(int) * (total ? total : 1));
    for (size_t i = 0; i < total; i++) w.v[i] = NONE;
    const int stop_q = main_diag_start - main_diag;      /* :2160 */
    int score = 0, slot = 0, result = -1;
    *wf(&w, WM, slot, qlen - 1) = -1;                    /* :2166: diagonal k = 0, before the first base */

    for (;;) {
        /* forward: gaps are left for free at the score they were reached with   (:2171-2184) */
        if (!reverse)
            for (int kind = WI; kind <= WD; kind++)
                for (int d = 0; d < w.nd; d++) {
                    const int q = *wf(&w, kind, slot, d), k = d + 1 - qlen;
                    if (q >= 0 && q < qlen && k + q >= 0 && k + q < tlen && q >= *wf(&w, WM, slot, d))
                        *wf(&w, WM, slot, d) = q;
                }
        /* free extension along matches, then the two exits   (:2187-2212) */
        for (int d = 0; d < w.nd; d++) {
            int q = *wf(&w, WM, slot, d);
            const int k = d + 1 - qlen;
            while ((k != main_diag || q + 1 < stop_q) && q != NONE && k + q >= -1 &&
                   q < qlen - 1 && k + q < tlen - 1 && query[q + 1] == truth[k + q + 1])
                q++;
            *wf(&w, WM, slot, d) = q;
            if (q + k == tlen - 1) { result = tlen - 1; goto done; }                     /* :2205-2207 */
            if (q == qlen - 1 && q + k >= 0 && q + k < tlen - 1) { result = q + k; goto done; }   /* :2208-2210 */
        }
        if (score == max_score) break;                   /* :2213 */

        /* next score   (:2225-2311) */
        score++;
        slot = slot + 1 == w.ring ? 0 : slot + 1;
        for (int kind = WM; kind < NW; kind++)           /* :2228-2232 clears I and D; M of this slot is */
            for (int d = 0; d < w.nd; d++)               /* overwritten below only through >= tests      */
                if (kind != WM) *wf(&w, kind, slot, d) = NONE;
        for (int d = 0; d < w.nd; d++) {
            const int k = d + 1 - qlen;
            int *m_cur = wf(&w, WM, slot, d), *i_cur = wf(&w, WI, slot, d), *d_cur = wf(&w, WD, slot, d);
            /* substitution   (:2239-2250) */
            if (score - x >= 0) {
                const int p = *wf(&w, WM, slot_of(slot, x, w.ring), d);
                if (p != NONE && p + 1 < qlen && k + p + 1 < tlen && p + 1 >= *m_cur) *m_cur = p + 1;
            }
            /* gap opening: o+e forwards, e in the reversed problem   (:2252-2275) */
            {
                const int cost = reverse ? e : o + e;
                if (score - cost >= 0) {
                    const int ps = slot_of(slot, cost, w.ring);
                    if (d > 0) {
                        const int p = *wf(&w, WM, ps, d - 1);
                        if (p != NONE && k + p < tlen && p >= *d_cur) *d_cur = p;
                    }
                    if (d < w.nd - 1) {
                        const int p = *wf(&w, WM, ps, d + 1);
                        if (p != NONE && p + 1 < qlen && k + p + 1 < tlen && k + p + 1 >= 0 && p + 1 >= *i_cur) *i_cur = p + 1;
                    }
                }
            }
            /* reversed problem: leaving a gap costs o   (:2277-2294) */
            if (reverse && score - o >= 0) {
                const int ps = slot_of(slot, o, w.ring);
                for (int kind = WI; kind <= WD; kind++) {
                    const int p = *wf(&w, kind, ps, d);
                    if (p >= 0 && p < qlen && k + p >= 0 && k + p < tlen && p > *m_cur) *m_cur = p;
                }
            }
            /* gap extension   (:2296-2311) */
            if (score - e >= 0) {
                const int ps = slot_of(slot, e, w.ring);
                if (d > 0) {
                    const int p = *wf(&w, WD, ps, d - 1);
                    if (p != NONE && k + p < tlen && p >= *d_cur) *d_cur = p;
                }
                if (d < w.nd - 1) {
                    const int p = *wf(&w, WI, ps, d + 1);
                    if (p != NONE && p + 1 < qlen && k + p + 1 < tlen && k + p + 1 >= 0 && p + 1 >= *i_cur) *i_cur = p + 1;
                }
            }
        }
    }
    /* the score budget is spent: furthest truth index over everything still in the ring   (:2316-2331) */
    result = 0;
    for (int kind = WM; kind < NW; kind++)
        for (int s = 0; s < w.ring; s++)
            for (int d = 0; d < w.nd; d++) {
                const int q = *wf(&w, kind, s, d), k = d + 1 - qlen;
                if (q >= 0 && q < qlen && k + q >= 0 && k + q < tlen && k + q > result) result = k + q;
            }
done:
    free(w.v);
    return result;
}


/* Restatement of the reference's wf_swg_align (/root/reference/src/dist.cpp:1510-1652) as far as its
 * score goes: global affine-gap (Gotoh) wavefront alignment, mismatch x, gap open o, gap extension e;
 * returns the first score at which some wavefront reaches the last base of both strings.  The
 * cluster-growing step uses only this score (src/cluster.cpp:1040-1046: the budget of the reach
 * searches); the backtrack pointers the reference also records are what `--distance` consumes
 *  and are not restated here.  Same diagonal / wavefront conventions as above; every
 * wavefront of a new score starts empty (:1583-1587).
 *
 * Not the textbook Gotoh distance: the search starts BEFORE the first base on the main diagonal and a
 * gap cannot be opened there (the insertion test needs a truth index >= 0, :1613-1617, and the
 * deletion moves to a diagonal whose query index -1 can never be extended into query[0] with truth[0]),
 * so the first bases of the two strings are always paired: "CA" vs "A" scores x + (o+e), not o+e.
 * The reference's callers always pass strings that start with the same flanking reference base
 * (src/cluster.cpp:1037-1041), where the two coincide; measured against a textbook implementation the
 * two differ in about one of six random pairs.  Anything that replaces this function must reproduce it
 * or keep that precondition.                                                                         */
int vdo_swg_score(const char *query, int qlen, const char *truth, int tlen, int x, int o, int e) {
    wf_t w;
    w.nd = qlen + tlen - 1;
    w.ring = (x > o + e ? x : o + e) + 1;
    const size_t total = (size_t)NW * w.ring * w.nd;
    w.v = (int *)malloc(sizeof(int) * (total ? total : 1));
    for (size_t i = 0; i < total; i++) w.v[i] = NONE;
    int score = 0, slot = 0;
    *wf(&w, WM, slot, qlen - 1) = -1;                                   /* :1528 */
    for (;;) {
        for (int kind = WI; kind <= WD; kind++)                           /* gaps are left for free (:1533-1547) */
            for (int d = 0; d < w.nd; d++) {
                const int q = *wf(&w, kind, slot, d), k = d + 1 - qlen;
                if (q >= 0 && q < qlen && k + q >= 0 && k + q < tlen && q >= *wf(&w, WM, slot, d))
                    *wf(&w, WM, slot, d) = q;
            }
        for (int d = 0; d < w.nd; d++) {                                  /* extension (:1550-1568) */
            int q = *wf(&w, WM, slot, d);
            const int k = d + 1 - qlen;
            while (q != NONE && k + q >= -1 && q < qlen - 1 && k + q < tlen - 1 && query[q + 1] == truth[k + q + 1]) q++;
            *wf(&w, WM, slot, d) = q;
            if (q == qlen - 1 && q + k == tlen - 1) { free(w.v); return score; }
        }
        score++;
        slot = slot + 1 == w.ring ? 0 : slot + 1;
        for (int kind = WM; kind < NW; kind++)
            for (int d = 0; d < w.nd; d++) *wf(&w, kind, slot, d) = NONE;
        for (int d = 0; d < w.nd; d++) {
            const int k = d + 1 - qlen;
            int *m_cur = wf(&w, WM, slot, d), *i_cur = wf(&w, WI, slot, d), *d_cur = wf(&w, WD, slot, d);
            if (score - x >= 0) {                                         /* :1592-1600 */
                const int p = *wf(&w, WM, slot_of(slot, x, w.ring), d);
                if (p != NONE && p + 1 < qlen && k + p + 1 < tlen && p + 1 >= *m_cur) *m_cur = p + 1;
            }
            if (score - (o + e) >= 0) {                                   /* :1602-1625 */
                const int ps = slot_of(slot, o + e, w.ring);
                if (d > 0) {
                    const int p = *wf(&w, WM, ps, d - 1);
                    if (p != NONE && k + p < tlen && p >= *d_cur) *d_cur = p;
                }
                if (d < w.nd - 1) {
                    const int p = *wf(&w, WM, ps, d + 1);
                    if (p != NONE && p + 1 < qlen && k + p + 1 < tlen && k + p + 1 >= 0 && p + 1 >= *i_cur) *i_cur = p + 1;
                }
            }
            if (score - e >= 0) {                                         /* :1627-1650 */
                const int ps = slot_of(slot, e, w.ring);
                if (d > 0) {
                    const int p = *wf(&w, WD, ps, d - 1);
                    if (p != NONE && k + p < tlen && p >= *d_cur) *d_cur = p;
                }
                if (d < w.nd - 1) {
                    const int p = *wf(&w, WI, ps, d + 1);
                    if (p != NONE && p + 1 < qlen && k + p + 1 < tlen && k + p + 1 >= 0 && p + 1 >= *i_cur) *i_cur = p + 1;
                }
            }
        }
    }
}


/* Restatement of the reference's wf_swg_align WITH its predecessor flags and of wf_swg_backtrack
 * (/root/reference/src/dist.cpp:1510-1652, :2625-2757): the affine-gap alignment the optional
 * `--distance` pass runs per supercluster and haplotype (edits_wrapper, :1908-2077).
 * Every score keeps its three wavefronts and one flag byte per diagonal: how the cell was entered
 * (F_INS / F_DEL: by leaving or extending that gap, F_SUB: by a substitution or by opening a gap,
 * F_MAT: the origin).  The walk back from the last cell prefers, on the M wavefront, leaving an
 * insertion over leaving a deletion over a substitution (:2662-2700) and, inside a gap, extending over
 * opening (:2716-2745).  cigar[] has |query|+|truth| entries, filled from the back as the reference
 * does: two entries per match / substitution, one per inserted / deleted base, zeros in front.
 * Returns the score.                                                                                */
enum { F_INS = 1, F_DEL = 2, F_MAT = 4, F_SUB = 8 };

typedef struct { int *off[NW]; unsigned char *flag[NW]; } wave_t;

int vdo_swg_cigar(const char *query, int qlen, const char *truth, int tlen, int x, int o, int e, int *cigar) {
    const int nd = qlen + tlen - 1;
    int cap = 16, score = 0;
    wave_t *W = (wave_t *)malloc(sizeof(wave_t) * cap);
#define NEW_WAVE(sidx) do { for (int k_ = 0; k_ < NW; k_++) { \
        W[sidx].off[k_] = (int *)malloc(sizeof(int) * nd); W[sidx].flag[k_] = (unsigned char *)calloc(nd, 1); \
        for (int d_ = 0; d_ < nd; d_++) W[sidx].off[k_][d_] = NONE; } } while (0)
    NEW_WAVE(0);
    W[0].off[WM][qlen - 1] = -1;                                        /* :1528-1529 */
    W[0].flag[WM][qlen - 1] = F_MAT;
    for (;;) {
        wave_t *c = &W[score];
        for (int kind = WI; kind <= WD; kind++)                           /* :1533-1547 */
            for (int d = 0; d < nd; d++) {
                const int q = c->off[kind][d], k = d + 1 - qlen;
                if (q >= 0 && q < qlen && k + q >= 0 && k + q < tlen && q >= c->off[WM][d]) {
                    c->off[WM][d] = q;
                    c->flag[WM][d] |= (kind == WI) ? F_INS : F_DEL;
                }
            }
        int done = 0;
        for (int d = 0; d < nd && !done; d++) {                           /* :1550-1568 */
            int q = c->off[WM][d];
            const int k = d + 1 - qlen;
            while (q != NONE && k + q >= -1 && q < qlen - 1 && k + q < tlen - 1 && query[q + 1] == truth[k + q + 1]) q++;
            c->off[WM][d] = q;
            if (q == qlen - 1 && q + k == tlen - 1) done = 1;
        }
        if (done) break;
        score++;
        if (score == cap) { cap *= 2; W = (wave_t *)realloc(W, sizeof(wave_t) * cap); }
        NEW_WAVE(score);
        c = &W[score];
        for (int d = 0; d < nd; d++) {
            const int k = d + 1 - qlen;
            if (score - x >= 0) {                                         /* :1592-1600 */
                const int p = W[score - x].off[WM][d];
                if (p != NONE && p + 1 < qlen && k + p + 1 < tlen && p + 1 >= c->off[WM][d]) { c->off[WM][d] = p + 1; c->flag[WM][d] |= F_SUB; }
            }
            if (score - (o + e) >= 0) {                                   /* :1602-1625 */
                const wave_t *pw = &W[score - (o + e)];
                if (d > 0) {
                    const int p = pw->off[WM][d - 1];
                    if (p != NONE && k + p < tlen && p >= c->off[WD][d]) { c->off[WD][d] = p; c->flag[WD][d] |= F_SUB; }
                }
                if (d < nd - 1) {
                    const int p = pw->off[WM][d + 1];
                    if (p != NONE && p + 1 < qlen && k + p + 1 < tlen && k + p + 1 >= 0 && p + 1 >= c->off[WI][d]) { c->off[WI][d] = p + 1; c->flag[WI][d] |= F_SUB; }
                }
            }
            if (score - e >= 0) {                                         /* :1627-1650 */
                const wave_t *pw = &W[score - e];
                if (d > 0) {
                    const int p = pw->off[WD][d - 1];
                    if (p != NONE && k + p < tlen && p >= c->off[WD][d]) { c->off[WD][d] = p; c->flag[WD][d] |= F_DEL; }
                }
                if (d < nd - 1) {
                    const int p = pw->off[WI][d + 1];
                    if (p != NONE && p + 1 < qlen && k + p + 1 < tlen && k + p + 1 >= 0 && p + 1 >= c->off[WI][d]) { c->off[WI][d] = p + 1; c->flag[WI][d] |= F_INS; }
                }
            }
        }
    }
    /* ---- walk back (:2648-2755) ---- */
    const int final_score = score;
    for (int i = 0; i < qlen + tlen; i++) cigar[i] = 0;
    int cp = qlen + tlen - 1, kind = WM, qi = qlen - 1, ti = tlen - 1, s = score, failed = 0;
    while ((qi >= 0 || ti >= 0) && !failed) {
        if (s < 0) { failed = 1; break; }
        const int d = qlen - 1 + (ti - qi);
        if (kind == WM) {
            const int f = W[s].flag[WM][d];
            if (f & (F_INS | F_DEL)) {                                    /* a gap was left here for free */
                const int g = (f & F_INS) ? WI : WD;
                const int stop = W[s].off[g][d];
                while (qi > stop) { cigar[cp--] = F_MAT; cigar[cp--] = F_MAT; qi--; ti--; if (qi < 0 || ti < 0) { failed = 1; break; } }
                kind = g;
            } else if (f & F_SUB) {
                if (s - x < 0) { failed = 1; break; }
                const int stop = W[s - x].off[WM][d] + 1;
                while (qi > stop) { cigar[cp--] = F_MAT; cigar[cp--] = F_MAT; qi--; ti--; if (qi < 0 || ti < 0) { failed = 1; break; } }
                if (failed) break;
                cigar[cp--] = F_SUB; cigar[cp--] = F_SUB; qi--; ti--;
                s -= x;
            } else if (f & F_MAT) {
                while (qi >= 0 && ti >= 0) { cigar[cp--] = F_MAT; cigar[cp--] = F_MAT; qi--; ti--; }
                if (qi >= 0 || ti >= 0) failed = 1;
            } else failed = 1;
        } else if (kind == WI) {
            const int f = W[s].flag[WI][d];
            if (f & F_INS) { cigar[cp--] = F_INS; qi--; s -= e; }
            else if (f & F_SUB) { cigar[cp--] = F_INS; qi--; kind = WM; s -= o + e; }
            else failed = 1;
        } else {
            const int f = W[s].flag[WD][d];
            if (f & F_DEL) { cigar[cp--] = F_DEL; ti--; s -= e; }
            else if (f & F_SUB) { cigar[cp--] = F_DEL; ti--; kind = WM; s -= o + e; }
            else failed = 1;
        }
        if (!(qi == -1 && ti == -1) && (qi < 0 || ti < 0)) failed = 1;
    }
    for (int i = 0; i <= final_score; i++)
        for (int k_ = 0; k_ < NW; k_++) { free(W[i].off[k_]); free(W[i].flag[k_]); }
    free(W);
#undef NEW_WAVE
    return failed ? -1 : final_score;                                    /* -1: the reference would ERROR() and exit */
}
