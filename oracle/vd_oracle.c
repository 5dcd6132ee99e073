/* TEST INFRASTRUCTURE ONLY — CPU restatement ("port" oracle) of the vcfdist v2.6.4
 * precision/recall hot path, in plain C.  Nothing in the product path may call this:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg use it, and only
 * as the checker for the CUDA path.
 *
 * Parity status: PINNED.  This restatement is validated bit-for-bit against the
 * reference's own object code (oracle/_ref/libvdref*.so, built from the unmodified
 * sources under /root/reference/src by oracle/Makefile) by tests/test_oracle_vs_ref.py
 * on the bundled demo and on seeded adversarial batches; golden outputs of the
 * reference are committed under tests/golden/.
 *
 * The reference explores the two-plane alignment graph with a Dijkstra-like BFS over
 * std::unordered_set (src/dist.cpp:251-443, 486-834); this file restates it as the
 * equivalent dense dynamic program, one column of the truth haplotype
 * at a time.  Each function cites the reference lines it follows.
 *
 * Tie rule: where the reference keeps the LAST writer of swap_pred_maps
 * (src/dist.cpp:347, :376; order = std::unordered_set iteration order), this oracle —
 * like the CUDA path — keeps the source with the larger row index and raises
 * VD_ST_TIE when the choice was ambiguous on a reachable cell.  libvdrefB.so is the
 * reference with exactly that rule patched in.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "vcfdist_b200.h"

/* src/defs.h:110-129 */
#define PTR_INS 1
#define PTR_DEL 2
#define PTR_MAT 4
#define PTR_SUB 8
#define PTR_SWP 16
#define F_VARIANT 1
#define F_VAR_BEG 2
#define F_VAR_END 4
#define F_INS_LOC 8

#define INF 0x3fffffff

typedef struct {
    int len;          /* haplotype string length                    */
    uint8_t *str;     /* haplotype bases                            */
    int *ptr;         /* hap -> ref pointers   (query_ptrs[PTRS])   */
    uint8_t *flg;     /* hap -> ref flags      (query_ptrs[FLAGS])  */
    int rlen;         /* ref string length built alongside          */
    int *rptr;        /* ref -> hap pointers   (ref_ptrs[PTRS])     */
    uint8_t *rflg;    /* ref -> hap flags      (ref_ptrs[FLAGS])    */
} hap_t;

static void hap_free(hap_t *h) {
    free(h->str); free(h->ptr); free(h->flg); free(h->rptr); free(h->rflg);
    memset(h, 0, sizeof(*h));
}

/* generate_ptrs_strs, src/dist.cpp:145-242.  Returns 0, or -1 on input the reference
 * itself cannot process (variant left of the cursor, or running past the window).   */
static int expand_hap(const vd_batch_in *in, int sc, int h, hap_t *out) {
    int64_t r0 = in->ref_off[sc];
    int win = (int)(in->ref_off[sc + 1] - r0);          /* end_pos - beg_pos + 1 */
    const uint8_t *fa = in->ref_seq + r0;
    int64_t vb = in->var_off[4 * sc + h], ve = in->var_off[4 * sc + h + 1];

    int cap = win;
    for (int64_t v = vb; v < ve; v++)
        cap += (int)(in->alt_off[v + 1] - in->alt_off[v]);
    out->str = (uint8_t *)malloc(cap + 1);
    out->ptr = (int *)malloc(sizeof(int) * (cap + 1));
    out->flg = (uint8_t *)malloc(cap + 1);
    out->rptr = (int *)malloc(sizeof(int) * (win + 1));
    out->rflg = (uint8_t *)malloc(win + 1);
    int Q = 0, R = 0;        /* query_str.size(), ref_str.size() */
    int64_t v = vb;
    int ref_pos = 0;
    while (ref_pos < win) {                              /* :163, ref_pos <= end_pos */
        if (v < ve && ref_pos == in->var_pos[v]) {       /* :165-166 */
            int alen = (int)(in->alt_off[v + 1] - in->alt_off[v]);
            const uint8_t *alt = in->alt_seq + in->alt_off[v];
            int rl = in->var_rlen[v];
            switch (in->var_type[v]) {
            case VD_TYPE_INS:                            /* :169-178 */
                if (alen < 1) return -1;
                for (int k = 0; k < alen; k++) {
                    out->ptr[Q + k] = R - 1;
                    out->flg[Q + k] = F_VARIANT;
                    out->str[Q + k] = alt[k];
                }
                out->flg[Q + alen - 1] |= F_VAR_END;
                out->flg[Q] |= F_VAR_BEG | F_INS_LOC;
                Q += alen;
                break;
            case VD_TYPE_DEL:                            /* :179-189 */
                if (rl < 1 || R + rl > win) return -1;
                for (int k = 0; k < rl; k++) {
                    out->rptr[R + k] = Q - 1;
                    out->rflg[R + k] = F_VARIANT;
                }
                out->rflg[R + rl - 1] |= F_VAR_END;
                out->rflg[R] |= F_VAR_BEG;
                R += rl;
                ref_pos += rl;
                break;
            case VD_TYPE_SUB:                            /* :190-198 */
                if (alen != 1 || rl != 1) return -1;
                out->rptr[R] = Q;
                out->rflg[R] = F_VARIANT | F_VAR_BEG | F_VAR_END;
                out->ptr[Q] = R;
                out->flg[Q] = F_VARIANT | F_VAR_BEG | F_VAR_END;
                out->str[Q] = alt[0];
                R++; Q++; ref_pos++;
                break;
            default:
                return -1;                               /* :199-201 */
            }
            v++;                                         /* :204 */
        } else {                                         /* :206-235 */
            int ref_end = (v < ve) ? in->var_pos[v] : win;
            if (ref_end < ref_pos || ref_end > win) return -1;
            int n = ref_end - ref_pos;
            for (int k = 0; k < n; k++) {
                out->ptr[Q + k] = R + k;
                out->flg[Q + k] = 0;
                out->rptr[R + k] = Q + k;
                out->rflg[R + k] = 0;
                out->str[Q + k] = fa[ref_pos + k];
            }
            Q += n; R += n; ref_pos = ref_end;
        }
    }
    if (R != win) return -1;
    out->len = Q;
    out->rlen = R;
    return 0;
}

/* plain unit-cost global edit distance; wf_ed (src/dist.cpp:1406-1506) computes
 * exactly this (only `s` is consumed by its caller, :1196-1199).                     */
static int lev(const uint8_t *a, int m, const uint8_t *b, int n) {
    if (!m) return n;                                    /* :1417 */
    if (!n) return m;                                    /* :1418 */
    int *row = (int *)malloc(sizeof(int) * (n + 1));
    for (int j = 0; j <= n; j++) row[j] = j;
    for (int i = 1; i <= m; i++) {
        int diag = row[0];
        row[0] = i;
        for (int j = 1; j <= n; j++) {
            int up = row[j];
            int best = diag + (a[i - 1] != b[j - 1]);
            if (up + 1 < best) best = up + 1;
            if (row[j - 1] + 1 < best) best = row[j - 1] + 1;
            row[j] = best;
            diag = up;
        }
    }
    int r = row[n];
    free(row);
    return r;
}

typedef struct { int hi, qri, ti; } cell_t;   /* hi: 0 QUERY plane, 1 REF plane */

/* swap-source lists: for destination row a of plane P, the rows b of the other plane
 * with src(b) and dest(b) == a   (src/dist.cpp:335-337, :364-366)                    */
typedef struct { int *off; int *src; } csr_t;

static void build_csr(const int *ptr, const uint8_t *flg, int nsrc, int ndst, csr_t *c) {
    c->off = (int *)calloc(ndst + 2, sizeof(int));
    c->src = (int *)malloc(sizeof(int) * (nsrc + 1));
    for (int b = 0; b < nsrc; b++) {
        int ok = !(flg[b] & F_VARIANT) || (flg[b] & F_VAR_END);
        int d = ptr[b] + 1;
        if (ok && d >= 0 && d < ndst) c->off[d + 1]++;
    }
    for (int a = 0; a < ndst; a++) c->off[a + 1] += c->off[a];
    int *fill = (int *)malloc(sizeof(int) * (ndst + 1));
    memcpy(fill, c->off, sizeof(int) * (ndst + 1));
    for (int b = 0; b < nsrc; b++) {                     /* ascending b => ascending lists */
        int ok = !(flg[b] & F_VARIANT) || (flg[b] & F_VAR_END);
        int d = ptr[b] + 1;
        if (ok && d >= 0 && d < ndst) c->src[fill[d]++] = b;
    }
    free(fill);
}

/* One alignment: query hap `q` (with its ref<->query maps) against truth hap `t`.
 * rseq = REF-plane string.  Writes the integer results for this alignment.           */
static void align_one(const vd_batch_in *in, vd_batch_out *out, int64_t n_var,
                      int sc, int ai, const hap_t *q, const hap_t *t, const uint8_t *rseq) {
    const int Lq = q->len, Lr = q->rlen, Lt = t->len;
    const int N = Lq + Lr;                               /* rows: Q plane then R plane */
    uint32_t status = 0;
    const size_t cells = (size_t)N * Lt;

    uint8_t *F = (uint8_t *)calloc(cells, 1);            /* forward flags, [c][row]      */
    int *SRC = (int *)malloc(sizeof(int) * cells);       /* chosen swap source row       */
    uint8_t *TIE = (uint8_t *)calloc(cells, 1);
    int *Dp = (int *)malloc(sizeof(int) * N), *Dc = (int *)malloc(sizeof(int) * N);

    csr_t toR, toQ;                                      /* sources in Q for dest rows of R, and v.v. */
    build_csr(q->ptr, q->flg, Lq, Lr, &toR);
    build_csr(q->rptr, q->rflg, Lr, Lq, &toQ);

    /* ---- forward pass: calc_prec_recall_aln, src/dist.cpp:251-443 ---- */
    for (int c = 0; c < Lt; c++) {
        int tok = c > 0 && (!(t->flg[c - 1] & F_VARIANT) || (t->flg[c - 1] & F_VAR_END)); /* :338-339 */
        for (int row = 0; row < N; row++) {
            int P = row >= Lq, a = P ? row - Lq : row;
            size_t idx = (size_t)c * N + row;
            if (a == 0 && c == 0) { Dc[row] = 0; F[idx] = PTR_MAT; continue; }   /* :299-305 */
            int m = (P ? rseq[a] : q->str[a]) == t->str[c];
            int diag = INF, ins = INF, del = INF, swp = INF, swsrc = -1, nsw = 0;
            if (a > 0 && c > 0) diag = Dp[row - 1] + (m ? 0 : 1);         /* :324-332, :415-422 */
            if (a > 0) ins = Dc[row - 1] + 1;                              /* :397-404 */
            if (c > 0) del = Dp[row] + 1;                                  /* :406-413 */
            if (tok && m) {                                                /* :334-349, :363-378 */
                const csr_t *cs = P ? &toR : &toQ;
                int obase = P ? 0 : Lq;                  /* row offset of the other plane */
                for (int k = cs->off[a]; k < cs->off[a + 1]; k++) {
                    int v = Dp[obase + cs->src[k]];
                    if (v < swp) { swp = v; swsrc = cs->src[k]; nsw = 1; }
                    else if (v == swp) { swsrc = cs->src[k]; nsw++; }      /* keep larger row */
                }
            }
            int d = diag;
            if (ins < d) d = ins;
            if (del < d) d = del;
            if (swp < d) d = swp;
            uint8_t f = 0;
            if (diag == d && d < INF) f |= m ? PTR_MAT : PTR_SUB;
            if (ins == d && d < INF) f |= PTR_INS;
            if (del == d && d < INF) f |= PTR_DEL;
            if (swp == d && d < INF) { f |= PTR_SWP; SRC[idx] = swsrc; TIE[idx] = nsw > 1; }
            Dc[row] = d;
            F[idx] = f;
        }
        int *tmp = Dp; Dp = Dc; Dc = tmp;
    }
    /* Dp now holds the last column.  :390-391, :436-440 */
    int dq = Dp[Lq - 1], dr = Dp[Lq + Lr - 1];
    int s = dq < dr ? dq : dr;
    int end_plane = (dq == s) ? 0 : 1;                   /* prefer QUERY */
    out->aln_score[4 * sc + ai] = s;
    out->aln_end_plane[4 * sc + ai] = (uint8_t)end_plane;

    /* ---- backward pass: calc_prec_recall_path, src/dist.cpp:486-834 ---- */
    int *T = (int *)malloc(sizeof(int) * cells);
    uint8_t *PF = (uint8_t *)calloc(cells, 1);
    for (size_t i = 0; i < cells; i++) T[i] = -1;        /* :527-530 */
    {
        int erow = end_plane ? Lq + Lr - 1 : Lq - 1;
        size_t e = (size_t)(Lt - 1) * N + erow;
        T[e] = 0; PF[e] = PTR_MAT;                       /* :543-545 */
    }
#define RELAX(yidx, val, ty) do { \
        if ((val) > T[yidx]) { T[yidx] = (val); PF[yidx] = (ty); } \
        else if ((val) == T[yidx]) PF[yidx] |= (ty); } while (0)
    for (int c = Lt - 1; c >= 0; c--) {
        for (int row = N - 1; row >= 0; row--) {
            size_t x = (size_t)c * N + row;
            if (T[x] < 0) continue;
            int P = row >= Lq, a = P ? row - Lq : row;
            uint8_t f = F[x];
            int tp = 0;                                  /* :572-574, :656-658, :709-711, :749-751 */
            if (!P && a > 0)
                tp = (q->ptr[a] != q->ptr[a - 1] + 1) || (q->flg[a] & F_VAR_BEG);
            if ((f & PTR_MAT) && a > 0 && c > 0) {       /* :556-595 */
                size_t y = (size_t)(c - 1) * N + row - 1;
                RELAX(y, T[x] + tp, PTR_MAT);
            }
            if ((f & PTR_SWP) && a > 0 && c > 0) {       /* :598-679 */
                uint8_t of = P ? q->rflg[a] : q->flg[a];
                if (!(of & F_VARIANT) || (of & F_VAR_BEG)) {
                    int zrow = (P ? 0 : Lq) + SRC[x];
                    size_t z = (size_t)(c - 1) * N + zrow;
                    if (TIE[x]) status |= VD_ST_TIE;
                    RELAX(z, T[x] + (P ? 0 : tp), PTR_SWP);   /* is_tp false on REF (:614) */
                }
            }
            if ((f & PTR_SUB) && a > 0 && c > 0) {       /* :692-731 */
                size_t y = (size_t)(c - 1) * N + row - 1;
                RELAX(y, T[x] + tp, PTR_SUB);
            }
            if ((f & PTR_INS) && a > 0) {                /* :734-771 */
                size_t y = (size_t)c * N + row - 1;
                RELAX(y, T[x] + tp, PTR_INS);
            }
            if ((f & PTR_DEL) && c > 0) {                /* :774-804 */
                size_t y = (size_t)(c - 1) * N + row;
                RELAX(y, T[x], PTR_DEL);
            }
        }
    }
#undef RELAX
    int beg_plane = (T[0] >= 0) ? 0 : 1;                 /* :811-814 (row 0 = (Q,0,0)) */
    out->aln_beg_plane[4 * sc + ai] = (uint8_t)beg_plane;

    /* ---- walk: get_prec_recall_path_sync, src/dist.cpp:842-999 ---- */
    int maxpath = N + Lt + 4;
    cell_t *path = (cell_t *)malloc(sizeof(cell_t) * maxpath);
    uint8_t *sync = (uint8_t *)malloc(maxpath + 1), *edit = (uint8_t *)malloc(maxpath + 1);
    int np = 0, ns = 0;                                  /* path entries, sync/edit entries */
    uint8_t *ref_has_ins = (uint8_t *)calloc(Lr + 1, 1); /* :886-894 */
    for (int j = 0; j < Lq; j++)
        if ((q->flg[j] & F_INS_LOC) && q->ptr[j] >= 0 && q->ptr[j] < Lr) ref_has_ins[q->ptr[j]] = 1;
    for (int j = 0; j < Lt; j++)
        if ((t->flg[j] & F_INS_LOC) && t->ptr[j] >= 0 && t->ptr[j] < Lr) ref_has_ins[t->ptr[j]] = 1;
    {
        int hi = beg_plane, qri = 0, ti = 0;
        sync[ns] = 1; edit[ns] = 0; ns++;                /* :897-898 */
        path[np].hi = hi; path[np].qri = qri; path[np].ti = ti; np++;
        while ((hi == 1 && qri < Lr - 1) || (hi == 0 && qri < Lq - 1) || ti < Lt - 1) {  /* :905 */
            uint8_t pf = PF[(size_t)ti * N + (hi ? Lq + qri : qri)];
            int ty;
            if (hi == 1 && (pf & PTR_SWP)) {             /* :907-912 */
                ty = PTR_SWP; hi = 0; qri = q->rptr[qri]; qri++; ti++; edit[ns] = 0;
            } else if (pf & PTR_MAT) { ty = PTR_MAT; qri++; ti++; edit[ns] = 0;   /* :914-916 */
            } else if (pf & PTR_SUB) { ty = PTR_SUB; qri++; ti++; edit[ns] = 1;   /* :918-920 */
            } else if (pf & PTR_INS) { ty = PTR_INS; qri++; edit[ns] = 1;         /* :922-924 */
            } else if (pf & PTR_DEL) { ty = PTR_DEL; ti++; edit[ns] = 1;          /* :926-928 */
            } else if (hi == 0 && (pf & PTR_SWP)) {      /* :930-934 */
                ty = PTR_SWP; hi = 1; qri = q->ptr[qri]; qri++; ti++; edit[ns] = 0;
            } else { status |= VD_ST_ERR_NO_POINTER; break; }                     /* :936-939 */
            /* edits[] was pushed above; sync[] is pushed below unless we break (:941) */
            if ((hi == 0 && qri >= Lq) || (hi == 1 && qri >= Lr) || ti >= Lt) { ns++; break; }
            if (np >= maxpath - 1) { status |= VD_ST_ERR_NO_POINTER; ns++; break; }
            path[np].hi = hi; path[np].qri = qri; path[np].ti = ti; np++;
            int in_truth_var = t->flg[ti] & F_VARIANT;                            /* :949-951 */
            if (ty & (PTR_MAT | PTR_SWP | PTR_SUB | PTR_DEL))
                in_truth_var = in_truth_var && !(t->flg[ti] & F_VAR_BEG);
            int in_query_var = (hi == 1) ? 0 : (q->flg[qri] & F_VARIANT);         /* :953-956 */
            if (hi == 0 && (ty & (PTR_MAT | PTR_SWP | PTR_SUB | PTR_DEL)))
                in_query_var = in_query_var && !(q->flg[qri] & F_VAR_BEG);
            int tref = t->ptr[ti];
            int qref = (hi == 1) ? qri : q->ptr[qri];
            int is_ins_loc = (tref >= 0 && tref < Lr && ref_has_ins[tref]) ||     /* :958-960 */
                             (qref >= 0 && qref < Lr && ref_has_ins[qref]);
            int is_sync = !in_truth_var && !in_query_var && !is_ins_loc &&        /* :964-967 */
                          tref == qref && (ty & (PTR_MAT | PTR_SWP | PTR_SUB));
            sync[ns] = (uint8_t)(is_sync ? 1 : 0);
            ns++;
        }
        /* when the break at :941 fires, edits has one more entry than sync; the final
         * pushes (:995-997) then append to both, as in the reference                  */
        if (status & VD_ST_ERR_NO_POINTER) goto done;
    }
    {
        /* ---- credit: calc_prec_recall, src/dist.cpp:1005-1401 ---- */
        /* The reference keeps sync and edits as separate vectors: one edits entry per
         * loop iteration, one sync entry per pushed path cell.  They advance together
         * unless the :941 break fired, in which case edits is one entry longer.       */
        int broke = (ns == np + 1);
        /* build the reference's vectors */
        int nsync = np + 1, nedit = np + 1 + broke;
        uint8_t *sv = (uint8_t *)malloc(nsync + 1), *ev = (uint8_t *)malloc(nedit + 1);
        for (int k = 0; k < np; k++) sv[k] = sync[k];
        sv[np] = 1;                                      /* :995 */
        for (int k = 0; k < np + broke; k++) ev[k] = edit[k];
        ev[np + broke] = 0;                              /* :997 */

        int swap = (ai == 1 || ai == 2);                 /* :1037 */
        int qh = ai >> 1, th = 2 + (ai & 1);
        int64_t qb = in->var_off[4 * sc + qh], qe = in->var_off[4 * sc + qh + 1];
        int64_t tb = in->var_off[4 * sc + th], te = in->var_off[4 * sc + th + 1];
        uint8_t *asg = out->assigned + (int64_t)swap * n_var;
        int32_t *sg = out->sync_group + (int64_t)swap * n_var;
        int32_t *red = out->ref_ed + (int64_t)swap * n_var;
        int32_t *qed = out->query_ed + (int64_t)swap * n_var;
        float *cq = out->callq + (int64_t)swap * n_var;

        int sync_group = 0;                              /* :1059 */
        int hi = end_plane;                              /* :1061 */
        int qri = (hi == 0 ? Lq : Lr) - 1;               /* :1062-1065 */
        int prev_hi = hi, prev_qri = qri;
        int prev_sync_ref_idx = Lr;                      /* :1066-1067 */
        int ti = Lt - 1, prev_ti = ti;
        int prev_sync_truth_idx = Lt;                    /* :1071-1072 */
        int query_ed = 0;
        int64_t qvp = qe - 1, prev_qvp = qvp;            /* :1074-1077 */
        int q_pos = (qvp >= qb) ? in->var_pos[qvp] : 0;
        int64_t tvp = te - 1, prev_tvp = tvp;            /* :1078-1081 */
        int t_pos = (tvp >= tb) ? in->var_pos[tvp] : 0;
        int sync_idx = nsync - 1;                        /* :1082 */
        (void)qri; (void)ti;

        while (sync_idx >= 0) {                          /* :1136 */
            int query_ref_pos = (prev_hi == 1) ? prev_qri : q->ptr[prev_qri];     /* :1139-1144 */
            while (query_ref_pos < q_pos && qvp >= qb) {                          /* :1147 */
                if (hi == 1) {                           /* :1157-1168 */
                    asg[qvp] = VD_ASSIGN_REF_FP;
                    sg[qvp] = sync_group++;
                    red[qvp] = 0; qed[qvp] = 0;
                    cq[qvp] = in->var_qual[qvp];
                }
                qvp--;
                q_pos = (qvp < qb) ? -1 : in->var_pos[qvp];                       /* :1175-1176 */
            }
            int truth_ref_pos = t->ptr[prev_ti];                                  /* :1180 */
            while (truth_ref_pos < t_pos && tvp >= tb) {                          /* :1181-1187 */
                tvp--;
                t_pos = (tvp < tb) ? -1 : in->var_pos[tvp];
            }
            if (sv[sync_idx]) {                          /* :1190 */
                int sync_ref_idx = ((prev_hi == 1) ? prev_qri : q->ptr[prev_qri]) + 1;   /* :1194 */
                int sync_truth_idx = prev_ti + 1;                                 /* :1195 */
                int rn = prev_sync_ref_idx - sync_ref_idx;                        /* std::string::substr */
                if (rn < 0 || sync_ref_idx + rn > Lr) rn = Lr - sync_ref_idx;
                int tn = prev_sync_truth_idx - sync_truth_idx;
                if (tn < 0 || sync_truth_idx + tn > Lt) tn = Lt - sync_truth_idx;
                int ref_ed = lev(rseq + sync_ref_idx, rn, t->str + sync_truth_idx, tn);   /* :1197-1199 */
                if (prev_tvp == tvp && ref_ed != 0) status |= VD_ST_WARN_REFED_NOTRUTH;   /* :1203 */
                if (prev_qvp == qvp && query_ed != ref_ed) status |= VD_ST_WARN_QED_NOQUERY; /* :1207 */
                if (query_ed > ref_ed) status |= VD_ST_WARN_QED_GT_REFED;         /* :1211 */
                if (ref_ed == 0 && tvp != prev_tvp) {                             /* :1219-1223 */
                    status |= VD_ST_WARN_ZERO_REFED;
                    ref_ed = 1;
                }
                float callq = in->max_qual;                                       /* :1284-1288 */
                for (int64_t v = prev_qvp; v > qvp; v--)
                    if (in->var_qual[v] < callq) callq = in->var_qual[v];
                for (int64_t v = prev_qvp; v > qvp; v--) {                        /* :1291-1322 */
                    if (asg[v] == VD_ASSIGN_NONE) {
                        asg[v] = VD_ASSIGN_SYNC;
                        sg[v] = sync_group; red[v] = ref_ed; qed[v] = query_ed; cq[v] = callq;
                    }
                }
                for (int64_t v = prev_tvp; v > tvp; v--) {                        /* :1325-1353 */
                    asg[v] = VD_ASSIGN_SYNC;
                    sg[v] = sync_group; red[v] = ref_ed; qed[v] = query_ed; cq[v] = callq;
                }
                if (qvp != prev_qvp || tvp != prev_tvp) sync_group++;             /* :1364-1367 */
                prev_qvp = qvp; prev_tvp = tvp;
                prev_sync_ref_idx = sync_ref_idx;
                prev_sync_truth_idx = sync_truth_idx;
                query_ed = 0;
            }
            query_ed += ev[sync_idx];                    /* :1382 */
            sync_idx--;
            if (sync_idx < 0) break;
            hi = prev_hi;                                /* :1387-1392 */
            prev_qri = path[sync_idx].qri;
            prev_ti = path[sync_idx].ti;
            prev_hi = path[sync_idx].hi;
            q_pos = (qvp < qb) ? -1 : in->var_pos[qvp];  /* :1395-1398 */
            t_pos = (tvp < tb) ? -1 : in->var_pos[tvp];
        }
        free(sv); free(ev);
    }
done:
    out->status[4 * sc + ai] = status;
    free(F); free(SRC); free(TIE); free(Dp); free(Dc); free(T); free(PF);
    free(path); free(sync); free(edit); free(ref_has_ins);
    free(toR.off); free(toR.src); free(toQ.off); free(toQ.src);
}

/* precision_recall_wrapper, src/dist.cpp:1731-1904, for every supercluster of the batch.
 * `out` arrays must be sized as documented in include/vcfdist_b200.h.                 */
int vdo_run(const vd_batch_in *in, vd_batch_out *out) {
    int64_t n_var = in->var_off[4 * (int64_t)in->n_sc];
    memset(out->assigned, 0, 2 * n_var);
    memset(out->sync_group, 0, sizeof(int32_t) * 2 * n_var);
    memset(out->ref_ed, 0, sizeof(int32_t) * 2 * n_var);
    memset(out->query_ed, 0, sizeof(int32_t) * 2 * n_var);
    memset(out->callq, 0, sizeof(float) * 2 * n_var);
    for (int sc = 0; sc < in->n_sc; sc++) {
        hap_t h[4];
        memset(h, 0, sizeof(h));
        int bad = 0;
        for (int k = 0; k < 4; k++) bad |= expand_hap(in, sc, k, &h[k]);
        if (bad) { for (int k = 0; k < 4; k++) hap_free(&h[k]); return VD_E_BADINPUT; }
        const uint8_t *rseq = (in->rplane_seq ? in->rplane_seq : in->ref_seq) + in->ref_off[sc];
        for (int ai = 0; ai < 4; ai++)
            align_one(in, out, n_var, sc, ai, &h[ai >> 1], &h[2 + (ai & 1)], rseq);
        for (int k = 0; k < 4; k++) hap_free(&h[k]);
    }
    return VD_OK;
}
