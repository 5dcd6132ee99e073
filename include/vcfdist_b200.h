/* vcfdist_b200 — C-ABI boundary of the B200-native precision/recall hot path.
 *
 * Replaces, for the reference TimD1/vcfdist v2.6.4, everything underneath
 *
 *     void precision_recall_threads_wrapper(
 *             std::shared_ptr<superclusterData> clusterdata_ptr,
 *             std::vector<std::vector<std::vector<int>>> sc_groups);
 *                                   (decl src/dist.h:245-247, def src/dist.cpp:1656-1727,
 *                                    called once from src/main.cpp:219)
 *
 * i.e. per supercluster: generate_ptrs_strs (src/dist.cpp:145-242), calc_prec_recall_aln
 * (:251-443), calc_prec_recall_path (:486-834), get_prec_recall_path_sync (:842-999),
 * the integer part of calc_prec_recall (:1005-1401) and wf_ed (:1406-1506).
 *
 * The reference has no FFI for this path; the seam is that one free function.  The
 * host-side drop-in (vcfdist_b200/host/pr_dropin.cpp) keeps its exact signature,
 * packs `superclusterData` into a `vd_batch_in`, calls vd_run() and scatters the
 * integers back, evaluating the two float threshold tests (src/dist.cpp:465-467,
 * :1293-1296, :1327-1328) on the host with the reference's own expression shapes.
 *
 * Conventions: plain pointers and sizes only; all buffers are caller-owned; every
 * entry point returns 0 on success or a negative VD_E* code and never exits or
 * throws.  There is no CPU fallback: without a CUDA device vd_create() fails.
 */
#ifndef VCFDIST_B200_H
#define VCFDIST_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VD_ABI_VERSION 2

/* variant types, identical to src/defs.h:31-35 */
#define VD_TYPE_SUB 1
#define VD_TYPE_INS 2
#define VD_TYPE_DEL 3

/* haplotype order inside a supercluster, and the four alignments (src/defs.h:101-104):
 * alignment i aligns query hap (i>>1) against truth hap (i&1)                         */
#define VD_HAP_Q1 0
#define VD_HAP_Q2 1
#define VD_HAP_T1 2
#define VD_HAP_T2 3

/* error codes */
#define VD_OK            0
#define VD_E_NODEVICE   -1   /* no CUDA device / wrong architecture                    */
#define VD_E_CUDA       -2   /* a CUDA runtime call failed (see vd_last_error)         */
#define VD_E_BADINPUT   -3   /* malformed batch (unsorted/overlapping variants, ...)   */
#define VD_E_TOOLARGE   -4   /* one supercluster alone exceeds the HBM scratch budget  */
#define VD_E_NOMEM      -5
#define VD_E_ALIGN      -6   /* an alignment hit a fatal reference condition
                                (src/dist.cpp:314, :440, :606, :644, :937)             */
#define VD_E_RANGE      -7   /* vd_run_packed: a value does not fit the 16-bit records
                                (call vd_run instead)                                  */

/* per-alignment status bits, vd_batch_out.status[4*sc + i] */
#define VD_ST_TIE            0x0001u  /* a swap edge on a reachable optimal cell had more than one
                                         equal-score source row: the reference's choice depends on
                                         std::unordered_set iteration order (src/dist.cpp:347,376);
                                         we pick the larger source row                           */
#define VD_ST_WARN_REFED_NOTRUTH 0x0002u  /* WARN src/dist.cpp:1203-1206 */
#define VD_ST_WARN_QED_NOQUERY   0x0004u  /* WARN src/dist.cpp:1207-1210 */
#define VD_ST_WARN_QED_GT_REFED  0x0008u  /* WARN src/dist.cpp:1211-1214 */
#define VD_ST_WARN_ZERO_REFED    0x0010u  /* WARN src/dist.cpp:1219-1223 (ref_ed forced to 1) */
#define VD_ST_ERR_NO_POINTER     0x0100u  /* ERROR src/dist.cpp:937  */
#define VD_ST_ERR_NO_SWAP_PRED   0x0200u  /* ERROR src/dist.cpp:606, :644 */
#define VD_ST_ERR_UNFINISHED     0x0400u  /* ERROR src/dist.cpp:314, :440 */
#define VD_ST_ERR_BADINPUT       0x0800u  /* malformed supercluster (see VD_E_BADINPUT) */
#define VD_ST_ERR_MASK           0xff00u

/* vd_batch_out.assigned[] values */
#define VD_ASSIGN_NONE    0   /* variant never visited: stays ERRTYPE_UN                  */
#define VD_ASSIGN_REF_FP  1   /* passed while the path ran on the REF plane: FP, credit 0
                                 (src/dist.cpp:1157-1168)                                 */
#define VD_ASSIGN_SYNC    2   /* credited at a sync section (src/dist.cpp:1291-1353); the
                                 host derives credit, TP/FP/FN from ref_ed and query_ed    */

/* One batch of superclusters, in the compact form the reference's per-supercluster
 * driver reads them (src/dist.cpp:1786-1822): the reference window plus the four
 * haplotypes' variant lists.  Haplotype strings and the query<->ref / truth->ref
 * pointer arrays of generate_ptrs_strs are expanded on the device.
 *
 * Variants of (supercluster s, hap h) are var_off[4*s+h] .. var_off[4*s+h+1]-1, h in
 * VD_HAP_* order, sorted by position, non-overlapping, inside the window.             */
typedef struct vd_batch_in {
    int32_t        n_sc;
    const int64_t *ref_off;     /* [n_sc+1] byte offsets into ref_seq; window s is
                                   fasta[ctg][begs[s] .. ends[s]] inclusive (src/dist.cpp:163,232) */
    const uint8_t *ref_seq;     /* FASTA bytes (upper-cased by src/fasta.h:19-20)                   */
    const uint8_t *rplane_seq;  /* optional, may be NULL: the REF-plane string (= ref_q1 of
                                   src/dist.cpp:1784-1792) when query-hap-1 REF alleles differ from
                                   the FASTA; same offsets as ref_seq                                */
    const int64_t *var_off;     /* [4*n_sc+1]                                                        */
    const int32_t *var_pos;     /* [n_var] poss[v] - begs[s]                (src/dist.cpp:166,1076)  */
    const int32_t *var_rlen;    /* [n_var] refs[v].size(): INS 0, SUB 1, DEL deleted length          */
    const uint8_t *var_type;    /* [n_var] VD_TYPE_*                                                 */
    const int64_t *alt_off;     /* [n_var+1] byte offsets into alt_seq                               */
    const uint8_t *alt_seq;     /* alts[v] bytes (INS: inserted bases, SUB: one base, DEL: none)     */
    const float   *var_qual;    /* [n_var] var_quals[v]; only query haps are read (:1163,:1287)      */
    float          max_qual;    /* float(g.max_qual)                        (src/dist.cpp:1284)      */
} vd_batch_in;

/* Results.  Per-variant arrays are indexed [v] with v the batch-global variant index;
 * each variant is written by exactly two alignments, one per phasing slot
 * (src/dist.cpp:1037-1048): slot 0 = original phasing (Q1T1,Q2T2), slot 1 = swapped
 * (Q1T2,Q2T1).  Element (slot, v) lives at [slot * n_var + v].                           */
typedef struct vd_batch_out {
    int32_t  *aln_score;       /* [4*n_sc] s[i]                         (src/dist.cpp:426)        */
    uint8_t  *aln_end_plane;   /* [4*n_sc] 0 QUERY plane, 1 REF plane   (src/dist.cpp:436-440)    */
    uint8_t  *aln_beg_plane;   /* [4*n_sc] plane of the path origin     (src/dist.cpp:811-814)    */
    uint32_t *status;          /* [4*n_sc] VD_ST_* bits                                            */
    uint8_t  *assigned;        /* [2*n_var] VD_ASSIGN_*                                            */
    int32_t  *sync_group;      /* [2*n_var]                                                        */
    int32_t  *ref_ed;          /* [2*n_var]                                                        */
    int32_t  *query_ed;        /* [2*n_var]                                                        */
    float    *callq;           /* [2*n_var] min var_qual over the section's query variants,
                                  starting from max_qual (src/dist.cpp:1284-1288); for
                                  VD_ASSIGN_REF_FP the variant's own quality (:1163)               */
} vd_batch_out;

/* The same results in 16-bit records: 20 bytes per supercluster and 20 per variant instead of 40 and 34, for
 * callers that keep them as they are (vd_finalize_packed reads them directly) - the copy out over PCIe is what
 * bounds an end-to-end step on a WGS batch.  Values that do not fit (a score, ref_ed or query_ed of 65535 or
 * more, 16384 or more sync groups in one alignment) make vd_run_packed return VD_E_RANGE.                      */
typedef struct vd_packed_out {
    uint16_t *aln_score;       /* [4*n_sc] 0xffff: no score (malformed supercluster)                            */
    uint8_t  *aln_planes;      /* [4*n_sc] bit 0 end plane, bit 1 origin plane                                  */
    uint16_t *status;          /* [4*n_sc] VD_ST_* bits                                                         */
    uint16_t *sync_group;      /* [2*n_var] assigned << 14 | sync_group                                         */
    uint16_t *ref_ed;          /* [2*n_var]                                                                     */
    uint16_t *query_ed;        /* [2*n_var]                                                                     */
    float    *callq;           /* [2*n_var]                                                                     */
} vd_packed_out;

/* The batch in compact form, for the copy in over PCIe (which bounds an end-to-end step once the results are 16-bit
 * records): lengths instead of 64-bit offsets, 16-bit positions and lengths - 47 instead of 111 bytes per supercluster
 * on a WGS batch.  The offsets are rebuilt on the GPU (prefix sums).  vd_compact_pack fills it from a vd_batch_in and
 * returns VD_E_RANGE when a value does not fit (a window or REF/ALT allele of 65536 bases or more, more than 255
 * variants on one haplotype of a supercluster): use vd_run / vd_run_packed for such a batch.                       */
#define VD_COMPACT_BLOCK 65536   /* superclusters per entry of the blk_* index (where the host pipeline may cut)     */
typedef struct vd_compact_in {
    int32_t n_sc;
    float   max_qual;
    int64_t n_var, ref_bytes, alt_bytes;
    const uint16_t *ref_len;     /* [n_sc] window length (ends - begs + 1)                                          */
    const uint8_t  *ref_seq;     /* [ref_bytes]                                                                      */
    const uint8_t  *rplane_seq;  /* [ref_bytes] or NULL, as in vd_batch_in                                           */
    const uint8_t  *hap_nvar;    /* [4*n_sc] variants per haplotype q1,q2,t1,t2                                      */
    const uint16_t *var_pos;     /* [n_var]                                                                          */
    const uint16_t *var_rlen;    /* [n_var]                                                                          */
    const uint16_t *alt_len;     /* [n_var]                                                                          */
    const uint8_t  *var_type;    /* [n_var]                                                                          */
    const uint8_t  *alt_seq;     /* [alt_bytes]                                                                      */
    const float    *var_qual;    /* [n_var]                                                                          */
    const int64_t  *blk_var;     /* [n_blk+1], n_blk = ceil(n_sc / VD_COMPACT_BLOCK): first variant of block b ...   */
    const int64_t  *blk_ref;     /* ... its first window byte ...                                                    */
    const int64_t  *blk_alt;     /* ... its first ALT byte; entry n_blk = the totals                                 */
} vd_compact_in;

/* counters of the last vd_run*/
typedef struct vd_stats {
    int64_t n_sc, n_var;
    int64_t cells;            /* sum over alignments of (Lq+Lr)*Lt                  */
    int64_t io_bytes;         /* algorithmic input+output bytes                        */
    int64_t spill_bytes;      /* 3 B/cell for alignments whose flag matrices live in HBM           */
    int64_t n_short, n_long;  /* alignments handled by the short / long kernels                    */
    int64_t n_launches;       /* kernels launched                                                  */
    int64_t h2d_bytes, d2h_bytes;
    float   ms_total;         /* CUDA-event time of the whole call on the handle's stream          */
    float   ms_short;         /* ... of the short-supercluster kernels (all classes)              */
    float   ms_long_fwd, ms_long_bwd, ms_long_walk;
    float   ms_plan;
    float   ms_long_wall;     /* wall time of the concurrent forward+backward region            */
    float   ms_small[3];      /* ... of the fused shared-memory kernels: thread-per-alignment class 0,
                                 class 1, warp-per-supercluster kernel (all its launches)           */
    int64_t n_small[3];       /* superclusters handled by each of them                             */
    int64_t io_small[3];      /* algorithmic input+output bytes of those superclusters             */
    int64_t n_hom;            /* homozygous superclusters among them (both query haplotypes identical and
                                 both truth haplotypes identical: one alignment computed, records replicated) */
    int64_t n_dense;          /* long alignments the banded warp kernels left to the dense block kernels        */
    float   ms_band;          /* wall time of the banded warp kernels (all rungs, forward + backward + walk)    */
    float   pad_;
    int64_t band_cells;       /* cells the banded forward sweeps visited in the alignments they solved           */
    int64_t band_rows;        /* ... rows (both planes) and truth columns of those alignments                    */
    int64_t band_cols;
} vd_stats;

/* Final per-variant / per-supercluster results in the reference's own terms
 * (src/variant.h:49-60, src/cluster.h:36-42, src/defs.h:66-72, :131-134).              */
typedef struct vd_final {
    uint8_t *errtypes;     /* [2*n_var] ERRTYPE_TP 0 / FP 1 / FN 2 / UN 5                       */
    float   *credit;       /* [2*n_var]                                                        */
    float   *callq;        /* [2*n_var]                                                        */
    int32_t *sync_group;   /* [2*n_var]                                                        */
    int32_t *ref_ed;       /* [2*n_var]                                                        */
    int32_t *query_ed;     /* [2*n_var]                                                        */
    int32_t *sc_phase;     /* [n_sc] PHASE_ORIG 0 / SWAP 1 / NONE 2                            */
    int32_t *orig_dist;    /* [n_sc] s[Q1T1]+s[Q2T2]                                           */
    int32_t *swap_dist;    /* [n_sc] s[Q2T1]+s[Q1T2]                                           */
} vd_final;

/* Host-only float step: store_phase (src/dist.cpp:449-475) and the credit / TP-FP-FN
 * decisions of calc_prec_recall (src/dist.cpp:1293-1352), evaluated with the reference's
 * own expression shapes (float division, comparison against a double threshold) from the
 * integers the device returned.  Needs no GPU.                                            */
int  vd_finalize(const vd_batch_in *in, const vd_batch_out *out,
                 double phase_threshold, double credit_threshold, vd_final *fin);

typedef struct vd_handle vd_handle;

int  vd_abi_version(void);

/* One handle per GPU and per host thread; owns a stream, pinned staging buffers and HBM
 * scratch.  `scratch_bytes` bounds the HBM used for spilled flag matrices (0 = default:
 * 60 % of free memory), the analogue of the reference's --max-ram ladder
 * (src/cluster.cpp:94-117, src/dist.cpp:1700-1721).                                      */
int  vd_create(int device, int64_t scratch_bytes, vd_handle **out);
void vd_destroy(vd_handle *h);

/* Synchronous.  Host buffers in, host buffers out: H2D copy, kernels, D2H copy.          */
int  vd_run(vd_handle *h, const vd_batch_in *in, vd_batch_out *out);
int  vd_run_packed(vd_handle *h, const vd_batch_in *in, vd_packed_out *out);
int  vd_finalize_packed(const vd_batch_in *in, const vd_packed_out *out,
                        double phase_threshold, double credit_threshold, vd_final *fin);

/* Host only: fills *out (whose array pointers the caller has allocated: sizes as commented in vd_compact_in) from *in. */
int  vd_compact_pack(const vd_batch_in *in, vd_compact_in *out);
/* vd_run_packed with the batch in compact form.                                                                    */
int  vd_run_compact(vd_handle *h, const vd_compact_in *in, vd_packed_out *out);

/* Page-locked host memory for the buffers of vd_run / vd_run_packed (plain cudaHostAlloc / cudaFreeHost, so that a
 * caller needs no CUDA headers): copies from and to pageable memory run at a fraction of the PCIe rate.
 * NULL when the allocation fails.                                                                             */
void *vd_host_alloc(int64_t bytes);
void  vd_host_free(void *p);

/* Same computation with every pointer of `in` and `out` already resident in this GPU's
 * HBM (n_var and the byte sizes are passed since the offsets live on the device).
 * Asynchronous work is enqueued on the handle's stream and synchronised before return.   */
int  vd_run_device(vd_handle *h, const vd_batch_in *in_dev, vd_batch_out *out_dev,
                   int64_t n_var, int64_t ref_bytes, int64_t alt_bytes);

/* A slice of a resident batch: superclusters [first_sc, first_sc + in_dev->n_sc) of the batch whose
 * arrays in_dev points at - ref_off and var_off are passed already advanced to the slice (ref_off +
 * first_sc, var_off + 4*first_sc), every other array whole, since the offsets stay batch-absolute.
 * first_var = var_off[4*first_sc] and n_var = the slice's variant count.  out_dev holds the SLICE's
 * own result arrays ([4*n_sc] and [2*n_var], slot stride n_var).  Lets a caller overlap the exchange
 * of one slice's results (all-gather, copy-out) with the kernels of the next.                       */
int  vd_run_device_slice(vd_handle *h, const vd_batch_in *in_dev, vd_batch_out *out_dev,
                         int64_t first_var, int64_t n_var, int64_t ref_bytes, int64_t alt_bytes);

/* 16-bit records from wide ones, both resident in this GPU's HBM: enqueued on the handle's stream behind the kernels
 * of vd_run_device / vd_run_device_slice (the exchange step of a multi-GPU run moves the narrow records).  packed->callq
 * may alias wide->callq or be null (nothing to narrow there).  vd_packed_overflow: nonzero once a value did not fit,
 * valid after the stream has been synchronised; reset by the next vd_run_device*.                                   */
int  vd_pack_device(vd_handle *h, const vd_batch_out *wide_dev, int64_t n_sc, int64_t n_var, const vd_packed_out *packed_dev);
int  vd_packed_overflow(const vd_handle *h);

int  vd_get_stats(const vd_handle *h, vd_stats *out);
const char *vd_last_error(const vd_handle *h);

/* The handle's CUDA stream (a cudaStream_t), so that a caller can order its own work,
 * e.g. torch.cuda.ExternalStream(vd_stream(h)).                                           */
void *vd_stream(const vd_handle *h);

/* ---- cluster-growing stage (the step before the path) --------------------------------
 * Batches of the two affine-gap wavefront searches wf_swg_cluster (src/cluster.cpp:954-1263) is made of:
 *   VD_WF_REACH  wf_swg_max_reach (src/dist.cpp:2150-2333; call sites src/cluster.cpp:1080-1090, :1139-1149): the
 *                furthest truth index reachable with a score <= max_score[i]; reverse[i] = 1 for the leftward
 *                search (the caller passes the reversed strings, as the reference does)
 *   VD_WF_SCORE  the score of wf_swg_align (src/dist.cpp:1510-1652; call site src/cluster.cpp:1040-1046)
 * Problem i: query = q_seq[q_off[i] .. q_off[i+1]), truth = t_seq[t_off[i] .. t_off[i+1]); penalties sub / open /
 * extend as g.query_sub etc.; host buffers in, result[n] out; one warp per problem on the device.            */
#define VD_WF_REACH 0
#define VD_WF_SCORE 1
int  vd_wf_batch(vd_handle *h, int mode, int n, const int64_t *q_off, const uint8_t *q_seq, const int64_t *t_off,
                 const uint8_t *t_seq, const int32_t *main_diag, const int32_t *main_diag_start, const int32_t *max_score,
                 const uint8_t *reverse, int sub, int open, int extend, int32_t *result);

/* ---- `--distance` pass ------------------------------------------------------------------
 * Batches of the affine-gap alignment edits_wrapper (src/dist.cpp:1908-2077) runs per supercluster and haplotype:
 * wf_swg_align with its predecessor flags (src/dist.cpp:1510-1652) + wf_swg_backtrack (:2625-2757).
 * score[i]: the alignment score, -1 where the reference's walk back would ERROR(); cigar: problem i at
 * cigar[q_off[i] + t_off[i] ..], |query| + |truth| entries filled from the back exactly as the reference fills its
 * vector (PTR_INS 1 / PTR_DEL 2 / PTR_MAT 4 / PTR_SUB 8; two entries per match or substitution), zeros in front. */
int  vd_swg_align_batch(vd_handle *h, int n, const int64_t *q_off, const uint8_t *q_seq, const int64_t *t_off,
                        const uint8_t *t_seq, int sub, int open, int extend, int32_t *score, int32_t *cigar);

#ifdef __cplusplus
}
#endif
#endif
