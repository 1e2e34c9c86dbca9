/*
 * daliid_b200 -- C-ABI of the B200-native retrieval-evaluation hot path of DaliID.
 *
 * The reference (pure Python, /root/reference/Person-ReID) has no FFI of its own:
 * its hot path is CPU torch/numpy plus the third-party `torchreid` evaluator.
 * Each entry point below replaces one reference expression / call (cited as
 * file:line under Person-ReID/); INTEGRATION.md shows the ctypes stub a reference
 * maintainer would add at those lines.
 *
 * Conventions
 *   - plain pointers and sizes, no C++/torch types, no exceptions across the ABI;
 *   - every function returns DALI_OK (0) or a negative DALI_ERR_* code and records
 *     a message retrievable with dali_last_error(ctx);
 *   - data pointers may be HOST or DEVICE memory (queried with
 *     cudaPointerGetAttributes); host buffers are staged through the context's
 *     stream (pinned host memory gives full PCIe rate), device buffers are used
 *     in place.  Label arrays (int32) and small results (cmc, mAP) are HOST.
 *   - the caller owns all buffers; the context owns its stream (unless one is
 *     attached), events and workspaces.  One context per host thread.  A call that
 *     writes any HOST output returns after that output is complete; a call whose
 *     outputs are all DEVICE buffers only enqueues work on the context's stream
 *     (stream-ordered, like a kernel launch) and may return before it has run;
 *   - there is NO CPU fallback: without a usable sm_100 device every compute
 *     entry point returns DALI_ERR_CUDA.
 */
#ifndef DALIID_B200_H_
#define DALIID_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library is built with -fvisibility=hidden */
#endif

#define DALI_ABI_VERSION 1

/* ---- status codes ------------------------------------------------------- */
#define DALI_OK 0
#define DALI_ERR_INVALID (-1)        /* bad argument                                      */
#define DALI_ERR_CUDA (-2)           /* CUDA runtime / driver failure, or no sm_100 GPU   */
#define DALI_ERR_NO_VALID_QUERY (-3) /* torchreid: AssertionError('Error: all query
                                        identities do not appear in gallery')             */
#define DALI_ERR_UNSUPPORTED (-4)
#define DALI_ERR_NOMEM (-5)
#define DALI_ERR_PEER_CAPACITY (-6) /* peer block smaller than the number of matches: re-create it */
#define DALI_ERR_PEER_TIMEOUT (-7)  /* a peer exchange gave up waiting for another rank               */
#define DALI_ERR_FUSED_FALLBACK (-8) /* internal: the fused path declined, the caller takes the matrix path */

/* ---- enums --------------------------------------------------------------- */
/* distance metric (a2 / a2') */
#define DALI_METRIC_COSINE 0      /* 1.0 - q.g          validateModels.py:47, evaluate.py:291      */
#define DALI_METRIC_SQEUCLIDEAN 1 /* |q|^2+|g|^2-2q.g   compute_distance_matrix(..,"euclidean"),
                                     commented call validateModels.py:44, evaluate.py:288          */
#define DALI_METRIC_EUCLIDEAN 2   /* torch.cdist(p=2)   commented call validateModels.py:45         */
#define DALI_METRIC_DOT 3         /* q.g similarity     validateModels.py:179, getFeatures.py:285  */

/* arithmetic of the Q x G x D contraction */
#define DALI_PREC_FP32 0   /* SIMT FFMA, one fmaf chain per element in k order (exact class) */
#define DALI_PREC_TF32X3 1 /* tcgen05 kind::tf32, hi/lo split, 3 MMAs (fp32 class)           */
#define DALI_PREC_TF32 2   /* tcgen05 kind::tf32, single pass (fast; <= 0.01 pp mAP)         */
#define DALI_PREC_TF32C 3  /* tcgen05: TF32 hi*hi + two bf16 correction MMAs (fp32 class,
                              error <= 2^-18 per product; 2 instead of 3 tensor passes)      */
#define DALI_PREC_F16X3 4  /* tcgen05 kind::f16: rows scaled by 2^12 and split into fp16 hi + fp16
                              residual (22 mantissa bits), hi*hi + hi*lo + lo*hi at the 16-bit
                              rate (1.5 TF32 passes, half the operand bytes; fp32 class, error
                              <= 2^-20 per product).  Needs unit rows: normalize != 0.        */
#define DALI_PREC_F16 5    /* tcgen05 kind::f16, single pass on the fp16 hi plane of F16X3: the
                              same 11-bit mantissa as TF32 at twice its rate and half its operand
                              bytes (fast; <= 0.01 pp mAP).  Needs unit rows: normalize != 0.   */

/* accumulation semantics of the CMC/AP reduction */
#define DALI_ACCUM_CY_F32 0 /* torchreid Cython path: C float, sequential in rank order */
#define DALI_ACCUM_PY_F64 1 /* torchreid Python path: float64 terms, numpy pairwise sum  */

typedef struct dali_ctx dali_ctx;
typedef struct dali_rank_plan dali_rank_plan;

/* ---- context --------------------------------------------------------------- */
int dali_abi_version(void);
/* device: CUDA ordinal.  Fails with DALI_ERR_CUDA when it is not compute capability 10.x. */
int dali_ctx_create(dali_ctx **out, int device);
void dali_ctx_destroy(dali_ctx *ctx);
/* Attach a caller-owned cudaStream_t (e.g. torch's current stream); NULL restores the
 * context's own stream.  All kernels and copies of this context are issued on it. */
int dali_ctx_set_stream(dali_ctx *ctx, void *cuda_stream);
void *dali_ctx_get_stream(dali_ctx *ctx);
const char *dali_last_error(dali_ctx *ctx);
const char *dali_strerror(int code);
/* Number of DMA streams host operands are currently copied with (1 or 2).  For host galleries of
 * 24 MB and more the library times its copy/compute pipeline with one and with two streams during
 * the first calls of a context and keeps the faster setting; DALI_H2D_STREAMS=1..4 fixes it. */
int dali_ctx_h2d_streams(const dali_ctx *ctx);

/* Per-kernel device timing.  When enabled, every kernel launch of this context is
 * bracketed by cudaEvents on the context's stream; dali_ctx_timing_read returns, for
 * kernel slot `which` (DALI_K_*), the number of launches and their summed duration in
 * milliseconds since the last dali_ctx_timing_reset.  Reading synchronises the stream. */
#define DALI_K_NORMALIZE 0
#define DALI_K_DISTMAT 1
#define DALI_K_RANK_COUNT 2
#define DALI_K_RANK_FINALIZE 3
#define DALI_K_TOPK 4
#define DALI_K_FUSE 5
#define DALI_K_RANK_GATHER 6
#define DALI_K_RERANK 7
#define DALI_K_MRFUSE 8
#define DALI_K_PEER_EXCHANGE 9 /* includes the wait for the slowest rank */
#define DALI_K_H2D 10          /* host -> device copies of operands (copy streams)  */
#define DALI_K_COUNT_ 11
int dali_ctx_timing_enable(dali_ctx *ctx, int on);
int dali_ctx_timing_reset(dali_ctx *ctx);
int dali_ctx_timing_read(dali_ctx *ctx, int which, int *launches, float *total_ms);
/* Number of this library's kernel launches issued by the context since creation. */
int64_t dali_ctx_launch_count(dali_ctx *ctx);
/* Number of fused calls (dali_topk_features_f32) that overflowed their candidate lists and were
 * redone through the materialised distance matrix (same result, slower). */
int64_t dali_ctx_fallback_count(dali_ctx *ctx);
/* Rank-plan cache.  The plan (gallery index by identity + per-query match lists) depends on the
 * four label arrays only; evaluation code calls the path again and again with the same query and
 * gallery sets (mainKIT.py:154-163 twice per epoch, evaluateCleanATModels.py 7 times per run).
 * The context keeps the last plan and reuses it when the next call's labels are byte-for-byte
 * equal (compared in full on every call).  Enabled by default; DALI_PLAN_CACHE=0 in the
 * environment or dali_ctx_plan_cache_enable(ctx, 0) turn it off. */
int dali_ctx_plan_cache_enable(dali_ctx *ctx, int on);
int64_t dali_ctx_plan_cache_hits(dali_ctx *ctx);
/* Fused distance + positive-rank counting (opt-in).  When enabled, dali_eval_features_f32 never
 * writes the Q x G matrix if the features are device resident (or a small host array), the
 * arithmetic is a tensor-core mode, no matrix is requested and no query has more than 64
 * same-identity gallery items (128 from D = 1536 on): the operands are prepared identity-sorted, the
 * few tiles holding the matches yield the positives' distances (kBand), and every tile is counted
 * against each query's sorted thresholds in the contraction's epilogue (kCount: a binary search per
 * column in shared memory).  Results are bit-identical to the matrix path of the same arithmetic.
 * DISABLED by default: at the Market shapes the extra band wave and the epilogue cost more than the
 * 0.1 ms the matrix read-back takes (DESIGN.md 4.9 has the measurements); DALI_FUSED_COUNT=1 / 0
 * overrides this switch.  dali_ctx_fused_count_calls: evaluations that took the fused path; a call
 * that found a non-finite positive distance is redone through the matrix and counted by
 * dali_ctx_fallback_count instead. */
int dali_ctx_fused_count_enable(dali_ctx *ctx, int on);
int64_t dali_ctx_fused_count_calls(dali_ctx *ctx);

/* ---- a1: row L2 normalisation -------------------------------------------- */
/* out[i,:] = x[i,:] / ||x[i,:]||  (no eps: a zero row yields NaN, as the reference does)
 * replaces  x/torch.norm(x, dim=1, keepdim=True)   validateModels.py:41-42,
 *           evaluate.py:285-286, evaluate_ensembled_models.py:278-279,
 *           evaluateCleanATModels.py:106-107,115-119.
 * norms_opt (may be NULL): [n] row norms (the "magnitudes" of
 * evaluateCleanATModels.py:252).  x, out, norms_opt: host or device, fp32, row-major. */
int dali_normalize_f32(dali_ctx *ctx, const float *x, int64_t n, int64_t d, int64_t ldx,
                       float *out, int64_t ldo, float *norms_opt);

/* ---- a2 / a2': query x gallery distance matrix ----------------------------- */
/* out[i,j] (fp32, row-major, leading dimension ld >= G) for q [Q,D], g [G,D] row-major
 * contiguous.  normalize != 0 applies a1 to both operands first (the reference always
 * does for cosine).  replaces  1.0 - torch.mm(q, g.T)  validateModels.py:47,
 * evaluate.py:260-267,291, evaluate_ensembled_models.py:281,300,
 * evaluateCleanATModels.py:109,121,124. */
int dali_distmat_f32(dali_ctx *ctx, const float *q, int64_t Q, const float *g, int64_t G,
                     int64_t D, int metric, int precision, int normalize, float *out,
                     int64_t ld);

/* ---- a2 + a4: one model's distance matrix folded into the running mean -------------------- */
/* The n_models matrices of an ensemble, one call per model (step = 0 .. n_models-1, in the
 * reference's order of addition), summed in the contraction's epilogue instead of by a separate
 * pass over n_models + 1 matrices:
 *   step 0: acc = d;   0 < step < n_models-1: acc = acc + d;   last step: acc = (acc + d) / n_models
 * with the roundings of dali_fuse_f32's mean, so acc ends bit-identical to fusing the separately
 * materialised matrices.  out_opt (may be NULL) also receives this model's own matrix (the
 * reference ranks every model before the ensemble).  acc and out_opt: DEVICE buffers, rows 16-byte
 * aligned (ld multiple of 4); tensor-core precisions only -- otherwise DALI_ERR_UNSUPPORTED, and
 * the caller takes dali_distmat_f32 + dali_fuse_f32.
 * replaces  (distmat01 + distmat02)/2  evaluate_ensembled_models.py:313,
 *           (d_backbone + d_head01 + d_head02)/3  evaluate.py:260-278. */
int dali_distmat_fuse_mean_f32(dali_ctx *ctx, const float *q, int64_t Q, const float *g, int64_t G,
                               int64_t D, int metric, int precision, int normalize, float *out_opt,
                               int64_t ld_out, float *acc, int64_t ld_acc, int step, int n_models);

/* Test hook: the mean's division x / n_models is carried out as a multiplication by RN(1/n) refined
 * by two FMAs (correctly rounded, i.e. what numpy / torch compute); this compares it with the IEEE
 * division over ALL 2^32 fp32 operands on the device and reports the number of differing results. */
int dali_selftest_mean_division(dali_ctx *ctx, int n, uint64_t *mismatches);

/* ---- a4: multi-model distance fusion -------------------------------------- */
/* wq == NULL: out = ((d[0]+d[1])+...+d[n-1]) / n   fp32, left to right, true division
 *   replaces evaluate.py:278, evaluate_ensembled_models.py:313, evaluateCleanATModels.py:127.
 * wq,wg != NULL: w_m[i,j] = max(wq[m][i], wg[m][j]);
 *   out = (w_0*d_0 + w_1*d_1 + ...) / (w_0 + w_1 + ...)   each product and sum rounded
 *   replaces evaluateCleanATModels.py:154-157,193-196,230-233.
 * d[m]: [Q,G] fp32 with leading dimension ld (all equal), host or device (all alike);
 * wq[m]: [Q], wg[m]: [G], host or device.  out may alias d[0]. */
int dali_fuse_f32(dali_ctx *ctx, const float *const *d, int n, const float *const *wq,
                  const float *const *wg, float *out, int64_t Q, int64_t G, int64_t ld);

/* ---- a5: rank + junk mask + CMC/mAP from a distance matrix ------------------ */
/* replaces torchreid.metrics.evaluate_rank(distmat, q_pids, g_pids, q_camids, g_camids,
 *          use_metric_cuhk03=False)   validateModels.py:68-69, evaluate.py:312-313,
 *          evaluate_ensembled_models.py:324-325, evaluateCleanATModels.py:266-267.
 * dist: [Q,G] fp32, leading dimension ld, host or device.  Labels: HOST int32 (any
 * values; only equality matters).  Canonical order: distance ascending, gallery index
 * ascending, NaN last, -0 == +0 (== numpy stable argsort).  max_rank is clamped to G.
 * Outputs (HOST): cmc[max_rank] float32, *mAP (double holding the value the chosen
 * accumulation produces), ap_opt[Q] per-query AP (NaN for invalid queries),
 * first_rank_opt[Q] 1-based kept rank of the first true match (-1 invalid),
 * num_valid_opt.  Returns DALI_ERR_NO_VALID_QUERY when no query has a valid match. */
int dali_eval_rank_f32(dali_ctx *ctx, const float *dist, int64_t Q, int64_t G, int64_t ld,
                       const int32_t *q_pid, const int32_t *g_pid, const int32_t *q_cam,
                       const int32_t *g_cam, int max_rank, int accum_mode, float *cmc,
                       double *mAP, double *ap_opt, int32_t *first_rank_opt,
                       int64_t *num_valid_opt);

/* ---- a1+a2+a5 fused: features in, CMC/mAP out -------------------------------- */
/* replaces validateModels.validate's arithmetic (validateModels.py:41-47,61-69) and the
 * single-model branch of evaluate.py (285-302).  q,g: host or device fp32 row-major.
 * distmat_opt (may be NULL): receives the [Q,G] matrix (ld_opt >= G), host or device. */
int dali_eval_features_f32(dali_ctx *ctx, const float *q, int64_t Q, const float *g, int64_t G,
                           int64_t D, const int32_t *q_pid, const int32_t *g_pid,
                           const int32_t *q_cam, const int32_t *g_cam, int metric,
                           int precision, int normalize, int max_rank, int accum_mode,
                           float *cmc, double *mAP, double *ap_opt, int32_t *first_rank_opt,
                           int64_t *num_valid_opt, float *distmat_opt, int64_t ld_opt);

/* ---- a7 / a8: top-k ---------------------------------------------------------- */
/* Per row the k best columns of dist [Q,G] (ld), smallest first (largest != 0: largest
 * first), ties by ascending column id.  replaces torch.argsort(distmat, dim=1)[:, :20]
 * validateModels.py:93 and torch.topk(S, k=5, largest=True) validateModels.py:180.
 * col_ids_opt (may be NULL): int32 [Q,ld] ids reported and used for the tie-break instead
 * of the column number (merging per-shard candidate lists).  d_out [Q,k] fp32 and
 * i_out [Q,k] int32: host or device.  k <= 128.  Rows shorter than k are padded with
 * (+inf | -inf, -1). */
int dali_topk_f32(dali_ctx *ctx, const float *dist, int64_t Q, int64_t G, int64_t ld, int k,
                  int largest, const int32_t *col_ids_opt, float *d_out, int32_t *i_out);

/* Fused a1+a2+a7 for 1:N identification without materialising Q x G (BASELINE config 5):
 * ids reported are g_base + column. */
int dali_topk_features_f32(dali_ctx *ctx, const float *q, int64_t Q, const float *g, int64_t G,
                           int64_t D, int metric, int precision, int normalize, int k,
                           int largest, int32_t g_base, float *d_out, int32_t *i_out);

/* Merge of per-shard top-k lists (gallery-sharded identification): vals / ids are DEVICE arrays
 * [parts][Q][k] (the layout an all-gather of the per-rank [Q,k] results produces), ids global
 * gallery ids (-1 = padding).  d_out / i_out: DEVICE [Q,k], the k best over all parts in the same
 * order as dali_topk_f32 (value, then id).  k <= 32.  Stream-ordered, no host synchronisation. */
int dali_topk_merge_f32(dali_ctx *ctx, const float *vals, const int32_t *ids, int parts, int64_t Q,
                        int k, int largest, float *d_out, int32_t *i_out);

/* ---- next row N1: k-reciprocal re-ranking ------------------------------------- */
/* out[Q,G] = torchreid.utils.re_ranking(qg, qq, gg, k1, k2, lambda_value): the hook the
 * reference keeps commented out at validateModels.py:49-53, evaluate.py:294-298,
 * evaluate_ensembled_models.py:284-288,303-307 (flag: validateModels.py:28-31).
 * qg [Q,G], qq [Q,Q], gg [G,G] fp32 with leading dimensions ld_*; all host or all device; out
 * [Q,G] (ld_out), host or device.  1 <= k1 <= 28, 1 <= k2 <= min(8, k1+1).  Neighbour sets are
 * exact (stable tie order); values agree with the numpy form to fp32 rounding. */
int dali_rerank_f32(dali_ctx *ctx, const float *qg, int64_t ld_qg, const float *qq, int64_t ld_qq,
                    const float *gg, int64_t ld_gg, int64_t Q, int64_t G, int k1, int k2,
                    double lambda_value, float *out, int64_t ld_out);

/* ---- next row N4: whole ranked lists ------------------------------------------- */
/* idx_out[q, r] = column holding the r-th smallest (largest when `descending`) value of row q:
 * torch.argsort(dist, dim=1, descending=..., stable=True).  replaces the full orderings of
 * getFeatures.py:303,347 (get_subset / get_subset_one_encoder: one row of similarities to the
 * whole training set, best first) and the `indices` matrix inside torchreid's evaluation.
 * Ties by ascending column, NaN after +inf (first when descending), -0 == +0.
 * dist fp32 [Q,ld] and idx_out int32 [Q,G] contiguous: each host or device. */
int dali_argsort_f32(dali_ctx *ctx, const float *dist, int64_t Q, int64_t G, int64_t ld,
                     int descending, int32_t *idx_out);

/* ---- next row N3: meta-recognition score fusion --------------------------------- */
/* fused[Q,G] (fp64) = Meta_Recognition.mrfuse(scores...) of evaluate.py:610-627 (call site, kept
 * commented by the reference: evaluate.py:277; same class in evaluate_ensembled_models.py:593-637):
 * per model and gallery column a 2-parameter Weibull is fitted (libmr.FitHigh / _fit,
 * evaluate.py:429-432, 531-580) to the column's Q-topk-1 largest scores after the top-`topk`
 * scores of every query row (use_columns = 0, what mrfuse uses) or of every gallery column
 * (use_columns = 1, metarec's default) were reduced by killscale * themselves; the weight of a
 * score is that Weibull's CDF (libmr.wscore, evaluate.py:434-473) and
 * fused = sum_m(w_m * s_m) / sum_m(w_m).
 * scores: n (1..3) pointers to fp32 [Q,ld] SIMILARITY matrices (1 - distance), all host or all
 * device.  fused: fp64 [Q,ld_out], host or device.  Optional outputs (NULL to skip; host or
 * device): fit_opt fp64 [n][G][2] = (shape, scale) exactly as libmr.wbFits holds them (0,0 for a
 * column whose Newton iteration did not converge in 100 steps, NaN,NaN when it turned NaN),
 * small_opt fp32 [n][G] = libmr.smallScoreTensor, weights_opt fp64 [n][Q][ld_out] = metarec's
 * return value.  Needs Q >= topk+2, topk <= 126, and G >= topk when use_columns = 0.
 * Values agree with the reference run on the same inputs to ~1e-6 relative (the reference's fp32
 * log / mean intermediates are not reproducible beyond that across libm implementations). */
int dali_mrfuse_f32(dali_ctx *ctx, const float *const *scores, int n, int64_t Q, int64_t G,
                    int64_t ld, int topk, int use_columns, float killscale, double *fused,
                    int64_t ld_out, double *fit_opt, float *small_opt, double *weights_opt);

/* ROC over all Q x G pairs without a sort: the verification branch of
 * evaluateCleanATModels.py:276-292 (label = same identity, score = 1.0 - distmat / 2.0 in fp32,
 * sklearn.metrics.roc_curve).  One streaming pass histograms the scores of the two classes over
 * nbins uniform bins of [lo, hi] (bin = floor((score - lo) * nbins / (hi - lo)), clamped; NaN -> bin
 * 0); the suffix sums of pos_hist / neg_hist are exact points (tps, fps) of the curve at the nbins
 * thresholds "smallest score of bin b".  dist: fp32 [Q, ld], host or device; labels: HOST int32;
 * pos_hist, neg_hist: HOST uint64 [nbins]. */
int dali_roc_hist_f32(dali_ctx *ctx, const float *dist, int64_t Q, int64_t G, int64_t ld,
                      const int32_t *q_pid, const int32_t *g_pid, int nbins, float lo, float hi,
                      uint64_t *pos_hist, uint64_t *neg_hist);

/* ---- (e) gallery-sharded building blocks ------------------------------------ */
/* One process per GPU holds all Q queries and a contiguous gallery slab
 * [g0, g0+Gs).  Labels of the WHOLE gallery are replicated (small).  The host side
 * (daliid_b200/sharded.py, torch.distributed) runs:
 *   plan -> gather_keys -> allreduce(sum) -> count -> allreduce(sum) -> finalize.
 * A "match" is a (query, gallery) pair with equal pid (valid positive or junk); the
 * plan lays all matches out in one array of length M (query-major, gallery ascending). */
int dali_rank_plan_create(dali_ctx *ctx, const int32_t *q_pid, const int32_t *g_pid,
                          const int32_t *q_cam, const int32_t *g_cam, int64_t Q, int64_t G,
                          dali_rank_plan **out);
void dali_rank_plan_destroy(dali_rank_plan *plan);
int64_t dali_rank_plan_num_matches(const dali_rank_plan *plan);
/* keys_out[M] (DEVICE uint32): order-preserving key of dist[q, g-g0] for matches whose
 * gallery item lies in the slab, 0 elsewhere (so a sum over ranks assembles all keys). */
int dali_rank_gather_keys(dali_ctx *ctx, const dali_rank_plan *plan, const float *dist_slab,
                          int64_t ld, int64_t g0, int64_t Gs, uint32_t *keys_out);
/* counts_out[M] (DEVICE int32): number of slab columns j with
 * (key(dist[q,j]), g0+j) <lex (keys[m], gallery id of match m). */
int dali_rank_count(dali_ctx *ctx, const dali_rank_plan *plan, const float *dist_slab,
                    int64_t ld, int64_t g0, int64_t Gs, const uint32_t *keys,
                    int32_t *counts_out);
/* keys, counts: DEVICE, already summed over ranks.  Outputs as dali_eval_rank_f32. */
int dali_rank_finalize(dali_ctx *ctx, const dali_rank_plan *plan, const uint32_t *keys,
                       const int32_t *counts, int max_rank, int accum_mode, float *cmc,
                       double *mAP, double *ap_opt, int32_t *first_rank_opt,
                       int64_t *num_valid_opt);

/* ---- (e) exchange over NVLink peer memory --------------------------------------- */
/* The two exchange steps above are sums of small int32 arrays (one word per match).  Instead of
 * an NCCL all-reduce each (host launch + tens of microseconds on the device), every rank keeps its
 * contribution in a block that the other ranks of the node map through CUDA IPC, and one kernel
 * per exchange signals, waits and reduces with peer loads (daliid_b200/csrc/peer.cu).
 *   dali_peer_create      allocate this rank's block: two buffers of `capacity` int32
 *   dali_peer_ipc_handle  64-byte cudaIpcMemHandle of the block (exchange it with any transport,
 *                         e.g. torch.distributed.all_gather_object)
 *   dali_peer_connect     map the blocks of all ranks; handles = world x 64 bytes, rank order
 *   dali_peer_buffer      DEVICE pointer of buffer `which` (0 / 1): pass it as keys_out of
 *                         dali_rank_gather_keys or counts_out of dali_rank_count
 *   dali_peer_allreduce_i32  out[i] = sum over ranks of buffer `which`[i], i < n.  Collective:
 *                         every rank must call it the same number of times in the same order.
 * One process per GPU, all GPUs on one node (NVLink / NVSwitch), world <= 8. */
typedef struct dali_peer dali_peer;
int dali_peer_create(dali_ctx *ctx, int rank, int world, int64_t capacity, dali_peer **out);
int dali_peer_ipc_handle(dali_peer *peer, void *handle64);
int dali_peer_connect(dali_peer *peer, const void *handles);
void dali_peer_destroy(dali_peer *peer);
int64_t dali_peer_capacity(const dali_peer *peer);
void *dali_peer_buffer(dali_peer *peer, int which);
int dali_peer_allreduce_i32(dali_ctx *ctx, dali_peer *peer, int which, int32_t *out, int64_t n);
/* DALI_OK, or DALI_ERR_PEER_TIMEOUT when an exchange since dali_peer_create gave up waiting for
 * another rank (limit: DALI_PEER_TIMEOUT_MS, default 20000).  The kernel does not trap: the
 * context stays usable, but results computed from that exchange are meaningless.  Synchronises. */
int dali_peer_status(dali_peer *peer);
/* The whole sharded evaluation of one rank in one call: slab contraction, plan, gather, exchange,
 * count, exchange, finalize (the sequence daliid_b200/sharded.py otherwise drives call by call).
 * q [Q,D] and g_slab [Gs,D]: host or device; g0 = first gallery index of the slab; labels of the
 * WHOLE gallery (G_total).  Outputs as dali_eval_rank_f32, identical on every rank.  Returns
 * DALI_ERR_PEER_CAPACITY with *matches_out = required capacity when the peer block is too small
 * (every rank gets the same answer: re-create the block collectively and call again).
 * Collective: all ranks call it together. */
int dali_eval_features_sharded_f32(dali_ctx *ctx, dali_peer *peer, const float *q, int64_t Q,
                                   const float *g_slab, int64_t Gs, int64_t D, int64_t g0,
                                   int64_t G_total, const int32_t *q_pid, const int32_t *g_pid_all,
                                   const int32_t *q_cam, const int32_t *g_cam_all, int metric,
                                   int precision, int normalize, int max_rank, int accum_mode,
                                   float *cmc, double *mAP, double *ap_opt, int32_t *first_rank_opt,
                                   int64_t *num_valid_opt, int64_t *matches_out);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* DALIID_B200_H_ */
