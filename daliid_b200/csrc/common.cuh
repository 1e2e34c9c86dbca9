// Internal definitions shared by the kernels and the C-ABI glue of libdaliid_b200.
// Nothing here is exported; the public surface is include/daliid_b200.h.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/daliid_b200.h"

namespace dali {

// ---------------------------------------------------------------------------
// Order-preserving key of an fp32 distance (the canonical order of SURVEY 8c):
// ascending value, -0.0 == +0.0, every NaN after +inf.  Largest real key is
// key(+inf) = 0xFF800000, NaN maps to 0xFFFFFFFE so key+1 never wraps.
// ---------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t dist_key(float d) {
  d = d + 0.0f;  // -0.0 -> +0.0 (IEEE; kept because fast-math is never enabled)
#ifdef __CUDA_ARCH__
  uint32_t b = __float_as_uint(d);
#else
  union { float f; uint32_t u; } cv; cv.f = d; uint32_t b = cv.u;
#endif
  uint32_t k = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
  return (d != d) ? 0xFFFFFFFEu : k;
}

__host__ __device__ __forceinline__ float key_to_dist(uint32_t k) {
  uint32_t b = (k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k;
#ifdef __CUDA_ARCH__
  return __uint_as_float(b);
#else
  union { float f; uint32_t u; } cv; cv.u = b; return cv.f;
#endif
}

// x / n for a small integer n (the mean of n matrices), correctly rounded like __fdiv_rn: for
// |x| in [2^-100, 2^100) two FMAs refine x * RN(1/n) -- q0 is faithful, r = x - q0 n is exact, and
// RN(q0 + r RN(1/n)) is then the correctly rounded quotient (Markstein); checked against
// __fdiv_rn over all 2^32 operands for n = 2 .. 8 by dali_selftest_mean_division.  Zero,
// subnormal-range, huge, infinite and NaN operands take the IEEE division.  ~6 instructions
// instead of ~20 with a call: the fused-mean epilogue and fuse.cu divide every element.
__device__ __forceinline__ float div_small_int(float x, float n, float rn) {
  const float ax = fabsf(x);
  if (ax >= 7.888609052210118e-31f && ax < 1.2676506002282294e30f) {
    const float q0 = __fmul_rn(x, rn);
    const float r = __fmaf_rn(-q0, n, x);
    return __fmaf_rn(r, rn, q0);
  }
  return __fdiv_rn(x, n);
}

__host__ __device__ __forceinline__ uint64_t composite(uint32_t key, uint32_t gid) {
  return (static_cast<uint64_t>(key) << 32) | gid;
}

// ---------------------------------------------------------------------------
// Context
// ---------------------------------------------------------------------------
struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
};

enum WsSlot {
  WS_QN = 0,   // normalised queries (fp32, padded rows) / hi+lo planes
  WS_GN,       // normalised gallery
  WS_QN16,     // bf16 hi / residual planes (TF32C)
  WS_GN16,
  WS_QIN,      // staged host queries
  WS_GIN,      // staged host gallery
  WS_QNORM,    // row norms
  WS_GNORM,
  WS_DIST,     // internal distance matrix
  WS_KEYS,     // match keys
  WS_COUNTS,   // match counts
  WS_RANKS,    // sorted kept ranks
  WS_AP,       // per-query AP (fp32)
  WS_FIRST,    // per-query first rank
  WS_CMC,      // cmc histogram + num_valid
  WS_STAGE_A,  // generic staging (host matrices etc.)
  WS_STAGE_B,
  WS_STAGE_C,
  WS_TOPK_D,
  WS_TOPK_I,
  WS_PTRS,     // device arrays of pointers (fusion)
  WS_MISC,
  WS_CAND,      // fused top-k: candidate lists [Q][cap] uint64
  WS_CAND_CNT,  // [Q] int32
  WS_THR,       // [Q] fp32 running k-th best distance
  WS_FLAG,      // overflow flag
  WS_RR_OD,     // re-ranking: normalised squared [N,N] matrix
  WS_RR_MISC,
  WS_RR_RANK,
  WS_RR_V0,
  WS_RR_V,
  WS_RR_CSC,
  WS_RR_TEMP,
  WS_MR_T,      // meta-recognition fusion: cleaned transpose [G,Q]
  WS_MR_MISC,   // kill / low lists, Weibull parameters
  WS_SORT,      // full row ordering: composite keys, payload, radix scratch
  WS_FZ_THR,    // fused counting: per-query sorted thresholds [M] fp32
  WS_FZ_SLOT,   // their match-list slots [M] int32
  WS_FZ_HIST,   // bucket histogram [M] int32
  WS_COUNT_
};

// ---- fused distance + positive-rank counting (distmat_umma2.cu: kBand / kCount) ----------------
// Operands of the fused path are written IDENTITY-SORTED by the preparation kernels (rows of the
// query planes in `qorder`, rows of the gallery planes in the plan's `order`), so that the matches
// of a 256-query tile sit in a handful of adjacent column tiles (the band).
struct FusedArgs {
  const uint32_t *tiles = nullptr;  // explicit tile list of this launch, (m << 16) | n
  int num_list = 0;
  const uint32_t *band = nullptr;   // kCount: the band tiles, whose distances wait in `scratch`
  int num_band = 0;
  float *scratch = nullptr;         // [num_band][256][256] fp32
  const int32_t *qorder = nullptr;  // [Q] query index of sorted row r
  const int64_t *off = nullptr;     // [Q + 1] match-list offsets (plan)
  const int64_t *lo = nullptr;      // [Q] first sorted-gallery column of the query's identity
  const int32_t *nv = nullptr;      // [Q] valid positives (they come first in the match list)
  const int32_t *gid = nullptr;     // [M] global gallery id of a match
  const int32_t *slot_of_seg = nullptr;  // [M] match-list slot of the i-th column of the segment
  const int32_t *order = nullptr;   // [G] local gallery row of sorted column j
  uint32_t *keys = nullptr;         // [M] order keys of the matches' distances (kBand writes them)
  const float *sorted_thr = nullptr;   // [M] per query: distances of its valid positives, ascending
  const int32_t *sorted_slot = nullptr;  // [M] match-list slot of the k-th smallest
  int32_t *hist = nullptr;          // [M] per query and sorted position k: columns with exactly k
                                    // thresholds (of the same pass of 31) <= them (red.add per tile)
  int32_t *counts = nullptr;        // [M] match-list order: tie corrections (kCount), then the prefix
                                    // sums of hist are added (launch_fused_prefix)
  int32_t *flag = nullptr;          // set when a threshold is not finite: the caller redoes the
                                    // evaluation through the materialised matrix
  int32_t gid_base = 0;             // global gallery id of local row 0 (sharded slabs)
  int dbg = 0;                      // DALI_FUSED_DBG (experiments): 1 skip the search, 2 skip the tile
};


}  // namespace dali

struct dali_rank_plan;
namespace dali { struct HostStager; }

struct dali_ctx {
  int device = 0;
  int num_sms = 0;
  int cc_major = 0, cc_minor = 0;
  int clock_khz = 1965000;  // SM clock the watchdogs convert milliseconds with
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  std::string err;
  dali::DevBuf ws[dali::WS_COUNT_];
  int32_t *pinned_flag = nullptr;  // fused counting: "a threshold was not finite" read back here
  void *pinned = nullptr;  // small pinned scratch for results
  size_t pinned_cap = 0;
  void *plan_stage = nullptr;  // pinned staging of the rank plan (async upload)
  size_t plan_stage_cap = 0;
  cudaEvent_t plan_stage_done = nullptr;  // the last upload out of plan_stage
  cudaStream_t plan_stream = nullptr;     // upload + expansion of a new plan run beside the contraction
  bool pool_ready = false;
  // H2D of host operands on side streams, overlapped with compute
  static constexpr int kCopyStreams = 4;
  cudaStream_t copy_streams[kCopyStreams] = {nullptr, nullptr, nullptr, nullptr};
  std::vector<cudaEvent_t> chunk_events;
  size_t next_event = 0;
  int h2d_streams = 1;  // DMA streams one H2D copy is split over (DALI_H2D_STREAMS)
  // operand preparation: between prep_defer_begin / prep_defer_end the first launch of the default
  // (fp16 planes) kind is held back and issued together with the second one as ONE launch
  bool prep_defer = false, prep_pending = false;
  int prep_kind = 0;
  unsigned prep_blocks = 0;
  alignas(8) unsigned char prep_params[192];
  // online choice between one and two DMA streams for pipelined host galleries: the first calls
  // of a context are timed (events around the copy/compute pipeline) with either setting
  int h2d_tune_calls = 0;        // pipelined calls seen so far
  int64_t h2d_tune_bytes = 0;    // gallery bytes of the calls being compared
  float h2d_tune_ms[2] = {0.f, 0.f};  // best pipeline time with 1 / 2 streams
  cudaEvent_t h2d_ev0 = nullptr, h2d_ev1 = nullptr;
  int h2d_ev_streams = 0;        // setting the pending event pair was recorded with (0: none)
  cudaEvent_t handover = nullptr;  // dali_ctx_set_stream: the new stream waits for the old one
  dali::HostStager *stager = nullptr;  // pageable host operands: threaded copy into pinned slots
  // timing
  bool timing = false;
  int t_launches[DALI_K_COUNT_] = {0};
  std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> t_pending;
  std::vector<cudaEvent_t> t_pool;
  float t_ms[DALI_K_COUNT_] = {0};
  // rank-plan cache: the plan depends on the label arrays only, and evaluation code calls the
  // path many times with the same query / gallery sets.  The last plan is kept and
  // reused when the next call's labels compare equal byte for byte.
  dali_rank_plan *cached_plan = nullptr;
  bool plan_cache = true;
  int64_t plan_cache_hits = 0;
  int64_t launches = 0;
  int64_t fallbacks = 0;  // fused calls that had to be redone through the materialised path
  bool fused_count = false;  // evaluations may take the fused distance + counting path (opt-in:
                             // measured slower than the matrix path at the Market shapes)
  int64_t fused_calls = 0;   // evaluations that did
  // opt-in dynamic shared memory already granted per kernel on THIS device (the attribute is
  // per device, so it cannot be a process-wide static)
  std::unordered_map<const void *, size_t> func_smem;
  // tensor-map encoder (driver entry point, resolved lazily)
  void *encode_tiled = nullptr;
};

struct dali_rank_plan {
  dali_ctx *ctx = nullptr;
  int refs = 1;                       // the creator, plus the context's cache while it holds it
  std::vector<int32_t> labels;        // q_pid | g_pid | q_cam | g_cam as passed (cache key)
  cudaEvent_t ready = nullptr;        // device image complete (expansion kernel done)
  int64_t Q = 0, G = 0, M = 0;
  int max_m = 0;   // largest number of matches (same identity) of one query
  int max_nv = 0;  // upper bound of the valid positives of one query (== max_m; the exact
                   // per-query numbers live on the device, written by the expansion kernel)
  std::vector<int64_t> h_off;  // [Q+1] offsets into the match arrays (host copy)
  // device image, one stream-ordered allocation:
  //   off [Q+1] i64 | lo [Q] i64 | nv [Q] i32 | njunk [Q] i32 | q_cam [Q] i32 | gid [M] i32 |
  //   order [G] i32 (gallery ids sorted by identity) | g_cam [G] i32
  void *d_block = nullptr;
  int64_t *d_off = nullptr;
  int64_t *d_lo = nullptr;    // start of each query's range inside `order`
  int32_t *d_nv = nullptr;    // number of valid positives (they come first in gid)
  int32_t *d_njunk = nullptr;
  int32_t *d_qcam = nullptr;
  int32_t *d_gid = nullptr;   // [M] gallery id of each match (valid ascending, then junk)
  int32_t *d_order = nullptr;
  int32_t *d_gcam = nullptr;
  int32_t *d_slot = nullptr;  // [M] match-list slot (relative to off[q]) of the i-th item of the query's
                              // identity segment in `order` (written by the expansion kernel)
  std::vector<int64_t> h_lo;  // [Q] host copy of lo
  // fused distance + counting (built on first use, dali::fused_plan_setup): queries sorted by the
  // position of their identity in `order`, the band tiles and all the other tiles
  bool fz_ready = false;
  void *d_fz = nullptr;
  int32_t *d_qorder = nullptr;   // [Q]
  uint32_t *d_band = nullptr;    // [n_band] (m << 16) | n
  uint32_t *d_main = nullptr;    // [n_main]
  int n_band = 0, n_main = 0;
};

namespace dali {

int set_err(dali_ctx *ctx, int code, const std::string &msg);
int ws_ensure(dali_ctx *ctx, int slot, size_t bytes, void **out);
// raise a kernel's dynamic shared-memory limit on the context's device (cached per context)
int ensure_dyn_smem(dali_ctx *ctx, const void *func, size_t bytes);

// Switches to the context's device for the duration of an entry point, restores the caller's.
struct DeviceGuard {
  int prev = -1;
  int enter(dali_ctx *ctx);
  ~DeviceGuard();
};

// RAII-less timing helpers: call before/after a kernel launch on ctx->stream.
struct KTimer {
  dali_ctx *ctx;
  int slot;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  KTimer(dali_ctx *c, int s);
  ~KTimer();
};

#define DALI_CUDA_OK(ctx, expr)                                                         \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess)                                                              \
      return dali::set_err((ctx), DALI_ERR_CUDA,                                        \
                           std::string(#expr) + ": " + cudaGetErrorString(_e));         \
  } while (0)

// ---- kernel launchers (defined in the .cu files) ---------------------------
// normalize.cu
// perm (device, may be null): output row r is prepared from input row perm[r]
int launch_prep(dali_ctx *ctx, const float *x, int64_t n, int64_t d, int64_t ldx, float *plane0,
                float *plane1, int64_t ldo, int64_t d_pad, int64_t rows_pad, int do_normalize,
                int round_mode, float *norms, float *sq, void *hi16 = nullptr,
                void *lo16 = nullptr, const int32_t *perm = nullptr);
void prep_defer_begin(dali_ctx *ctx);
int prep_defer_end(dali_ctx *ctx);  // launches a held-back preparation, if any
float f16x3_hi_grid(int64_t d_pad);
int launch_selftest_div(dali_ctx *ctx, int n, unsigned long long *bad_dev);  // fuse.cu
// distmat_simt.cu
int launch_distmat_simt(dali_ctx *ctx, const float *qn, const float *gn, int64_t Q, int64_t G,
                        int64_t D, int64_t ldq, int64_t ldg, int metric, const float *qsq,
                        const float *gsq, float *out, int64_t ld);
// distmat_umma.cu
int launch_distmat_umma(dali_ctx *ctx, const float *q32, const float *g32, const void *q16,
                        const void *g16, int64_t Q, int64_t G, int64_t Dp, int64_t q_rows_pad,
                        int64_t g_rows_pad, int64_t g_row0, int precision, int metric,
                        const float *qsq, const float *gsq, float *out, int64_t ld, float *acc = nullptr,
                        int64_t ld_acc = 0, int acc_mode = 0, float acc_div = 1.0f);
// distmat_umma2.cu
int launch_distmat_filter_umma(dali_ctx *ctx, const float *q32, const float *g32, const void *q16,
                               const void *g16, int64_t Q, int64_t G, int64_t Dp,
                               int64_t q_rows_pad, int64_t g_rows_pad, int64_t g_row0,
                               int precision, int metric, const float *qsq, const float *gsq,
                               const float *thr, int32_t *cand_cnt, uint64_t *cand, int cap,
                               int largest, int direct, int32_t id_base);
bool fused_count_supports(int precision);
int launch_distmat_band_umma(dali_ctx *ctx, const float *q32, const float *g32, const void *q16,
                             const void *g16, int64_t Q, int64_t G, int64_t Dp, int64_t q_rows_pad,
                             int64_t g_rows_pad, int64_t g_row0, int precision, int metric,
                             const float *qsq, const float *gsq, const FusedArgs &fa);
int launch_distmat_count_umma(dali_ctx *ctx, const float *q32, const float *g32, const void *q16,
                              const void *g16, int64_t Q, int64_t G, int64_t Dp, int64_t q_rows_pad,
                              int64_t g_rows_pad, int64_t g_row0, int precision, int metric,
                              const float *qsq, const float *gsq, const FusedArgs &fa);
// rank.cu
int launch_fused_sort_thresholds(dali_ctx *ctx, const dali_rank_plan *plan, const uint32_t *keys,
                                 float *sorted_thr, int32_t *sorted_slot, int32_t *flag);
int launch_fused_prefix(dali_ctx *ctx, const dali_rank_plan *plan, const int32_t *hist,
                        const int32_t *sorted_slot, int32_t *counts, int pass);
int launch_plan_expand(dali_ctx *ctx, const dali_rank_plan *plan, cudaStream_t stream = nullptr);
int launch_rank_gather(dali_ctx *ctx, const dali_rank_plan *plan, const float *dist, int64_t ld,
                       int64_t g0, int64_t Gs, uint32_t *keys);
int launch_rank_count(dali_ctx *ctx, const dali_rank_plan *plan, const float *dist, int64_t ld,
                      int64_t g0, int64_t Gs, const uint32_t *keys, int32_t *counts);
int launch_rank_fused(dali_ctx *ctx, const dali_rank_plan *plan, const float *dist, int64_t ld,
                      int max_rank, int32_t *ranks_sorted, float *ap, int32_t *first_rank,
                      int32_t *cmc_cnt, int *done);
int launch_rank_finalize(dali_ctx *ctx, const dali_rank_plan *plan, const uint32_t *keys,
                         const int32_t *counts, int max_rank, int32_t *ranks_sorted, float *ap,
                         int32_t *first_rank, int32_t *cmc_cnt /* [max_rank+1], last = num_valid */);
// topk.cu
int launch_topk(dali_ctx *ctx, const float *dist, int64_t Q, int64_t G, int64_t ld, int k,
                int largest, const int32_t *col_ids, int32_t id_base, float *d_out,
                int32_t *i_out);
int launch_topk_merge(dali_ctx *ctx, const float *vals, const int32_t *ids, int parts, int64_t Q, int k,
                      int largest, float *d_out, int32_t *i_out);
int launch_topk_compact(dali_ctx *ctx, uint64_t *cand, int32_t *cand_cnt, float *thr, int64_t Q,
                        int cap, int k, int largest, int fixed_cnt, int32_t *overflow, float *d_out,
                        int32_t *i_out, int32_t *row_flags = nullptr);
// rerank.cu
int launch_rerank(dali_ctx *ctx, const float *qg, int64_t ld_qg, const float *qq, int64_t ld_qq,
                  const float *gg, int64_t ld_gg, int64_t Q, int64_t G, int k1, int k2, double lambda,
                  float *out, int64_t ld_out);
// mrfuse.cu
int launch_mrfuse(dali_ctx *ctx, const float *const *s, int n, int64_t Q, int64_t G, int64_t ld,
                  int topk, int use_columns, float killscale, double *out, int64_t ld_out,
                  double *fit_opt, float *small_opt, double *weights_opt);
// sortrows.cu
int launch_argsort_rows(dali_ctx *ctx, const float *dist, int64_t Q, int64_t G, int64_t ld,
                        int descending, int32_t *idx_out);
// roc.cu
int launch_roc_hist(dali_ctx *ctx, const float *dist, int64_t Q, int64_t G, int64_t ld, const int32_t *qpid,
                    const int32_t *gpid, int nbins, float lo, float hi, unsigned long long *pos_hist,
                    unsigned long long *neg_hist);
// fuse.cu
int launch_fuse(dali_ctx *ctx, const float *const *d_ptrs_dev, int n, const float *const *wq_dev,
                const float *const *wg_dev, float *out, int64_t Q, int64_t G, int64_t ld);

}  // namespace dali
