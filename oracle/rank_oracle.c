/*
 * CPU ORACLE (test infrastructure, NOT product code): plain-C restatement of the
 * compiled evaluator the reference reaches through
 *     torchreid.metrics.evaluate_rank(...)        validateModels.py:68-69,
 *                                                 evaluate.py:312-313,
 *                                                 evaluate_ensembled_models.py:324-325,
 *                                                 evaluateCleanATModels.py:266-267
 * i.e. torchreid/metrics/rank_cylib/rank_cy.pyx::eval_market1501_cy (third party,
 * not vendored in /root/reference, version unpinned -> PARITY UNPINNED; restated
 * from the upstream's published semantics).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this.  It doubles as the compiled CPU baseline
 * (kind "port"): like the upstream, the caller does the argsort with numpy and
 * hands the index matrix in; everything after that is the loop below.
 *
 * Semantics kept from the upstream Cython: C `float` state, full-length loops
 * over the gallery, AP summed sequentially in rank order with each term formed
 * in double and stored back to float, mAP a sequential float sum in query order.
 * One behaviour is *defined* here because it is undefined upstream (stale buffer
 * read): when fewer than max_rank gallery items survive junk removal the CMC row
 * saturates at its last value.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef __cplusplus
extern "C" {
#endif

/* returns 0 on success, 1 when no query is valid (upstream: AssertionError) */
int oracle_eval_market1501_cy(
    const int64_t *indices, /* [num_q, num_g] argsort of the distance rows     */
    const int64_t *q_pids, const int64_t *g_pids, const int64_t *q_camids,
    const int64_t *g_camids, int64_t num_q, int64_t num_g, int64_t max_rank,
    float *avg_cmc,      /* out [max_rank] (max_rank already clamped to num_g) */
    float *mAP_out,      /* out scalar                                         */
    float *all_AP,       /* out [num_q], 0 for invalid queries                 */
    int64_t *first_rank, /* out [num_q], -1 for invalid queries (1-based)      */
    int64_t *num_valid_out) {
  float *raw_cmc = (float *)calloc((size_t)num_g, sizeof(float));
  float *cmc = (float *)calloc((size_t)num_g, sizeof(float));
  float *tmp_cmc = (float *)calloc((size_t)num_g, sizeof(float));
  float *all_cmc = (float *)calloc((size_t)(num_q * max_rank), sizeof(float));
  float num_valid_q = 0.f;
  if (!raw_cmc || !cmc || !tmp_cmc || !all_cmc) return -1;

  for (int64_t q = 0; q < num_q; ++q) {
    const int64_t *order = indices + q * num_g;
    const int64_t q_pid = q_pids[q], q_camid = q_camids[q];
    int64_t num_g_real = 0;
    int meet_condition = 0;
    all_AP[q] = 0.f;
    first_rank[q] = -1;
    for (int64_t g = 0; g < num_g; ++g) {
      const int64_t o = order[g];
      if (g_pids[o] != q_pid || g_camids[o] != q_camid) {
        const float m = (g_pids[o] == q_pid) ? 1.f : 0.f; /* matches[q][g] */
        raw_cmc[num_g_real] = m;
        if (m > 1e-31f) {
          if (!meet_condition) first_rank[q] = num_g_real + 1;
          meet_condition = 1;
        }
        num_g_real++;
      }
    }
    if (!meet_condition) continue;

    /* cmc = clip(cumsum(raw_cmc), 1) */
    float run = 0.f;
    for (int64_t g = 0; g < num_g_real; ++g) {
      run += raw_cmc[g];
      tmp_cmc[g] = run;
      cmc[g] = run > 1.f ? 1.f : run;
    }
    for (int64_t r = 0; r < max_rank; ++r)
      all_cmc[q * max_rank + r] = r < num_g_real ? cmc[r] : cmc[num_g_real - 1];
    num_valid_q += 1.f;

    float num_rel = 0.f, tmp_cmc_sum = 0.f;
    for (int64_t g = 0; g < num_g_real; ++g) {
      tmp_cmc_sum += (tmp_cmc[g] / (g + 1.)) * raw_cmc[g]; /* double term, float store */
      num_rel += raw_cmc[g];
    }
    all_AP[q] = tmp_cmc_sum / num_rel;
  }

  *num_valid_out = (int64_t)num_valid_q;
  int rc = 0;
  if (num_valid_q > 0.f) {
    for (int64_t r = 0; r < max_rank; ++r) {
      float acc = 0.f;
      for (int64_t q = 0; q < num_q; ++q) acc += all_cmc[q * max_rank + r];
      avg_cmc[r] = acc / num_valid_q;
    }
    float mAP = 0.f;
    for (int64_t q = 0; q < num_q; ++q) mAP += all_AP[q];
    *mAP_out = mAP / num_valid_q;
  } else {
    rc = 1;
  }
  free(raw_cmc); free(cmc); free(tmp_cmc); free(all_cmc);
  return rc;
}

/* Stable argsort of one fp32 row: distance ascending, index ascending, NaN last,
 * -0.0 == +0.0 (numpy kind='stable' order).  Bottom-up merge sort. */
static int lt_f32(float a, float b) {
  /* numpy's float less-than for sorting: NaNs sort to the end */
  return a < b || (b != b && a == a);
}

void oracle_stable_argsort_rows(const float *dist, int64_t num_q, int64_t num_g,
                                int64_t ld, int64_t *indices) {
  int64_t *tmp = (int64_t *)malloc((size_t)num_g * sizeof(int64_t));
  for (int64_t q = 0; q < num_q; ++q) {
    const float *d = dist + q * ld;
    int64_t *a = indices + q * num_g, *b = tmp;
    for (int64_t i = 0; i < num_g; ++i) a[i] = i;
    for (int64_t w = 1; w < num_g; w *= 2) {
      for (int64_t lo = 0; lo < num_g; lo += 2 * w) {
        int64_t mid = lo + w < num_g ? lo + w : num_g;
        int64_t hi = lo + 2 * w < num_g ? lo + 2 * w : num_g;
        int64_t i = lo, j = mid, k = lo;
        while (i < mid && j < hi) b[k++] = lt_f32(d[a[j]], d[a[i]]) ? a[j++] : a[i++];
        while (i < mid) b[k++] = a[i++];
        while (j < hi) b[k++] = a[j++];
      }
      int64_t *t = a; a = b; b = t;
    }
    if (a != indices + q * num_g) memcpy(indices + q * num_g, a, (size_t)num_g * sizeof(int64_t));
  }
  free(tmp);
}

#ifdef __cplusplus
}
#endif
