// ROC over all Q x G pairs as ONE streaming pass over the distance matrix (the
// verification branch of evaluateCleanATModels.py:276-292 -- label = same identity, score =
// 1.0 - distmat / 2.0, sklearn.metrics.roc_curve over Q*G scores; 53.6 M pairs at the Market shape).
//
// An ROC point is (#negatives, #positives) with score >= t.  Instead of sorting the scores, the
// kernel histograms them: NB uniform bins over [lo, hi] per class, bin = floor((s - lo) * NB / (hi -
// lo)) in fp32 (monotone in s), so the suffix sums of the two histograms are EXACT points of the
// curve at NB thresholds (the smallest score of every bin).  HBM bound: 4 B per pair, no sort, no
// Q x G temporaries (the reference builds two int64 label matrices and a flattened score copy).
//
// One CTA per (row, column split): the row's identity is one register, gallery identities are
// read coalesced (L2 resident), counts go to a shared-memory histogram of the bins this CTA
// touches most (a window of 4096 bins around the row's first score; scores of one row cluster) and
// to global atomics otherwise; the window is flushed with one global add per non-empty bin.
#include "common.cuh"

namespace dali {
namespace {

constexpr int kRocThreads = 256;
constexpr int kRocWindow = 4096;  // bins per class held in shared memory

__global__ void __launch_bounds__(kRocThreads)
roc_hist_kernel(const float *__restrict__ dist, int64_t G, int64_t ld, const int32_t *__restrict__ qpid,
                const int32_t *__restrict__ gpid, int nbins, float lo, float scale, int64_t per_split,
                unsigned long long *__restrict__ pos_hist, unsigned long long *__restrict__ neg_hist) {
  __shared__ uint32_t s_hist[2][kRocWindow];
  __shared__ int s_w0;
  const int64_t q = blockIdx.x;
  const int64_t c0 = static_cast<int64_t>(blockIdx.y) * per_split;
  const int64_t c1 = c0 + per_split < G ? c0 + per_split : G;
  if (c0 >= c1) return;
  const float *row = dist + q * ld;
  const int32_t pid = qpid[q];
  auto bin_of = [&](float d) {
    const float s = __fsub_rn(1.0f, __fmul_rn(d, 0.5f));   // the reference's 1.0 - distmat / 2.0 (fp32)
    const float t = (s - lo) * scale;
    int b = t >= 0.f ? static_cast<int>(fminf(t, 2.0e9f)) : 0;   // NaN: below every threshold
    if (s != s) b = 0;
    return b < nbins ? b : nbins - 1;
  };
  for (int i = threadIdx.x; i < 2 * kRocWindow; i += kRocThreads) (&s_hist[0][0])[i] = 0u;
  if (threadIdx.x == 0) {
    const int b = bin_of(row[c0]);
    s_w0 = max(0, min(nbins - kRocWindow, b - kRocWindow / 2));
  }
  __syncthreads();
  const int w0 = s_w0;
  for (int64_t c = c0 + threadIdx.x; c < c1; c += kRocThreads) {
    const int b = bin_of(__ldg(row + c));
    const int cls = __ldg(gpid + c) == pid ? 1 : 0;
    const int rel = b - w0;
    if (rel >= 0 && rel < kRocWindow) atomicAdd(&s_hist[cls][rel], 1u);
    else atomicAdd((cls ? pos_hist : neg_hist) + b, 1ull);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * kRocWindow; i += kRocThreads) {
    const int cls = i / kRocWindow, rel = i - cls * kRocWindow;
    const uint32_t v = s_hist[cls][rel];
    if (v && w0 + rel < nbins) atomicAdd((cls ? pos_hist : neg_hist) + w0 + rel, static_cast<unsigned long long>(v));
  }
}

}  // namespace

// pos_hist / neg_hist: device uint64 [nbins], zeroed by the caller.
int launch_roc_hist(dali_ctx *ctx, const float *dist, int64_t Q, int64_t G, int64_t ld, const int32_t *qpid,
                    const int32_t *gpid, int nbins, float lo, float hi, unsigned long long *pos_hist,
                    unsigned long long *neg_hist) {
  if (Q == 0 || G == 0) return DALI_OK;
  int64_t nsplit = std::max<int64_t>(1, std::min<int64_t>((G + 8191) / 8192, (8ll * ctx->num_sms + Q - 1) / Q));
  int64_t per_split = (G + nsplit - 1) / nsplit;
  nsplit = (G + per_split - 1) / per_split;
  if (Q > 2147483647ll || nsplit > 65535) return set_err(ctx, DALI_ERR_UNSUPPORTED, "roc: matrix too large for one launch");
  const float scale = static_cast<float>(nbins) / (hi - lo);
  KTimer t(ctx, DALI_K_TOPK);
  roc_hist_kernel<<<dim3(static_cast<unsigned>(Q), static_cast<unsigned>(nsplit)), kRocThreads, 0, ctx->stream>>>(
      dist, G, ld, qpid, gpid, nbins, lo, scale, per_split, pos_hist, neg_hist);
  DALI_CUDA_OK(ctx, cudaGetLastError());
  return DALI_OK;
}

}  // namespace dali
