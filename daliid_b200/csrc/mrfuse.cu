// Meta-recognition score fusion: per gallery column, a 2-parameter Weibull is
// fitted (maximum likelihood, Newton on the shape) to the high tail of the column's scores, every
// score is turned into a weight by that Weibull's CDF, and the models are fused as the weighted mean.
//
// replaces  Meta_Recognition.mrfuse / metarec       evaluate.py:583-627 (call site, commented: :277),
//                                                   evaluate_ensembled_models.py:593-637
//           libmr.FitHigh / _weibullFitting / _fit  evaluate.py:429-432, 475-493, 531-580
//           libmr.wscore (Weibull CDF)              evaluate.py:434-473
//
// Data flow per model (S = [Q,G] fp32 similarity scores; T = its cleaned transpose [G,Q]):
//   top-`topk` of every row of S (torch.topk, NaN counts as largest)      -> kill list
//   T[g][q] = nan_to_num(S[q][g]); killed entries x -> x - killscale*x    (transpose + scatter)
//   (topk+2) smallest of every row of T: the tail is everything but the topk+1 smallest,
//       `small` is the (topk+2)-th smallest = the last element of the reference's sorted tail
//   fit: one CTA per gallery column, logs staged in shared memory (recomputed per step for very
//        long columns), Newton in fp64
//   fuse: one pass over the n score matrices, fp64 out
// The reference sorts the tail (torch.topk of Q-topk-1 values); only sums over the tail enter the
// fit, so no sort is needed: excluding the topk+1 smallest by index gives the same multiset.
//
// Differences to the reference that are below 1e-12 relative: the Newton loop runs per column
// until |dk| < eps plus two more steps (the reference keeps updating all columns until the
// slowest one converged); sums are tree reductions.  fp32 log / mean are correctly rounded from
// fp64 (torch's CPU fp32 log is within 1 ulp of that), which moves the fitted shape by ~1e-7
// relative -- the reproducibility limit of the reference's own fp32 intermediates.
#include <cfloat>
#include <cmath>
#include <cstdlib>

#include "common.cuh"

namespace dali {

namespace {

constexpr int kFitThreads = 256;
constexpr int kIters = 100;
constexpr double kEps = 1e-6;

__device__ __forceinline__ float clean(float x) {  // torch.nan_to_num(x, 0)
  if (isnan(x)) return 0.f;
  if (isinf(x)) return x > 0 ? FLT_MAX : -FLT_MAX;
  return x;
}

// T[g][q] = clean(S[q][g])
__global__ void __launch_bounds__(256)
mr_transpose_kernel(const float *__restrict__ S, int64_t ld, int64_t Q, int64_t G,
                    float *__restrict__ T, int64_t ldT, int raw) {
  __shared__ float tile[32][33];
  const int64_t g0 = static_cast<int64_t>(blockIdx.x) * 32, q0 = static_cast<int64_t>(blockIdx.y) * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    const int64_t q = q0 + r, g = g0 + tx;
    tile[r][tx] = (q < Q && g < G) ? S[q * ld + g] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    const int64_t g = g0 + r, q = q0 + tx;
    if (g < G && q < Q) T[g * ldT + q] = raw ? tile[tx][r] : clean(tile[tx][r]);
  }
}

// rows of S (use_columns == 0): kill list idx [Q][topk] holds gallery ids -> T[idx][q]
// rows of T (use_columns != 0): kill list idx [G][topk] holds query ids   -> T[g][idx]
__global__ void mr_kill_kernel(const int32_t *__restrict__ idx, int64_t rows, int topk, int by_cols,
                               const float *__restrict__ S, int64_t ld, float killscale,
                               float *__restrict__ T, int64_t ldT) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= rows * topk) return;
  const int64_t r = i / topk;
  const int64_t c = idx[i];
  const int64_t q = by_cols ? c : r, g = by_cols ? r : c;
  const float x = S[q * ld + g];
  T[g * ldT + q] = clean(x - killscale * x);
}

__global__ void mr_clean_kernel(float *__restrict__ T, int64_t n) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) T[i] = clean(T[i]);
}

struct FitOut {
  double *shape, *scale;   // [G] the fit (NaN / 0 as the reference leaves them)
  double *kk, *sign;       // [G] 1/(1/shape) and sign(1/shape)*sign(scale), as Weibull.cdf uses them
  float *small;            // [G]
};

__device__ __forceinline__ double block_sum(double v, double *red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();  // red[] free again
  if (l == 0) red[w] = v;
  __syncthreads();
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < kFitThreads / 32; ++i) s += red[i];
  return s;
}

__device__ __forceinline__ double sgn(double x) { return x > 0 ? 1.0 : (x < 0 ? -1.0 : (x == 0 ? 0.0 : x)); }

// one CTA per gallery column g; T row g holds the column's Q cleaned scores.  STAGED: ln(x) of the
// column is kept in shared memory (8 B per query, Q <= ~28k); otherwise it is recomputed from the
// row every Newton step (any Q; one more fp64 log per element and step).  The topk+1 smallest
// entries, which are not part of the tail, are marked with a NaN.
template <bool STAGED>
__global__ void __launch_bounds__(kFitThreads)
mr_fit_kernel(float *__restrict__ T, int64_t ldT, int64_t Q, const float *__restrict__ low_v,
              const int32_t *__restrict__ low_i, int nlow /* topk+2 */, FitOut out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *lnd = reinterpret_cast<double *>(smem_raw);
  __shared__ double red[kFitThreads / 32];
  const int64_t g = blockIdx.x;
  const int tid = threadIdx.x;
  const float small = low_v[g * nlow + (nlow - 1)];
  float *row = T + g * ldT;
  auto ln_of = [&](float t) {  // sortedTensor + translateAmount - smallScoreTensor, then the log
    return log(static_cast<double>((t + 1.0f) - small));
  };
  if (STAGED) {
    for (int64_t q = tid; q < Q; q += kFitThreads) lnd[q] = ln_of(row[q]);
    __syncthreads();
    if (tid < nlow - 1) lnd[low_i[g * nlow + tid]] = nan("");
  } else {
    if (tid < nlow - 1) row[low_i[g * nlow + tid]] = __int_as_float(0x7fc00000);  // this CTA owns row g
  }
  __syncthreads();
  auto ln_at = [&](int64_t q) { return STAGED ? lnd[q] : ln_of(row[q]); };
  const double n_tail = static_cast<double>(Q - (nlow - 1));
  double sl = 0.0;
  for (int64_t q = tid; q < Q; q += kFitThreads) {
    const double l = ln_at(q);
    if (!isnan(l)) sl += static_cast<double>(static_cast<float>(l));  // the reference's log is fp32
  }
  sl = block_sum(sl, red);
  const double mean_ln = static_cast<double>(static_cast<float>(sl / n_tail));  // torch.mean of fp32

  double k = 1.0, k_prev = 1.0;
  bool open = true, saw_nan = false;
  int extra = 0;
  for (int it = 0; it < kIters; ++it) {
    double fg = 0.0, ff = 0.0, fp = 0.0;
    for (int64_t q = tid; q < Q; q += kFitThreads) {
      const double ld_ = ln_at(q);
      if (isnan(ld_)) continue;
      const double e = exp(k * ld_);
      const double l = static_cast<double>(static_cast<float>(ld_));
      const double t = e * l;
      fg += e;
      ff += t;
      fp += t * l;
    }
    fg = block_sum(fg, red);
    ff = block_sum(ff, red);
    fp = block_sum(fp, red);
    // every thread holds the same sums: the Newton step is computed redundantly (no broadcast)
    const double r = ff / fg;
    const double f = r - mean_ln - 1.0 / k;
    const double f_prime = (fp / fg - r * r) + 1.0 / (k * k);
    k -= f / f_prime;
    if (open && isnan(f)) saw_nan = true;
    if (fabs(k - k_prev) < kEps) open = false;
    k_prev = k;
    if (!open && ++extra > 2) break;
  }
  double shape = 0.0, scale = 0.0;
  if (!open) {
    double fg = 0.0;
    for (int64_t q = tid; q < Q; q += kFitThreads) {
      const double ld_ = ln_at(q);
      if (!isnan(ld_)) fg += exp(k * ld_);
    }
    fg = block_sum(fg, red);
    shape = k;
    scale = pow(fg / n_tail, 1.0 / k);
  } else if (saw_nan) {
    shape = scale = nan("");
  }
  if (tid == 0) {
    out.shape[g] = shape;
    out.scale[g] = scale;
    const double expo = 1.0 / shape;
    out.kk[g] = 1.0 / expo;
    out.sign[g] = sgn(expo) * sgn(scale);
    out.small[g] = small;
  }
}

struct FuseIn {
  const float *s[3];
  const double *scale[3], *kk[3], *sign[3];
  const float *small[3];
  int n;
};

__device__ __forceinline__ double weibull_weight(float s, float small, double scale, double kk, double sign) {
  float d = (s + 1.0f) - small;
  d = d < 0.f ? 0.f : d;  // clamp(min=0), NaN stays NaN
  const double y = static_cast<double>(d) / scale;
  const double z = pow(y, kk);
  const double v = 1.0 - exp(-z);
  double w = sign * (v - 0.5) + 0.5;
  if (isnan(w)) w = 0.0;  // nan_to_num
  else if (isinf(w)) w = w > 0 ? DBL_MAX : -DBL_MAX;
  return w;
}

__global__ void __launch_bounds__(256)
mr_fuse_kernel(FuseIn in, int64_t ld, int64_t Q, int64_t G, double *__restrict__ out, int64_t ld_out,
               double *__restrict__ weights, int64_t w_stride) {
  const int64_t g = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (g >= G) return;
  double scale[3], kk[3], sign[3];
  float small[3];
#pragma unroll
  for (int m = 0; m < 3; ++m)
    if (m < in.n) {
      scale[m] = in.scale[m][g]; kk[m] = in.kk[m][g]; sign[m] = in.sign[m][g]; small[m] = in.small[m][g];
    }
  for (int64_t q = blockIdx.y; q < Q; q += gridDim.y) {
    double num = 0.0, den = 0.0;
#pragma unroll
    for (int m = 0; m < 3; ++m)
      if (m < in.n) {
        const float s = in.s[m][q * ld + g];
        const double w = weibull_weight(s, small[m], scale[m], kk[m], sign[m]);
        if (weights) weights[m * w_stride + q * ld_out + g] = w;
        const double t = w * static_cast<double>(s);
        num = m == 0 ? t : num + t;
        den = m == 0 ? w : den + w;
      }
    out[q * ld_out + g] = num / den;
  }
}

}  // namespace

// s[m]: device [Q, ld] fp32; out: device [Q, ld_out] fp64; fit_opt: device [n][G][2] (shape, scale);
// small_opt: device [n][G]; weights_opt: device [n][Q][ld_out] fp64
int launch_mrfuse(dali_ctx *ctx, const float *const *s, int n, int64_t Q, int64_t G, int64_t ld,
                  int topk, int use_columns, float killscale, double *out, int64_t ld_out,
                  double *fit_opt, float *small_opt, double *weights_opt) {
  if (n < 1 || n > 3) return set_err(ctx, DALI_ERR_INVALID, "mrfuse: 1 <= n <= 3 score matrices");
  if (topk < 1 || topk + 2 > 128) return set_err(ctx, DALI_ERR_INVALID, "mrfuse: 1 <= topk <= 126");
  if (Q < topk + 2 || G < 1 || (!use_columns && G < topk))
    return set_err(ctx, DALI_ERR_INVALID, "mrfuse: needs Q >= topk+2 scores per gallery column (tail = Q-topk-1 >= 1) "
                                          "and G >= topk");
  const char *env_staged = getenv("DALI_MRFUSE_STAGED");  // 0: force the recompute path (cross-check)
  const bool staged = static_cast<size_t>(Q) * 8 <= 225 * 1024 && !(env_staged && atoi(env_staged) == 0);
  const size_t smem = staged ? static_cast<size_t>(Q) * 8 : 0;
  KTimer timer(ctx, DALI_K_MRFUSE);
  const int64_t ldT = (Q + 3) / 4 * 4;
  const int nlow = topk + 2;
  const int64_t kill_rows = use_columns ? G : Q;
  void *p;
  int rc;
  if ((rc = ws_ensure(ctx, WS_MR_T, sizeof(float) * G * ldT, &p))) return rc;
  float *T = static_cast<float *>(p);
  // misc: kill list values/ids | low values/ids | per model: shape, scale, kk, sign (fp64), small
  const size_t kill_n = static_cast<size_t>(kill_rows) * topk, low_n = static_cast<size_t>(G) * nlow;
  const size_t par_bytes = static_cast<size_t>(G) * (4 * sizeof(double) + sizeof(float));
  size_t off_kill_v = 0, off_kill_i = off_kill_v + sizeof(float) * kill_n;
  size_t off_low_v = off_kill_i + sizeof(int32_t) * kill_n, off_low_i = off_low_v + sizeof(float) * low_n;
  size_t off_par = (off_low_i + sizeof(int32_t) * low_n + 15) / 16 * 16;
  if ((rc = ws_ensure(ctx, WS_MR_MISC, off_par + 3 * (par_bytes + 16), &p))) return rc;
  char *base = static_cast<char *>(p);
  float *kill_v = reinterpret_cast<float *>(base + off_kill_v);
  int32_t *kill_i = reinterpret_cast<int32_t *>(base + off_kill_i);
  float *low_v = reinterpret_cast<float *>(base + off_low_v);
  int32_t *low_i = reinterpret_cast<int32_t *>(base + off_low_i);
  FitOut fo[3];
  FuseIn fi;
  fi.n = n;
  for (int m = 0; m < 3; ++m) {
    char *b = base + off_par + m * ((par_bytes + 15) / 16 * 16);
    fo[m].shape = reinterpret_cast<double *>(b);
    fo[m].scale = fo[m].shape + G;
    fo[m].kk = fo[m].scale + G;
    fo[m].sign = fo[m].kk + G;
    fo[m].small = reinterpret_cast<float *>(fo[m].sign + G);
    fi.s[m] = m < n ? s[m] : nullptr;
    fi.scale[m] = fo[m].scale; fi.kk[m] = fo[m].kk; fi.sign[m] = fo[m].sign; fi.small[m] = fo[m].small;
  }
  if ((rc = ensure_dyn_smem(ctx, reinterpret_cast<const void *>(&mr_fit_kernel<true>), smem))) return rc;
  for (int m = 0; m < n; ++m) {
    dim3 tg(static_cast<unsigned>((G + 31) / 32), static_cast<unsigned>((Q + 31) / 32));
    ctx->launches++;
    mr_transpose_kernel<<<tg, 256, 0, ctx->stream>>>(s[m], ld, Q, G, T, ldT, use_columns ? 1 : 0);
    DALI_CUDA_OK(ctx, cudaGetLastError());
    if (!use_columns) {
      if ((rc = launch_topk(ctx, s[m], Q, G, ld, topk, 1, nullptr, 0, kill_v, kill_i))) return rc;
    } else {
      // per gallery column: rows of the raw transpose (a NaN counts as the largest score, as in
      // torch.topk); T is cleaned after the kill, like the reference's nan_to_num
      if ((rc = launch_topk(ctx, T, G, Q, ldT, topk, 1, nullptr, 0, kill_v, kill_i))) return rc;
      ctx->launches++;
      mr_clean_kernel<<<static_cast<unsigned>((G * ldT + 255) / 256), 256, 0, ctx->stream>>>(T, G * ldT);
      DALI_CUDA_OK(ctx, cudaGetLastError());
    }
    ctx->launches++;
    mr_kill_kernel<<<static_cast<unsigned>((kill_n + 255) / 256), 256, 0, ctx->stream>>>(
        kill_i, kill_rows, topk, use_columns, s[m], ld, killscale, T, ldT);
    DALI_CUDA_OK(ctx, cudaGetLastError());
    if ((rc = launch_topk(ctx, T, G, Q, ldT, nlow, 0, nullptr, 0, low_v, low_i))) return rc;
    ctx->launches++;
    if (staged)
      mr_fit_kernel<true><<<static_cast<unsigned>(G), kFitThreads, smem, ctx->stream>>>(T, ldT, Q, low_v, low_i, nlow, fo[m]);
    else
      mr_fit_kernel<false><<<static_cast<unsigned>(G), kFitThreads, 0, ctx->stream>>>(T, ldT, Q, low_v, low_i, nlow, fo[m]);
    DALI_CUDA_OK(ctx, cudaGetLastError());
    if (fit_opt) {
      DALI_CUDA_OK(ctx, cudaMemcpy2DAsync(fit_opt + static_cast<size_t>(m) * G * 2, 2 * sizeof(double), fo[m].shape,
                                          sizeof(double), sizeof(double), G, cudaMemcpyDefault, ctx->stream));
      DALI_CUDA_OK(ctx, cudaMemcpy2DAsync(fit_opt + static_cast<size_t>(m) * G * 2 + 1, 2 * sizeof(double),
                                          fo[m].scale, sizeof(double), sizeof(double), G,
                                          cudaMemcpyDefault, ctx->stream));
    }
    if (small_opt)
      DALI_CUDA_OK(ctx, cudaMemcpyAsync(small_opt + static_cast<size_t>(m) * G, fo[m].small, sizeof(float) * G,
                                        cudaMemcpyDefault, ctx->stream));
  }
  dim3 fg(static_cast<unsigned>((G + 255) / 256),
          static_cast<unsigned>(std::min<int64_t>(Q, std::max<int64_t>(1, 8 * ctx->num_sms / ((G + 255) / 256)))));
  ctx->launches++;
  mr_fuse_kernel<<<fg, 256, 0, ctx->stream>>>(fi, ld, Q, G, out, ld_out, weights_opt,
                                              static_cast<int64_t>(Q) * ld_out);
  DALI_CUDA_OK(ctx, cudaGetLastError());
  return DALI_OK;
}

}  // namespace dali
