// Multi-model distance fusion: element-wise, HBM bound, with the
// reference's exact operation order so the fused matrix is bit-identical to numpy/torch.
//
// mean      ((d0+d1)+d2)/n             evaluate.py:278, evaluate_ensembled_models.py:313,
//                                      evaluateCleanATModels.py:127
// weighted  w_m[i,j] = max(wq_m[i], wg_m[j]);  (w0*d0 + w1*d1 + ..)/(w0 + w1 + ..)
//                                      evaluateCleanATModels.py:154-157,193-196,230-233
#include <algorithm>

#include "common.cuh"

namespace dali {

namespace {

constexpr int kMaxFuse = 8;

struct FuseParams {
  const float *d[kMaxFuse];
  const float *wq[kMaxFuse];
  const float *wg[kMaxFuse];
  float *out;
  int n;
  int weighted;
  int64_t Q, G, ld;
};

__device__ __forceinline__ float fuse_one(const FuseParams &p, const float (&v)[kMaxFuse],
                                          const float (&wqv)[kMaxFuse], int64_t col) {
  if (!p.weighted) {
    float acc = v[0];
#pragma unroll
    for (int m = 1; m < kMaxFuse; ++m)
      if (m < p.n) acc = __fadd_rn(acc, v[m]);
    return __fdiv_rn(acc, static_cast<float>(p.n));
  }
  float w = fmaxf(wqv[0], __ldg(p.wg[0] + col));
  // torch.maximum propagates NaN; fmaxf does not
  if (wqv[0] != wqv[0] || __ldg(p.wg[0] + col) != __ldg(p.wg[0] + col)) w = NAN;
  float num = __fmul_rn(w, v[0]);
  float den = w;
#pragma unroll
  for (int m = 1; m < kMaxFuse; ++m) {
    if (m < p.n) {
      const float g = __ldg(p.wg[m] + col);
      float wm = fmaxf(wqv[m], g);
      if (wqv[m] != wqv[m] || g != g) wm = NAN;
      num = __fadd_rn(num, __fmul_rn(wm, v[m]));
      den = __fadd_rn(den, wm);
    }
  }
  return __fdiv_rn(num, den);
}

template <int VEC>
__global__ void __launch_bounds__(256) fuse_kernel(FuseParams p) {
  const int64_t q = blockIdx.y;
  float wqv[kMaxFuse];
#pragma unroll
  for (int m = 0; m < kMaxFuse; ++m) wqv[m] = (p.weighted && m < p.n) ? __ldg(p.wq[m] + q) : 0.f;
  const int64_t nvec = (p.G + VEC - 1) / VEC;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < nvec;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t c = i * VEC;
    if (VEC == 4 && c + 3 < p.G) {
      float4 x[kMaxFuse];
#pragma unroll
      for (int m = 0; m < kMaxFuse; ++m)
        if (m < p.n) x[m] = __ldcs(reinterpret_cast<const float4 *>(p.d[m] + q * p.ld + c));
      float4 o;
      float v[kMaxFuse];
#pragma unroll
      for (int m = 0; m < kMaxFuse; ++m) v[m] = (m < p.n) ? x[m].x : 0.f;
      o.x = fuse_one(p, v, wqv, c);
#pragma unroll
      for (int m = 0; m < kMaxFuse; ++m) v[m] = (m < p.n) ? x[m].y : 0.f;
      o.y = fuse_one(p, v, wqv, c + 1);
#pragma unroll
      for (int m = 0; m < kMaxFuse; ++m) v[m] = (m < p.n) ? x[m].z : 0.f;
      o.z = fuse_one(p, v, wqv, c + 2);
#pragma unroll
      for (int m = 0; m < kMaxFuse; ++m) v[m] = (m < p.n) ? x[m].w : 0.f;
      o.w = fuse_one(p, v, wqv, c + 3);
      *reinterpret_cast<float4 *>(p.out + q * p.ld + c) = o;
    } else {
      for (int64_t e = c; e < c + VEC && e < p.G; ++e) {
        float v[kMaxFuse];
#pragma unroll
        for (int m = 0; m < kMaxFuse; ++m) v[m] = (m < p.n) ? __ldg(p.d[m] + q * p.ld + e) : 0.f;
        p.out[q * p.ld + e] = fuse_one(p, v, wqv, e);
      }
    }
  }
}

// Contiguous matrices (ld == G, what numpy / torch hand over): the Q x G elements are one flat
// array, so 128-bit accesses work for any G (15913 is odd); the (row, column) of an element is
// only needed for the weighted form.
__global__ void __launch_bounds__(256) fuse_flat_kernel(FuseParams p) {
  const int64_t total = p.Q * p.G;
  const int64_t nvec = (total + 3) >> 2;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < nvec;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t e0 = i << 2;
    int64_t q = 0, col = e0;
    if (p.weighted) {
      q = e0 / p.G;
      col = e0 - q * p.G;
    }
    float xs[kMaxFuse][4];
    const bool full = e0 + 3 < total;
#pragma unroll
    for (int m = 0; m < kMaxFuse; ++m) {
      if (m < p.n) {
        if (full) {
          const float4 x = __ldcs(reinterpret_cast<const float4 *>(p.d[m] + e0));
          xs[m][0] = x.x; xs[m][1] = x.y; xs[m][2] = x.z; xs[m][3] = x.w;
        } else {
#pragma unroll
          for (int t = 0; t < 4; ++t) xs[m][t] = e0 + t < total ? __ldg(p.d[m] + e0 + t) : 0.f;
        }
      }
    }
    float o[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      float v[kMaxFuse], wqv[kMaxFuse];
#pragma unroll
      for (int m = 0; m < kMaxFuse; ++m) {
        v[m] = (m < p.n) ? xs[m][t] : 0.f;
        wqv[m] = (p.weighted && m < p.n && e0 + t < total) ? __ldg(p.wq[m] + q) : 0.f;
      }
      o[t] = (e0 + t < total) ? fuse_one(p, v, wqv, col) : 0.f;
      if (++col == p.G) { col = 0; ++q; }
    }
    if (full) {
      *reinterpret_cast<float4 *>(p.out + e0) = make_float4(o[0], o[1], o[2], o[3]);
    } else {
#pragma unroll
      for (int t = 0; t < 4; ++t)
        if (e0 + t < total) p.out[e0 + t] = o[t];
    }
  }
}

}  // namespace

// All pointers are DEVICE pointers here (capi.cu stages host operands first).
int launch_fuse(dali_ctx *ctx, const float *const *d, int n, const float *const *wq,
                const float *const *wg, float *out, int64_t Q, int64_t G, int64_t ld) {
  if (n < 1 || n > kMaxFuse) return set_err(ctx, DALI_ERR_INVALID, "fusion supports 1..8 matrices");
  if (Q == 0 || G == 0) return DALI_OK;
  FuseParams p;
  bool aligned = (ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  for (int m = 0; m < kMaxFuse; ++m) {
    p.d[m] = m < n ? d[m] : nullptr;
    p.wq[m] = (wq && m < n) ? wq[m] : nullptr;
    p.wg[m] = (wg && m < n) ? wg[m] : nullptr;
    if (m < n && (reinterpret_cast<uintptr_t>(d[m]) & 15)) aligned = false;
  }
  p.out = out; p.n = n; p.weighted = (wq && wg) ? 1 : 0; p.Q = Q; p.G = G; p.ld = ld;
  bool base_aligned = (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  for (int m = 0; m < n; ++m) base_aligned = base_aligned && (reinterpret_cast<uintptr_t>(d[m]) & 15) == 0;
  if ((ld == G || Q == 1) && base_aligned) {
    KTimer t(ctx, DALI_K_FUSE);
    const int64_t nvec = (Q * G + 3) / 4;
    const int64_t want = (nvec + 255) / 256;
    const int blocks = static_cast<int>(std::min<int64_t>(want, 8ll * ctx->num_sms));
    fuse_flat_kernel<<<blocks, 256, 0, ctx->stream>>>(p);
    DALI_CUDA_OK(ctx, cudaGetLastError());
    return DALI_OK;
  }
  if (Q > 65535 * 1ll) {
    // grid.y limit: process in row bands
    return set_err(ctx, DALI_ERR_UNSUPPORTED, "fusion of more than 65535 rows: call per band");
  }
  const int vec = aligned ? 4 : 1;
  const int64_t nvec = (G + vec - 1) / vec;
  int bx = static_cast<int>((nvec + 255) / 256);
  if (bx > 64) bx = 64;
  dim3 grid(bx, static_cast<unsigned>(Q));
  KTimer t(ctx, DALI_K_FUSE);
  if (aligned)
    fuse_kernel<4><<<grid, 256, 0, ctx->stream>>>(p);
  else
    fuse_kernel<1><<<grid, 256, 0, ctx->stream>>>(p);
  DALI_CUDA_OK(ctx, cudaGetLastError());
  return DALI_OK;
}

}  // namespace dali
