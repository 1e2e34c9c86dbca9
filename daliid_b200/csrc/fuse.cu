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

// Flat form: the matrices are one array of Q * ld floats each (ld == G: what numpy / torch hand over,
// 128-bit accesses then work for any G -- 15913 is odd; or rows padded to a multiple of four floats:
// what compute_distance_matrix returns on CUDA, the padding columns are fused along and never read).
// N is a compile-time constant (the 8-way predicated unroll over a run-time n cost a third of the
// bandwidth), two float4 per matrix are in flight per thread, and the mean of 2 / 4 / 8 matrices
// multiplies by the exact reciprocal instead of dividing.
template <int N, bool WEIGHTED>
__global__ void __launch_bounds__(256) fuse_flat_kernel(FuseParams p) {
  const int64_t total = p.Q * p.ld;
  const int64_t nvec = (total + 3) >> 2;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  auto one = [&](const float (&v)[N], int64_t q, int64_t col) -> float {
    if (!WEIGHTED) {
      float acc = v[0];
#pragma unroll
      for (int m = 1; m < N; ++m) acc = __fadd_rn(acc, v[m]);
      if (N == 1) return acc;
      if (N == 2 || N == 4 || N == 8) return __fmul_rn(acc, 1.0f / N);  // exact: a power of two
      return div_small_int(acc, static_cast<float>(N), __frcp_rn(static_cast<float>(N)));
    }
    if (col >= p.G) return 0.f;  // padding column
    float num = 0.f, den = 0.f;
#pragma unroll
    for (int m = 0; m < N; ++m) {
      const float a = __ldg(p.wq[m] + q), g = __ldg(p.wg[m] + col);
      float w = fmaxf(a, g);
      if (a != a || g != g) w = NAN;  // torch.maximum propagates NaN; fmaxf does not
      num = m == 0 ? __fmul_rn(w, v[0]) : __fadd_rn(num, __fmul_rn(w, v[m]));
      den = m == 0 ? w : __fadd_rn(den, w);
    }
    return __fdiv_rn(num, den);
  };
  auto body = [&](int64_t i) {
    const int64_t e0 = i << 2;
    if (e0 + 3 < total) {
      float4 x[N];
#pragma unroll
      for (int m = 0; m < N; ++m) x[m] = __ldcs(reinterpret_cast<const float4 *>(p.d[m] + e0));
      int64_t q = 0, col = 0;
      if (WEIGHTED) { q = e0 / p.ld; col = e0 - q * p.ld; }  // ld % 4 == 0 or ld == G: see the launcher
      float v[N];
      float4 o;
#pragma unroll
      for (int m = 0; m < N; ++m) v[m] = x[m].x;
      o.x = one(v, q, col);
      if (WEIGHTED && ++col == p.ld) { col = 0; ++q; }
#pragma unroll
      for (int m = 0; m < N; ++m) v[m] = x[m].y;
      o.y = one(v, q, col);
      if (WEIGHTED && ++col == p.ld) { col = 0; ++q; }
#pragma unroll
      for (int m = 0; m < N; ++m) v[m] = x[m].z;
      o.z = one(v, q, col);
      if (WEIGHTED && ++col == p.ld) { col = 0; ++q; }
#pragma unroll
      for (int m = 0; m < N; ++m) v[m] = x[m].w;
      o.w = one(v, q, col);
      __stcs(reinterpret_cast<float4 *>(p.out + e0), o);
    } else {
      for (int64_t e = e0; e < total; ++e) {
        float v[N];
#pragma unroll
        for (int m = 0; m < N; ++m) v[m] = __ldg(p.d[m] + e);
        const int64_t q = e / p.ld;
        p.out[e] = one(v, q, e - q * p.ld);
      }
    }
  };
  int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  for (; i + stride < nvec; i += 2 * stride) {  // two independent vectors per trip
    body(i);
    body(i + stride);
  }
  if (i < nvec) body(i);
}

template <bool WEIGHTED>
static void launch_flat(int n, int blocks, cudaStream_t st, const FuseParams &p) {
  switch (n) {
    case 1: fuse_flat_kernel<1, WEIGHTED><<<blocks, 256, 0, st>>>(p); break;
    case 2: fuse_flat_kernel<2, WEIGHTED><<<blocks, 256, 0, st>>>(p); break;
    case 3: fuse_flat_kernel<3, WEIGHTED><<<blocks, 256, 0, st>>>(p); break;
    case 4: fuse_flat_kernel<4, WEIGHTED><<<blocks, 256, 0, st>>>(p); break;
    case 5: fuse_flat_kernel<5, WEIGHTED><<<blocks, 256, 0, st>>>(p); break;
    case 6: fuse_flat_kernel<6, WEIGHTED><<<blocks, 256, 0, st>>>(p); break;
    case 7: fuse_flat_kernel<7, WEIGHTED><<<blocks, 256, 0, st>>>(p); break;
    default: fuse_flat_kernel<8, WEIGHTED><<<blocks, 256, 0, st>>>(p); break;
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
  // flat form: contiguous matrices, or rows padded to whole float4 (then a vector never straddles
  // two rows and the padding columns are simply fused along)
  if ((ld == G || Q == 1 || ld % 4 == 0) && base_aligned) {
    if (Q == 1) p.ld = G;
    KTimer t(ctx, DALI_K_FUSE);
    const int64_t nvec = (p.Q * p.ld + 3) / 4;
    const int64_t want = (nvec + 511) / 512;
    const int blocks = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(want, 16ll * ctx->num_sms)));
    if (p.weighted) launch_flat<true>(n, blocks, ctx->stream, p);
    else launch_flat<false>(n, blocks, ctx->stream, p);
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

// ---- exhaustive check of div_small_int against the IEEE division (test hook) -------------------
namespace {
__global__ void __launch_bounds__(256) selftest_div_kernel(float n, unsigned long long *bad) {
  const float rn = __frcp_rn(n);
  unsigned long long mine = 0;
  for (uint64_t b = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; b < (1ull << 32);
       b += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    const float x = __uint_as_float(static_cast<uint32_t>(b));
    const float a = div_small_int(x, n, rn), e = __fdiv_rn(x, n);
    if (__float_as_uint(a) != __float_as_uint(e) && !(a != a && e != e)) ++mine;
  }
  if (mine) atomicAdd(bad, mine);
}
}  // namespace

int launch_selftest_div(dali_ctx *ctx, int n, unsigned long long *bad_dev) {
  selftest_div_kernel<<<ctx->num_sms * 8, 256, 0, ctx->stream>>>(static_cast<float>(n), bad_dev);
  DALI_CUDA_OK(ctx, cudaGetLastError());
  return DALI_OK;
}

}  // namespace dali
