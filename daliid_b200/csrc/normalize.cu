// Row preparation: L2 normalisation fused with the operand layout the
// contraction kernels read: zero-padded rows of d_pad floats, optionally rounded to TF32
// (plane 0) with the rounding residual as a second plane (3xTF32 split).
//
// replaces  x/torch.norm(x, dim=1, keepdim=True)
//   validateModels.py:41-42, evaluate.py:251-258,285-286,
//   evaluate_ensembled_models.py:278-279,297-298, evaluateCleanATModels.py:106-107,115-119,252-254
// No eps, as in the reference: a zero row divides 0/0 and becomes NaN.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace dali {

namespace {

constexpr int kPrepThreads = 256;

__device__ __forceinline__ float block_sum(float v, float *s_red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();  // s_red may still be read from a previous call
  if (l == 0) s_red[w] = v;
  __syncthreads();
  float t = (l < (kPrepThreads / 32)) ? s_red[l] : 0.f;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  return t;  // every thread holds the block total
}

__device__ __forceinline__ float round_tf32(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

struct PrepParams {
  const float *x;
  int64_t n, d, ldx;
  float *plane0;
  float *plane1;  // nullable: residual plane
  int64_t ldo, d_pad, rows_pad;
  int do_normalize;
  int round_mode;  // 0 keep fp32, 1 round plane0 to tf32, 2 fp16 hi/residual of 2^12 * x (no fp32 plane)
  float *norms;    // nullable: ||x|| of the input row
  float *sq;       // nullable: sum of squares of the OUTPUT row (fp32 values before tf32 rounding)
  __nv_bfloat16 *hi16;  // nullable: bf16 copy of plane0 (TF32C correction operand)
  __nv_bfloat16 *lo16;  // nullable: bf16 copy of the residual x - plane0
};

__global__ void __launch_bounds__(kPrepThreads) prep_rows_kernel(PrepParams p) {
  __shared__ float s_red[kPrepThreads / 32];
  const int64_t r = blockIdx.x;
  float *o0 = p.plane0 ? p.plane0 + r * p.ldo : nullptr;
  float *o1 = p.plane1 ? p.plane1 + r * p.ldo : nullptr;
  __nv_bfloat16 *h16 = p.hi16 ? p.hi16 + r * p.ldo : nullptr;
  __nv_bfloat16 *l16 = p.lo16 ? p.lo16 + r * p.ldo : nullptr;
  if (r >= p.n) {  // padding rows: zeros
    for (int64_t c = threadIdx.x; c < p.d_pad; c += kPrepThreads) {
      if (o0) o0[c] = 0.f;
      if (o1) o1[c] = 0.f;
      if (h16) { h16[c] = __float2bfloat16_rn(0.f); l16[c] = __float2bfloat16_rn(0.f); }
    }
    return;
  }
  const float *xr = p.x + r * p.ldx;
  float nrm = 1.f;
  if (p.do_normalize || p.norms) {
    float acc = 0.f;
    for (int64_t c = threadIdx.x; c < p.d; c += kPrepThreads) {
      const float v = __ldg(xr + c);
      acc = fmaf(v, v, acc);
    }
    nrm = sqrtf(block_sum(acc, s_red));
    if (p.norms && threadIdx.x == 0) p.norms[r] = nrm;
  }
  float acc2 = 0.f;
  for (int64_t c = threadIdx.x; c < p.d_pad; c += kPrepThreads) {
    float v = 0.f;
    if (c < p.d) {
      v = __ldg(xr + c);
      if (p.do_normalize) v = v / nrm;  // IEEE division, like the reference's x / norm
    }
    acc2 = fmaf(v, v, acc2);
    if (p.round_mode == 2) {
      // unit rows: |v| <= 1, so 4096 v fits fp16 (max 65504) and a typical 1/sqrt(D) entry and
      // its residual stay in the normal range; the contraction's epilogue multiplies by 2^-24
      const float sv = v * 4096.0f;
      const __half hi = __float2half_rn(sv);
      reinterpret_cast<__half *>(h16)[c] = hi;
      reinterpret_cast<__half *>(l16)[c] = __float2half_rn(sv - __half2float(hi));
    } else if (p.round_mode) {
      const float hi = round_tf32(v);
      o0[c] = hi;
      if (o1) o1[c] = round_tf32(v - hi);
      if (h16) {
        h16[c] = __float2bfloat16_rn(hi);
        l16[c] = __float2bfloat16_rn(v - hi);
      }
    } else {
      o0[c] = v;
    }
  }
  if (p.sq) {
    const float t = block_sum(acc2, s_red);
    if (threadIdx.x == 0) p.sq[r] = t;
  }
}

}  // namespace

int launch_prep(dali_ctx *ctx, const float *x, int64_t n, int64_t d, int64_t ldx, float *plane0,
                float *plane1, int64_t ldo, int64_t d_pad, int64_t rows_pad, int do_normalize,
                int round_mode, float *norms, float *sq, void *hi16, void *lo16) {
  if (rows_pad == 0) return DALI_OK;
  PrepParams p{x, n, d, ldx, plane0, plane1, ldo, d_pad, rows_pad, do_normalize, round_mode, norms, sq,
               static_cast<__nv_bfloat16 *>(hi16), static_cast<__nv_bfloat16 *>(lo16)};
  KTimer t(ctx, DALI_K_NORMALIZE);
  prep_rows_kernel<<<static_cast<unsigned>(rows_pad), kPrepThreads, 0, ctx->stream>>>(p);
  DALI_CUDA_OK(ctx, cudaGetLastError());
  return DALI_OK;
}

}  // namespace dali
