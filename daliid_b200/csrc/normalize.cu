// Row preparation: L2 normalisation fused with the operand layout the
// contraction kernels read: zero-padded rows of d_pad floats, optionally rounded to TF32
// (plane 0) with the rounding residual as a second plane (3xTF32 split).
//
// replaces  x/torch.norm(x, dim=1, keepdim=True)
//   validateModels.py:41-42, evaluate.py:251-258,285-286,
//   evaluate_ensembled_models.py:278-279,297-298, evaluateCleanATModels.py:106-107,115-119,252-254
// No eps, as in the reference: a zero row divides 0/0 and becomes NaN.
#include <cstdlib>
#include <cstring>

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

// round_mode 2 with a fixed-point hi plane: hi = nearest multiple of `grid` (then fp16, which is
// exact below 2048 * grid and coarser -- still a multiple -- above).  Products hi * hi are then
// multiples of grid^2, which the tensor core adds into its fp32 accumulator without dropping low
// bits at the alignment step (see distmat_umma2.cu, f16x3_schedule).
__device__ __forceinline__ float hi_quant(float s, float grid_inv, float grid) {
  return grid_inv != 0.f ? rintf(s * grid_inv) * grid : s;
}

struct PrepParams {
  const float *x;
  int64_t n, d, ldx;
  float *plane0;
  float *plane1;  // nullable: residual plane
  int64_t ldo, d_pad, rows_pad;
  int do_normalize;
  int round_mode;  // 0 keep fp32, 1 round plane0 to tf32, 2 fp16 hi/residual of 2^12 * x (no fp32 plane)
  float hi_grid_inv, hi_grid;  // round_mode 2: hi is first rounded to a multiple of hi_grid (0: off)
  float *norms;    // nullable: ||x|| of the input row
  float *sq;       // nullable: sum of squares of the OUTPUT row (fp32 values before tf32 rounding)
  __nv_bfloat16 *hi16;  // nullable: bf16 copy of plane0 (TF32C correction operand)
  __nv_bfloat16 *lo16;  // nullable: bf16 copy of the residual x - plane0
  const int32_t *perm;  // nullable: output row r comes from input row perm[r]
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
  const float *xr = p.x + (p.perm ? static_cast<int64_t>(__ldg(p.perm + r)) : r) * p.ldx;
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
      const __half hi = __float2half_rn(hi_quant(sv, p.hi_grid_inv, p.hi_grid));
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

// Single-pass variant for rows of at most 4096 elements with 16-byte aligned rows (every
// BASELINE shape): the row is read once with 128-bit loads and kept in registers between the
// norm reduction and the plane writes; 16-bit planes are written 8 bytes at a time.
constexpr int kVecCache = 4;  // float4 per thread: 256 threads * 4 * 4 = 4096 elements

__device__ __forceinline__ uint2 pack_bf16x4(float a, float b, float c, float d) {
  const __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
  return make_uint2(*reinterpret_cast<const uint32_t *>(&lo), *reinterpret_cast<const uint32_t *>(&hi));
}
__device__ __forceinline__ uint2 pack_f16x4(__half a, __half b, __half c, __half d) {
  const __half2 lo = __halves2half2(a, b), hi = __halves2half2(c, d);
  return make_uint2(*reinterpret_cast<const uint32_t *>(&lo), *reinterpret_cast<const uint32_t *>(&hi));
}

__global__ void __launch_bounds__(kPrepThreads) prep_rows_vec_kernel(PrepParams p) {
  __shared__ float s_red[kPrepThreads / 32];
  const int64_t r = blockIdx.x;
  float4 *o0 = p.plane0 ? reinterpret_cast<float4 *>(p.plane0 + r * p.ldo) : nullptr;
  float4 *o1 = p.plane1 ? reinterpret_cast<float4 *>(p.plane1 + r * p.ldo) : nullptr;
  uint2 *h16 = p.hi16 ? reinterpret_cast<uint2 *>(p.hi16 + r * p.ldo) : nullptr;
  uint2 *l16 = p.lo16 ? reinterpret_cast<uint2 *>(p.lo16 + r * p.ldo) : nullptr;
  const int nv_pad = static_cast<int>(p.d_pad >> 2);
  if (r >= p.n) {  // padding rows: zeros
    for (int c = threadIdx.x; c < nv_pad; c += kPrepThreads) {
      if (o0) o0[c] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (o1) o1[c] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (h16) { h16[c] = make_uint2(0u, 0u); l16[c] = make_uint2(0u, 0u); }
    }
    return;
  }
  const float4 *xr = reinterpret_cast<const float4 *>(p.x + (p.perm ? static_cast<int64_t>(__ldg(p.perm + r)) : r) * p.ldx);
  const int nv = static_cast<int>(p.d >> 2);
  float4 cache[kVecCache];
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < kVecCache; ++i) {
    const int c = threadIdx.x + i * kPrepThreads;
    cache[i] = c < nv ? __ldg(xr + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    acc = fmaf(cache[i].x, cache[i].x, acc);
    acc = fmaf(cache[i].y, cache[i].y, acc);
    acc = fmaf(cache[i].z, cache[i].z, acc);
    acc = fmaf(cache[i].w, cache[i].w, acc);
  }
  float nrm = 1.f;
  if (p.do_normalize || p.norms) {
    nrm = sqrtf(block_sum(acc, s_red));
    if (p.norms && threadIdx.x == 0) p.norms[r] = nrm;
  }
  float acc2 = 0.f;
#pragma unroll
  for (int i = 0; i < kVecCache; ++i) {
    const int c = threadIdx.x + i * kPrepThreads;
    if (c >= nv_pad) break;
    float4 v = cache[i];  // zeros beyond d
    if (p.do_normalize && c < nv) {  // IEEE division, like the reference's x / norm
      v.x = v.x / nrm; v.y = v.y / nrm; v.z = v.z / nrm; v.w = v.w / nrm;
    }
    acc2 = fmaf(v.x, v.x, acc2); acc2 = fmaf(v.y, v.y, acc2);
    acc2 = fmaf(v.z, v.z, acc2); acc2 = fmaf(v.w, v.w, acc2);
    if (p.round_mode == 2) {
      const float sx = v.x * 4096.0f, sy = v.y * 4096.0f, sz = v.z * 4096.0f, sw = v.w * 4096.0f;
      const float gi = p.hi_grid_inv, gg = p.hi_grid;
      const __half hx = __float2half_rn(hi_quant(sx, gi, gg)), hy = __float2half_rn(hi_quant(sy, gi, gg)),
                   hz = __float2half_rn(hi_quant(sz, gi, gg)), hw = __float2half_rn(hi_quant(sw, gi, gg));
      h16[c] = pack_f16x4(hx, hy, hz, hw);
      l16[c] = pack_f16x4(__float2half_rn(sx - __half2float(hx)), __float2half_rn(sy - __half2float(hy)),
                          __float2half_rn(sz - __half2float(hz)), __float2half_rn(sw - __half2float(hw)));
    } else if (p.round_mode) {
      const float4 hi = make_float4(round_tf32(v.x), round_tf32(v.y), round_tf32(v.z), round_tf32(v.w));
      o0[c] = hi;
      if (o1) o1[c] = make_float4(round_tf32(v.x - hi.x), round_tf32(v.y - hi.y),
                                  round_tf32(v.z - hi.z), round_tf32(v.w - hi.w));
      if (h16) {
        h16[c] = pack_bf16x4(hi.x, hi.y, hi.z, hi.w);
        l16[c] = pack_bf16x4(v.x - hi.x, v.y - hi.y, v.z - hi.z, v.w - hi.w);
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

// The default operand format (round_mode 2: fp16 hi / residual planes of 2^12 * x_hat) gets its own
// lean kernel: the generic one above spends ~160 instructions per float4 (four IEEE divisions, run-time
// mode branches) and was issue bound (ncu r01f: 73 % issue active at 4.7 TB/s).  Here a row is scaled
// by ONE factor 4096 / ||x|| (the planes are internal: a last-ulp difference to x / ||x|| changes a
// distance by < 1e-8), conversions are packed, and NVEC is a compile-time constant.
template <int NVEC>
__device__ __forceinline__ void prep_f16_body(const PrepParams &p, const int64_t r, float *s_red) {
  uint2 *h16 = reinterpret_cast<uint2 *>(p.hi16 + r * p.ldo);
  uint2 *l16 = reinterpret_cast<uint2 *>(p.lo16 + r * p.ldo);
  const int nv_pad = static_cast<int>(p.d_pad >> 2);
  if (r >= p.n) {  // padding rows: zeros
#pragma unroll
    for (int i = 0; i < NVEC; ++i) {
      const int c = threadIdx.x + i * kPrepThreads;
      if (c < nv_pad) { h16[c] = make_uint2(0u, 0u); l16[c] = make_uint2(0u, 0u); }
    }
    return;
  }
  const float4 *xr = reinterpret_cast<const float4 *>(p.x + (p.perm ? static_cast<int64_t>(__ldg(p.perm + r)) : r) * p.ldx);
  const int nv = static_cast<int>(p.d >> 2);
  float4 cache[NVEC];
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < NVEC; ++i) {
    const int c = threadIdx.x + i * kPrepThreads;
    cache[i] = c < nv ? __ldg(xr + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    acc = fmaf(cache[i].x, cache[i].x, acc);
    acc = fmaf(cache[i].y, cache[i].y, acc);
    acc = fmaf(cache[i].z, cache[i].z, acc);
    acc = fmaf(cache[i].w, cache[i].w, acc);
  }
  float scale = 4096.0f;
  if (p.do_normalize || p.norms) {
    const float nrm = sqrtf(block_sum(acc, s_red));
    if (p.norms && threadIdx.x == 0) p.norms[r] = nrm;
    if (p.do_normalize) scale = 4096.0f / nrm;  // zero row: inf, 0 * inf = NaN like the reference's 0 / 0
  }
  float acc2 = 0.f;
#pragma unroll
  for (int i = 0; i < NVEC; ++i) {
    const int c = threadIdx.x + i * kPrepThreads;
    if (c < nv_pad) {
      const float4 v = cache[i];  // zeros beyond d (0 * inf: only in a NaN row anyway)
      const float sx = c < nv ? v.x * scale : 0.f, sy = c < nv ? v.y * scale : 0.f;
      const float sz = c < nv ? v.z * scale : 0.f, sw = c < nv ? v.w * scale : 0.f;
      const float gi = p.hi_grid_inv, gg = p.hi_grid;
      const __half2 h01 = __floats2half2_rn(hi_quant(sx, gi, gg), hi_quant(sy, gi, gg)),
                    h23 = __floats2half2_rn(hi_quant(sz, gi, gg), hi_quant(sw, gi, gg));
      const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
      const __half2 l01 = __floats2half2_rn(sx - f01.x, sy - f01.y), l23 = __floats2half2_rn(sz - f23.x, sw - f23.y);
      h16[c] = make_uint2(*reinterpret_cast<const uint32_t *>(&h01), *reinterpret_cast<const uint32_t *>(&h23));
      l16[c] = make_uint2(*reinterpret_cast<const uint32_t *>(&l01), *reinterpret_cast<const uint32_t *>(&l23));
      if (p.sq) {
        const float k = 1.0f / 4096.0f;
        acc2 = fmaf(sx * k, sx * k, acc2); acc2 = fmaf(sy * k, sy * k, acc2);
        acc2 = fmaf(sz * k, sz * k, acc2); acc2 = fmaf(sw * k, sw * k, acc2);
      }
    }
  }
  if (p.sq) {
    const float t = block_sum(acc2, s_red);
    if (threadIdx.x == 0) p.sq[r] = t;
  }
}

template <int NVEC>
__global__ void __launch_bounds__(kPrepThreads) prep_f16_kernel(PrepParams p) {
  __shared__ float s_red[kPrepThreads / 32];
  prep_f16_body<NVEC>(p, blockIdx.x, s_red);
}
// two operands (queries and gallery of one evaluation) in ONE launch: the first `blocks_a` CTAs take
// `pa`, the others `pb` -- saves a launch gap and the ramp of the short query launch
template <int NVEC>
__global__ void __launch_bounds__(kPrepThreads) prep_f16_pair_kernel(PrepParams pa, PrepParams pb, unsigned blocks_a) {
  __shared__ float s_red[kPrepThreads / 32];
  if (blockIdx.x < blocks_a) prep_f16_body<NVEC>(pa, blockIdx.x, s_red);
  else prep_f16_body<NVEC>(pb, blockIdx.x - blocks_a, s_red);
}

// Rows of at most 1024 elements (D = 512 / 768: most BASELINE shapes): one WARP per row, eight rows
// per CTA, reductions by shuffle only.  A 256-thread CTA per 3 KB row with two block barriers ran at
// 1.4 TB/s (0.085 ms for queries + gallery at the Market/ViT shape).
template <int NV>
__device__ __forceinline__ void prep_f16_warp_body(const PrepParams &p, const unsigned blk) {
  const int lane = threadIdx.x & 31;
  const int64_t r = static_cast<int64_t>(blk) * (kPrepThreads / 32) + (threadIdx.x >> 5);
  if (r >= p.rows_pad) return;
  uint2 *h16 = reinterpret_cast<uint2 *>(p.hi16 + r * p.ldo);
  uint2 *l16 = reinterpret_cast<uint2 *>(p.lo16 + r * p.ldo);
  const int nv_pad = static_cast<int>(p.d_pad >> 2);
  if (r >= p.n) {  // padding rows: zeros
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + i * 32;
      if (c < nv_pad) { h16[c] = make_uint2(0u, 0u); l16[c] = make_uint2(0u, 0u); }
    }
    return;
  }
  const float4 *xr = reinterpret_cast<const float4 *>(p.x + (p.perm ? static_cast<int64_t>(__ldg(p.perm + r)) : r) * p.ldx);
  const int nv = static_cast<int>(p.d >> 2);
  float4 cache[NV];
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + i * 32;
    cache[i] = c < nv ? __ldg(xr + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    acc = fmaf(cache[i].x, cache[i].x, acc);
    acc = fmaf(cache[i].y, cache[i].y, acc);
    acc = fmaf(cache[i].z, cache[i].z, acc);
    acc = fmaf(cache[i].w, cache[i].w, acc);
  }
  float scale = 4096.0f;
  if (p.do_normalize || p.norms) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    const float nrm = sqrtf(acc);
    if (p.norms && lane == 0) p.norms[r] = nrm;
    if (p.do_normalize) scale = 4096.0f / nrm;
  }
  float acc2 = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = lane + i * 32;
    if (c < nv_pad) {
      const float4 v = cache[i];
      const float sx = c < nv ? v.x * scale : 0.f, sy = c < nv ? v.y * scale : 0.f;
      const float sz = c < nv ? v.z * scale : 0.f, sw = c < nv ? v.w * scale : 0.f;
      const float gi = p.hi_grid_inv, gg = p.hi_grid;
      const __half2 h01 = __floats2half2_rn(hi_quant(sx, gi, gg), hi_quant(sy, gi, gg)),
                    h23 = __floats2half2_rn(hi_quant(sz, gi, gg), hi_quant(sw, gi, gg));
      const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
      const __half2 l01 = __floats2half2_rn(sx - f01.x, sy - f01.y), l23 = __floats2half2_rn(sz - f23.x, sw - f23.y);
      h16[c] = make_uint2(*reinterpret_cast<const uint32_t *>(&h01), *reinterpret_cast<const uint32_t *>(&h23));
      l16[c] = make_uint2(*reinterpret_cast<const uint32_t *>(&l01), *reinterpret_cast<const uint32_t *>(&l23));
      if (p.sq) {
        const float k = 1.0f / 4096.0f;
        acc2 = fmaf(sx * k, sx * k, acc2); acc2 = fmaf(sy * k, sy * k, acc2);
        acc2 = fmaf(sz * k, sz * k, acc2); acc2 = fmaf(sw * k, sw * k, acc2);
      }
    }
  }
  if (p.sq) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc2 += __shfl_xor_sync(0xffffffffu, acc2, o);
    if (lane == 0) p.sq[r] = acc2;
  }
}

template <int NV>
__global__ void __launch_bounds__(kPrepThreads) prep_f16_warp_kernel(PrepParams p) {
  prep_f16_warp_body<NV>(p, blockIdx.x);
}
template <int NV>
__global__ void __launch_bounds__(kPrepThreads) prep_f16_warp_pair_kernel(PrepParams pa, PrepParams pb, unsigned blocks_a) {
  if (blockIdx.x < blocks_a) prep_f16_warp_body<NV>(pa, blockIdx.x);
  else prep_f16_warp_body<NV>(pb, blockIdx.x - blocks_a);
}

static_assert(sizeof(PrepParams) <= sizeof(dali_ctx::prep_params), "dali_ctx::prep_params holds a PrepParams");

// kind = 10 + float4 per lane (warp per row: 2 / 4 / 6 / 8) or 20 + float4 per thread (CTA per row: 1 .. 4)
int launch_f16_kind(dali_ctx *ctx, int kind, const PrepParams &p, unsigned blocks, const PrepParams *pb, unsigned blocks_b) {
  KTimer t(ctx, DALI_K_NORMALIZE);
  const unsigned total = blocks + (pb ? blocks_b : 0u);
#define DALI_PREP_CASE(K, SINGLE, PAIR)                                                       \
  case K:                                                                                     \
    if (pb) PAIR<<<total, kPrepThreads, 0, ctx->stream>>>(p, *pb, blocks);                    \
    else SINGLE<<<total, kPrepThreads, 0, ctx->stream>>>(p);                                  \
    break
  switch (kind) {
    DALI_PREP_CASE(12, prep_f16_warp_kernel<2>, prep_f16_warp_pair_kernel<2>);
    DALI_PREP_CASE(14, prep_f16_warp_kernel<4>, prep_f16_warp_pair_kernel<4>);
    DALI_PREP_CASE(16, prep_f16_warp_kernel<6>, prep_f16_warp_pair_kernel<6>);
    DALI_PREP_CASE(18, prep_f16_warp_kernel<8>, prep_f16_warp_pair_kernel<8>);
    DALI_PREP_CASE(21, prep_f16_kernel<1>, prep_f16_pair_kernel<1>);
    DALI_PREP_CASE(22, prep_f16_kernel<2>, prep_f16_pair_kernel<2>);
    DALI_PREP_CASE(23, prep_f16_kernel<3>, prep_f16_pair_kernel<3>);
    DALI_PREP_CASE(24, prep_f16_kernel<4>, prep_f16_pair_kernel<4>);
    default: return set_err(ctx, DALI_ERR_INVALID, "operand preparation: unknown kernel kind");
  }
#undef DALI_PREP_CASE
  DALI_CUDA_OK(ctx, cudaGetLastError());
  return DALI_OK;
}

int flush_pending_prep(dali_ctx *ctx) {
  if (!ctx->prep_pending) return DALI_OK;
  ctx->prep_pending = false;
  PrepParams q;
  std::memcpy(&q, ctx->prep_params, sizeof(PrepParams));
  return launch_f16_kind(ctx, ctx->prep_kind, q, ctx->prep_blocks, nullptr, 0);
}

}  // namespace

void prep_defer_begin(dali_ctx *ctx) { ctx->prep_defer = true; }
int prep_defer_end(dali_ctx *ctx) {
  ctx->prep_defer = false;
  return flush_pending_prep(ctx);
}

// Grid of the fixed-point hi plane (0 = plain fp16 rounding); DALI_F16X3_GRID overrides (probes).
float f16x3_hi_grid(int64_t) {
  static const char *env = getenv("DALI_F16X3_GRID");
  if (env) return static_cast<float>(atof(env));
  return 0.5f;
}

int launch_prep(dali_ctx *ctx, const float *x, int64_t n, int64_t d, int64_t ldx, float *plane0,
                float *plane1, int64_t ldo, int64_t d_pad, int64_t rows_pad, int do_normalize,
                int round_mode, float *norms, float *sq, void *hi16, void *lo16, const int32_t *perm) {
  if (rows_pad == 0) return DALI_OK;
  // round_mode 3 = round_mode 2 with the fixed-point hi plane of the three-pass fp16 arithmetic
  // (the single-pass F16 mode keeps plain fp16 rounding: it has no residual to absorb a coarser hi)
  const float hgrid = round_mode == 3 ? f16x3_hi_grid(d_pad) : 0.f;
  if (round_mode == 3) round_mode = 2;
  PrepParams p{x, n, d, ldx, plane0, plane1, ldo, d_pad, rows_pad, do_normalize, round_mode,
               hgrid > 0.f ? 1.0f / hgrid : 0.f, hgrid, norms, sq,
               static_cast<__nv_bfloat16 *>(hi16), static_cast<__nv_bfloat16 *>(lo16), perm};
  const bool vec = d % 4 == 0 && d_pad <= 4 * kPrepThreads * kVecCache && ldx % 4 == 0 && ldo % 4 == 0 &&
                   (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(plane0) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(plane1) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(hi16) & 7) == 0 && (reinterpret_cast<uintptr_t>(lo16) & 7) == 0;
  static const char *env_generic = getenv("DALI_PREP_GENERIC");  // cross-check: the generic kernel
  const unsigned grid = static_cast<unsigned>(rows_pad);
  if (vec && round_mode == 2 && hi16 && lo16 && !(env_generic && atoi(env_generic))) {
    const int nvec = static_cast<int>((d_pad / 4 + kPrepThreads - 1) / kPrepThreads);
    const int nvw = static_cast<int>((d_pad / 4 + 31) / 32);  // float4 per lane with one warp per row
    const unsigned gridw = static_cast<unsigned>((rows_pad + kPrepThreads / 32 - 1) / (kPrepThreads / 32));
    static const char *env_warp = getenv("DALI_PREP_WARP");  // 0: one CTA per row also for short rows
    const bool warp_rows = nvw <= 8 && !(env_warp && atoi(env_warp) == 0);
    const int kind = warp_rows ? 10 + (nvw <= 2 ? 2 : nvw <= 4 ? 4 : nvw <= 6 ? 6 : 8) : 20 + std::min(std::max(nvec, 1), 4);
    const unsigned blocks = warp_rows ? gridw : grid;
    if (ctx->prep_pending && ctx->prep_kind == kind) {  // the held-back operand and this one: one launch
      ctx->prep_pending = false;
      PrepParams first;
      std::memcpy(&first, ctx->prep_params, sizeof(PrepParams));
      return launch_f16_kind(ctx, kind, first, ctx->prep_blocks, &p, blocks);
    }
    if (int rc = flush_pending_prep(ctx)) return rc;
    if (ctx->prep_defer) {
      std::memcpy(ctx->prep_params, &p, sizeof(PrepParams));
      ctx->prep_kind = kind;
      ctx->prep_blocks = blocks;
      ctx->prep_pending = true;
      return DALI_OK;
    }
    return launch_f16_kind(ctx, kind, p, blocks, nullptr, 0);
  }
  if (int rc = flush_pending_prep(ctx)) return rc;  // (keeps the stream order of the preparations)
  KTimer t(ctx, DALI_K_NORMALIZE);
  if (vec)
    prep_rows_vec_kernel<<<grid, kPrepThreads, 0, ctx->stream>>>(p);
  else
    prep_rows_kernel<<<static_cast<unsigned>(rows_pad), kPrepThreads, 0, ctx->stream>>>(p);
  DALI_CUDA_OK(ctx, cudaGetLastError());
  return DALI_OK;
}

}  // namespace dali
