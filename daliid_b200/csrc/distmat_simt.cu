// Exact-fp32 distance contraction on the FP32 pipe (SURVEY 8a rows a2/a2', precision
// DALI_PREC_FP32): every output element is ONE fmaf chain over k = 0..D-1 in ascending
// order, independent of the tile it lands in -- so a gallery slab computed on another
// GPU produces bit-identical distances.  This is the verification path for
// the tcgen05 kernels in distmat_umma.cu.
//
// replaces  1.0 - torch.mm(q, g.T)              validateModels.py:47, evaluate.py:291
//           compute_distance_matrix(euclidean)  commented validateModels.py:44
//           torch.cdist(p=2)                    commented validateModels.py:45
//           torch.mm(val, centers.T)            validateModels.py:179
#include "common.cuh"

namespace dali {

namespace {

constexpr int BM = 128, BN = 128, BK = 16;
constexpr int kThreads = 256;
constexpr int PAD = 4;

__device__ __forceinline__ float epilogue(float acc, int metric, float qs, float gs) {
  switch (metric) {
    case DALI_METRIC_COSINE: return 1.0f - acc;
    case DALI_METRIC_SQEUCLIDEAN: return fmaf(-2.0f, acc, qs + gs);
    case DALI_METRIC_EUCLIDEAN: return sqrtf(fmaxf(fmaf(-2.0f, acc, qs + gs), 1e-30f));
    default: return acc;
  }
}

// A: [rows >= gridDim.y*BM, lda] zero padded, B likewise; K = Dp multiple of BK.
__global__ void __launch_bounds__(kThreads, 2)
distmat_simt_kernel(const float *__restrict__ A, const float *__restrict__ B, int64_t lda,
                    int64_t ldb, int64_t Q, int64_t G, int64_t Dp, int metric,
                    const float *__restrict__ qsq, const float *__restrict__ gsq,
                    float *__restrict__ out, int64_t ld) {
  __shared__ __align__(16) float As[2][BK][BM + PAD];
  __shared__ __align__(16) float Bs[2][BK][BN + PAD];
  const int tid = threadIdx.x;
  const int64_t m0 = static_cast<int64_t>(blockIdx.y) * BM;
  const int64_t n0 = static_cast<int64_t>(blockIdx.x) * BN;
  // loader mapping: 4 threads cover 16 k of one row; 64 rows per pass, 2 passes
  const int lrow = tid >> 2, lk = (tid & 3) * 4;
  const float *Ap = A + (m0 + lrow) * lda + lk;
  const float *Bp = B + (n0 + lrow) * ldb + lk;
  // compute mapping: 16 x 16 threads, each 2x2 blocks of 4x4
  const int ty = tid >> 4, tx = tid & 15;

  // accumulators as fp32 pairs: the inner product runs on packed FFMA2 (sm_100 fma.rn.f32x2),
  // two IEEE fmaf per instruction -- half the issue slots and operand reads of scalar FFMA, the
  // same bits (each half is an ordinary round-to-nearest fma in ascending k)
  float2 acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = make_float2(0.f, 0.f);

  float4 ra[2], rb[2];
  auto gload = [&](int64_t k0) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      ra[h] = __ldg(reinterpret_cast<const float4 *>(Ap + h * 64 * lda + k0));
      rb[h] = __ldg(reinterpret_cast<const float4 *>(Bp + h * 64 * ldb + k0));
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = lrow + h * 64;
      As[buf][lk + 0][r] = ra[h].x; As[buf][lk + 1][r] = ra[h].y;
      As[buf][lk + 2][r] = ra[h].z; As[buf][lk + 3][r] = ra[h].w;
      Bs[buf][lk + 0][r] = rb[h].x; Bs[buf][lk + 1][r] = rb[h].y;
      Bs[buf][lk + 2][r] = rb[h].z; Bs[buf][lk + 3][r] = rb[h].w;
    }
  };

  const int nk = static_cast<int>(Dp / BK);
  gload(0);
  sstore(0);
  __syncthreads();
  for (int kb = 0; kb < nk; ++kb) {
    const int buf = kb & 1;
    if (kb + 1 < nk) gload(static_cast<int64_t>(kb + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4 *>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4 *>(&As[buf][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4 *>(&Bs[buf][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4 *>(&Bs[buf][k][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float2 b[4] = {make_float2(b0.x, b0.y), make_float2(b0.z, b0.w), make_float2(b1.x, b1.y),
                           make_float2(b1.z, b1.w)};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float2 ai = make_float2(a[i], a[i]);
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = __ffma2_rn(ai, b[j], acc[i][j]);
      }
    }
    if (kb + 1 < nk) {
      sstore(buf ^ 1);
      __syncthreads();
    }
  }

  // epilogue + store (rows ty*4+i and 64+ty*4+i; cols tx*4+j and 64+tx*4+j)
  const bool vec_ok = ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t r = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (r >= Q) continue;
    const float qs = qsq ? __ldg(qsq + r) : 0.f;
#pragma unroll
    for (int jb = 0; jb < 2; ++jb) {
      const int64_t c = n0 + (jb ? 64 : 0) + tx * 4;
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float gs = (gsq && c + j < G) ? __ldg(gsq + c + j) : 0.f;
        const float2 pr = acc[i][jb * 2 + (j >> 1)];
        v[j] = epilogue((j & 1) ? pr.y : pr.x, metric, qs, gs);
      }
      float *dst = out + r * ld + c;
      if (vec_ok && c + 3 < G) {
        *reinterpret_cast<float4 *>(dst) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (c + j < G) dst[j] = v[j];
      }
    }
  }
}

}  // namespace

// qn/gn: zero-padded planes ([rows_pad, ld*], rows_pad multiple of 128, Dp multiple of 32)
int launch_distmat_simt(dali_ctx *ctx, const float *qn, const float *gn, int64_t Q, int64_t G,
                        int64_t Dp, int64_t ldq, int64_t ldg, int metric, const float *qsq,
                        const float *gsq, float *out, int64_t ld) {
  if (Q == 0 || G == 0) return DALI_OK;
  dim3 grid(static_cast<unsigned>((G + BN - 1) / BN), static_cast<unsigned>((Q + BM - 1) / BM));
  if (grid.y > 65535) return set_err(ctx, DALI_ERR_UNSUPPORTED, "Q too large for the SIMT path (chunk it)");
  KTimer t(ctx, DALI_K_DISTMAT);
  distmat_simt_kernel<<<grid, kThreads, 0, ctx->stream>>>(qn, gn, ldq, ldg, Q, G, Dp, metric, qsq,
                                                         gsq, out, ld);
  DALI_CUDA_OK(ctx, cudaGetLastError());
  return DALI_OK;
}

}  // namespace dali
