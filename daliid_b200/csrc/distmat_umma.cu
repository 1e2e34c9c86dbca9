// Q x G x D distance contraction on the 5th-generation tensor cores (SURVEY 8a rows a2/a2',
// precisions DALI_PREC_TF32 and DALI_PREC_TF32X3).
//
//   out[i,j] = epilogue( sum_k A[i,k] * B[j,k] )        A = prepared queries, B = gallery
//
// Both operands are K-major fp32 planes written by normalize.cu (zero padded, rounded to
// TF32; the 3xTF32 mode adds the residual plane and issues hi*hi + hi*lo + lo*hi into the
// same TMEM accumulator, which recovers fp32-class accuracy on the tensor pipe).
//
// Structure (one persistent CTA per SM, 192 threads):
//   warp 0    TMA producer: cp.async.bulk.tensor 2D tiles (128B swizzle) into a ring of
//             shared-memory stages, completion on mbarriers
//   warp 1    allocates TMEM (512 columns = two 128x256 fp32 accumulators) and issues
//             tcgen05.mma.cta_group::1.kind::tf32 (M=128, N=256, K=8) from one elected lane
//   warps 2-5 epilogue: tcgen05.ld the finished accumulator (32 lanes x 32 columns at a
//             time), apply the metric, transpose through shared memory and store coalesced
//             rows; overlaps the next tile's MMAs through the second accumulator
//
// replaces  1.0 - torch.mm(q, g.T)   validateModels.py:47, evaluate.py:260-267,291,
//           evaluate_ensembled_models.py:281,300, evaluateCleanATModels.py:109,121,124
#include <cuda.h>

#include "common.cuh"

namespace dali {

namespace {

constexpr int BM = 128;        // UMMA M (TMEM lanes)
constexpr int BN = 256;        // UMMA N (TMEM columns per accumulator)
constexpr int BK = 32;         // floats per stage row = 128 bytes = one swizzle atom
constexpr int UMMA_K = 8;      // tf32: 32 bytes of K per instruction
constexpr int kThreads = 192;
constexpr int A_BYTES = BM * BK * 4;  // 16 KiB
constexpr int B_BYTES = BN * BK * 4;  // 32 KiB
constexpr int EPI_LD = 33;
constexpr int EPI_BYTES = 4 * 32 * EPI_LD * 4;
constexpr uint64_t kWatchdogCycles = 4000000000ull;  // ~2 s: trap instead of hanging the box

template <int NPL>
struct Cfg {
  static constexpr int kStages = (NPL == 1) ? 4 : 2;
  static constexpr int kStageBytes = NPL * (A_BYTES + B_BYTES);
  static constexpr int kSmemBytes = kStages * kStageBytes + EPI_BYTES + 256 + 1024;
};

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const uint64_t t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > kWatchdogCycles) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar,
                                            int32_t x, int32_t y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(map), "r"(x), "r"(y), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
               ::"r"(bar)
               : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
        "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]),
        "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tc_wait_ld() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major operand tile, 128-byte swizzle: rows of 128 B, 8-row groups 1024 B apart.
// start address [0,14) (>>4), LBO [16,30) = 1 (unused for swizzled K-major), SBO [32,46) = 64
// (1024 B), descriptor version [46,48) = 1 (sm_100), layout type [61,64) = 2 (SWIZZLE_128B).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  return static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) |
         (1ull << 46) | (2ull << 61);
}

// kind::tf32, fp32 accumulate, A and B K-major, N=256, M=128 (cute::UMMA::InstrDescriptor).
constexpr uint32_t kInstrDesc = (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(BN >> 3) << 17) |
                                (uint32_t(BM >> 4) << 24);

__device__ __forceinline__ float epilogue(float acc, int metric, float qs, float gs) {
  switch (metric) {
    case DALI_METRIC_COSINE: return 1.0f - acc;
    case DALI_METRIC_SQEUCLIDEAN: return fmaf(-2.0f, acc, qs + gs);
    case DALI_METRIC_EUCLIDEAN: return sqrtf(fmaxf(fmaf(-2.0f, acc, qs + gs), 1e-30f));
    default: return acc;
  }
}

struct UmmaParams {
  int64_t Q, G;
  int num_m_tiles, num_n_tiles, num_kb;
  int32_t a_plane_rows, b_plane_rows;  // row offset of plane 1 inside the tensor maps
  int metric;
  const float *qsq, *gsq;
  float *out;
  int64_t ld;
};

template <int NPL>
__global__ void __launch_bounds__(kThreads, 1)
distmat_umma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const UmmaParams p) {
  using C = Cfg<NPL>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                              ~uintptr_t(1023));
  float *epi = reinterpret_cast<float *>(smem + C::kStages * C::kStageBytes);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + C::kStages * C::kStageBytes + EPI_BYTES);
  // bars: [0,S) full, [S,2S) empty, [2S,2S+2) tmem_full, [2S+2,2S+4) tmem_empty, then tmem ptr
  uint32_t *tmem_ptr_s = reinterpret_cast<uint32_t *>(bars + 2 * C::kStages + 4);
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (C::kStages + s); };
  auto tfull_bar = [&](int s) { return bar0 + 8u * (2 * C::kStages + s); };
  auto tempty_bar = [&](int s) { return bar0 + 8u * (2 * C::kStages + 2 + s); };

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.num_m_tiles * p.num_n_tiles;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 4);  // one arrive per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;"
                 ::"r"(smem_u32(tmem_ptr_s))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int m = t % p.num_m_tiles, n = t / p.num_m_tiles;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sbase = smem_u32(smem + stage * C::kStageBytes);
          mbar_expect_tx(full_bar(stage), C::kStageBytes);
#pragma unroll
          for (int pl = 0; pl < NPL; ++pl) {
            tma_load_2d(sbase + pl * A_BYTES, &tmA, full_bar(stage), kb * BK,
                        pl * p.a_plane_rows + m * BM);
            tma_load_2d(sbase + NPL * A_BYTES + pl * B_BYTES, &tmB, full_bar(stage), kb * BK,
                        pl * p.b_plane_rows + n * BN);
          }
          if (++stage == C::kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
        const int as = it & 1;
        mbar_wait(tempty_bar(as), ((it >> 1) & 1u) ^ 1u);  // epilogue drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(as * BN);
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sbase = smem_u32(smem + stage * C::kStageBytes);
          const uint32_t a0 = sbase, b0 = sbase + NPL * A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint32_t koff = k * UMMA_K * 4;  // bytes inside the 128 B swizzle atom
            const uint32_t first = (kb | k) ? 1u : 0u;
            if (NPL == 1) {
              tc_mma_tf32(tmem_d, make_smem_desc(a0 + koff), make_smem_desc(b0 + koff), kInstrDesc,
                          first);
            } else {
              const uint64_t ahi = make_smem_desc(a0 + koff);
              const uint64_t alo = make_smem_desc(a0 + A_BYTES + koff);
              const uint64_t bhi = make_smem_desc(b0 + koff);
              const uint64_t blo = make_smem_desc(b0 + B_BYTES + koff);
              tc_mma_tf32(tmem_d, alo, bhi, kInstrDesc, first);
              tc_mma_tf32(tmem_d, ahi, blo, kInstrDesc, 1u);
              tc_mma_tf32(tmem_d, ahi, bhi, kInstrDesc, 1u);
            }
          }
          tc_commit(empty_bar(stage));  // frees the stage when these MMAs retire
          if (++stage == C::kStages) { stage = 0; phase ^= 1u; }
        }
        tc_commit(tfull_bar(as));  // accumulator complete
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue warps =====================
    const int quarter = warp & 3;  // TMEM lanes [32*quarter, 32*quarter+32)
    float *stg = epi + (warp - 2) * 32 * EPI_LD;
    int it = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      const int m = t % p.num_m_tiles, n = t / p.num_m_tiles;
      const int as = it & 1;
      mbar_wait(tfull_bar(as), (it >> 1) & 1u);
      tc_fence_after();
      const int64_t row0 = static_cast<int64_t>(m) * BM + quarter * 32;
      const int64_t colt = static_cast<int64_t>(n) * BN;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t v[32];
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                               static_cast<uint32_t>(as * BN + c * 32);
        tc_ld_32x32(taddr, v);
        tc_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; ++j) stg[lane * EPI_LD + j] = __uint_as_float(v[j]);
        __syncwarp();
        const int64_t col = colt + c * 32 + lane;
        const bool col_ok = col < p.G;
        const float gs = (p.gsq && col_ok) ? __ldg(p.gsq + col) : 0.f;
#pragma unroll 4
        for (int rr = 0; rr < 32; ++rr) {
          const int64_t r = row0 + rr;
          if (r < p.Q && col_ok) {
            const float qs = p.qsq ? __ldg(p.qsq + r) : 0.f;
            p.out[r * p.ld + col] = epilogue(stg[rr * EPI_LD + lane], p.metric, qs, gs);
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(as));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base)
                 : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                  const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_map(dali_ctx *ctx, CUtensorMap *map, const float *base, int64_t rows, int64_t Dp,
             int box_rows) {
  if (!ctx->encode_tiled) {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn)
      return set_err(ctx, DALI_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
    ctx->encode_tiled = fn;
  }
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(Dp), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(Dp) * 4};
  cuuint32_t box[2] = {BK, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = reinterpret_cast<EncodeTiledFn>(ctx->encode_tiled)(
      map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), dims, strides, box, estr,
      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_err(ctx, DALI_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string(r));
  return DALI_OK;
}

template <int NPL>
int launch_t(dali_ctx *ctx, const CUtensorMap &tmA, const CUtensorMap &tmB, const UmmaParams &p) {
  using C = Cfg<NPL>;
  static bool attr_set = false;
  if (!attr_set) {
    DALI_CUDA_OK(ctx, cudaFuncSetAttribute(distmat_umma_kernel<NPL>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           C::kSmemBytes));
    attr_set = true;
  }
  const int tiles = p.num_m_tiles * p.num_n_tiles;
  const int grid = tiles < ctx->num_sms ? tiles : ctx->num_sms;
  KTimer t(ctx, DALI_K_DISTMAT);
  distmat_umma_kernel<NPL><<<grid, kThreads, C::kSmemBytes, ctx->stream>>>(tmA, tmB, p);
  DALI_CUDA_OK(ctx, cudaGetLastError());
  return DALI_OK;
}

}  // namespace

// q_planes: [NPL][q_rows_pad][Dp] fp32 (plane 0 = tf32-rounded operand, plane 1 = residual),
// g_planes likewise; rows_pad multiples of 256, Dp multiple of 32.
int launch_distmat_umma(dali_ctx *ctx, const float *q_planes, const float *g_planes, int64_t Q,
                        int64_t G, int64_t Dp, int64_t q_rows_pad, int64_t g_rows_pad, int split3,
                        int metric, const float *qsq, const float *gsq, float *out, int64_t ld) {
  if (Q == 0 || G == 0) return DALI_OK;
  const int npl = split3 ? 2 : 1;
  if (Dp % BK != 0 || q_rows_pad % BM != 0 || g_rows_pad % BN != 0)
    return set_err(ctx, DALI_ERR_INVALID, "umma operands must be padded (rows 128/256, D 32)");
  if (q_rows_pad * npl > INT32_MAX || g_rows_pad * npl > INT32_MAX)
    return set_err(ctx, DALI_ERR_UNSUPPORTED, "operand too tall for one tensor map");
  CUtensorMap tmA, tmB;
  int rc = make_map(ctx, &tmA, q_planes, q_rows_pad * npl, Dp, BM);
  if (rc) return rc;
  rc = make_map(ctx, &tmB, g_planes, g_rows_pad * npl, Dp, BN);
  if (rc) return rc;
  UmmaParams p;
  p.Q = Q; p.G = G;
  p.num_m_tiles = static_cast<int>((Q + BM - 1) / BM);
  p.num_n_tiles = static_cast<int>((G + BN - 1) / BN);
  p.num_kb = static_cast<int>(Dp / BK);
  p.a_plane_rows = static_cast<int32_t>(q_rows_pad);
  p.b_plane_rows = static_cast<int32_t>(g_rows_pad);
  p.metric = metric; p.qsq = qsq; p.gsq = gsq; p.out = out; p.ld = ld;
  if (static_cast<int64_t>(p.num_m_tiles) * p.num_n_tiles > INT32_MAX)
    return set_err(ctx, DALI_ERR_UNSUPPORTED, "too many tiles (chunk the queries)");
  return split3 ? launch_t<2>(ctx, tmA, tmB, p) : launch_t<1>(ctx, tmA, tmB, p);
}

}  // namespace dali
