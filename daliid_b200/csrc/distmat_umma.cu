// Q x G x D distance contraction on the 5th-generation tensor cores, one CTA per tile (SURVEY 8a
// rows a2/a2', precisions DALI_PREC_TF32, DALI_PREC_TF32X3 and DALI_PREC_TF32C).  The first
// tcgen05 kernel of this repo; the default path is now the CTA-pair kernel of distmat_umma2.cu and
// this one is kept as its bit-identical cross-check (DALI_UMMA_2CTA=0, see
// tests/test_gpu_distmat.py::test_two_cta_kernel_equals_one_cta_kernel).
//
//   out[i,j] = epilogue( sum_k A[i,k] * B[j,k] )        A = prepared queries, B = gallery
//
// Operands are K-major planes written by normalize.cu (zero padded):
//   hi    fp32 rounded to TF32                              (all modes)
//   lo    fp32 residual x - hi, rounded to TF32             (TF32X3: hi*hi + hi*lo + lo*hi)
//   hi16, lo16  bf16 copies of hi and of the residual       (TF32C : hi*hi on kind::tf32 plus the
//               two correction products on kind::f16/bf16 at twice the rate; error <= 2^-18/product)
// All MMAs of one output tile accumulate into the same fp32 TMEM accumulator.
//
// Structure (one persistent CTA per SM, 192 threads):
//   warp 0    TMA producer: cp.async.bulk.tensor 2D tiles (128B / 64B swizzle) into a ring of four
//             48 KiB shared-memory slots, completion on mbarriers
//   warp 1    allocates TMEM (512 columns = two 128x256 fp32 accumulators) and issues
//             tcgen05.mma.cta_group::1 (M=128, N=256; K=8 tf32 / K=16 bf16) from one elected lane
//   warps 2-5 epilogue: tcgen05.ld the finished accumulator (32 lanes x 32 columns at a
//             time), apply the metric, transpose through shared memory and store coalesced
//             rows; overlaps the next tile's MMAs through the second accumulator
//
// replaces  1.0 - torch.mm(q, g.T)   validateModels.py:47, evaluate.py:260-267,291,
//           evaluate_ensembled_models.py:281,300, evaluateCleanATModels.py:109,121,124
#include <cstdlib>

#include "umma_common.cuh"

namespace dali {

namespace {

using namespace umma;
constexpr int kThreads = 192;

// MH = number of 128-row UMMA halves per CTA tile.
//   MH = 1: 128 x 256 tile, 4 slots of 48 KiB, two accumulators (epilogue overlaps the next tile)
//   MH = 2: 256 x 256 tile, 3 slots of 64 KiB, both TMEM accumulators belong to one tile: a third
//           less operand traffic per FLOP on the SM->L2 request path, which ncu showed saturated
//           (l1tex2xbar 81 %) with the 128-row tile; the epilogue (~3 % of a tile) is not hidden.
template <int MH>
struct Cfg {
  static constexpr int BM = UM * MH;
  static constexpr int A_BYTES = BM * BK * 4;    // fp32 tile (rows of 128 B, SWIZZLE_128B)
  static constexpr int B_BYTES = BN * BK * 4;    // 32 KiB
  static constexpr int A16_BYTES = BM * BK * 2;  // bf16 tile (rows of 64 B, SWIZZLE_64B)
  static constexpr int B16_BYTES = BN * BK * 2;  // 16 KiB
  static constexpr int kSlotBytes = A_BYTES + B_BYTES;  // == 2*A16 + 2*B16
  static constexpr int kSlots = MH == 1 ? 4 : 3;
  static constexpr int kAcc = MH == 1 ? 2 : 1;   // accumulator stages
  static constexpr int kSmemBytes = kSlots * kSlotBytes + EPI_BYTES + 256 + 1024;
};



struct UmmaParams {
  int64_t Q, G;
  int num_m_tiles, num_n_tiles, num_kb;
  int32_t a_plane_rows, b_plane_rows;  // row offset of plane 1 inside the tensor maps
  int32_t b_row0;                      // first gallery row of this slab inside the B planes
  int metric;
  const float *qsq, *gsq;
  float *out;
  int64_t ld;
  uint64_t hint_a, hint_b;  // L2 eviction policies of the operand loads (0 = none)
  int stream_out;           // 1: evict-first stores of the distance tile
};

// Slots per k-block: kTf32 1 (A_hi|B_hi); kTf32x3 2 (A_hi|B_hi, A_lo|B_lo);
// kTf32c 2 (A_hi|B_hi fp32, then A_hi16|A_lo16|B_hi16|B_lo16 bf16).
template <int MODE, int MH>
__global__ void __launch_bounds__(kThreads, 1)
distmat_umma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmA16,
                    const __grid_constant__ CUtensorMap tmB16, const UmmaParams p) {
  using C = Cfg<MH>;
  constexpr int kSlots = C::kSlots, kSlotBytes = C::kSlotBytes, kAcc = C::kAcc, BM = C::BM;
  constexpr int A_BYTES = C::A_BYTES, A16_BYTES = C::A16_BYTES, B16_BYTES = C::B16_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                              ~uintptr_t(1023));
  float *epi = reinterpret_cast<float *>(smem + kSlots * kSlotBytes);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + kSlots * kSlotBytes + EPI_BYTES);
  // bars: [0,S) full, [S,2S) empty, [2S,2S+2) tmem_full, [2S+2,2S+4) tmem_empty, then tmem ptr
  uint32_t *tmem_ptr_s = reinterpret_cast<uint32_t *>(bars + 2 * kSlots + 4);
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (kSlots + s); };
  auto tfull_bar = [&](int s) { return bar0 + 8u * (2 * kSlots + s); };
  auto tempty_bar = [&](int s) { return bar0 + 8u * (2 * kSlots + 2 + s); };
  auto slot_addr = [&](int s) { return smem_u32(smem + s * kSlotBytes); };

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.num_m_tiles * p.num_n_tiles;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    if (MODE == kTf32c) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA16) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB16) : "memory");
    }
    for (int s = 0; s < kSlots; ++s) {
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
      int slot = 0;
      uint32_t phase = 0;
      auto advance = [&]() { if (++slot == kSlots) { slot = 0; phase ^= 1u; } };
      auto load_a = [&](uint32_t dst, const CUtensorMap *mp, uint32_t bar, int x, int y) {
        if (p.hint_a) tma_load_2d_hint(dst, mp, bar, x, y, p.hint_a); else tma_load_2d(dst, mp, bar, x, y);
      };
      auto load_b = [&](uint32_t dst, const CUtensorMap *mp, uint32_t bar, int x, int y) {
        if (p.hint_b) tma_load_2d_hint(dst, mp, bar, x, y, p.hint_b); else tma_load_2d(dst, mp, bar, x, y);
      };
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int m = t % p.num_m_tiles, n = t / p.num_m_tiles;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          // slot 0 of the k-block: TF32 operands
          mbar_wait(empty_bar(slot), phase ^ 1u);
          mbar_expect_tx(full_bar(slot), kSlotBytes);
          load_a(slot_addr(slot), &tmA, full_bar(slot), kb * BK, m * BM);
          load_b(slot_addr(slot) + A_BYTES, &tmB, full_bar(slot), kb * BK, p.b_row0 + n * BN);
          advance();
          if (MODE == kTf32x3) {  // residual planes (fp32, TF32-rounded)
            mbar_wait(empty_bar(slot), phase ^ 1u);
            mbar_expect_tx(full_bar(slot), kSlotBytes);
            load_a(slot_addr(slot), &tmA, full_bar(slot), kb * BK, p.a_plane_rows + m * BM);
            load_b(slot_addr(slot) + A_BYTES, &tmB, full_bar(slot), kb * BK,
                   p.b_plane_rows + p.b_row0 + n * BN);
            advance();
          } else if (MODE == kTf32c) {  // bf16 hi and residual planes
            mbar_wait(empty_bar(slot), phase ^ 1u);
            mbar_expect_tx(full_bar(slot), kSlotBytes);
            const uint32_t sb = slot_addr(slot);
            load_a(sb, &tmA16, full_bar(slot), kb * BK, m * BM);
            load_a(sb + A16_BYTES, &tmA16, full_bar(slot), kb * BK, p.a_plane_rows + m * BM);
            load_b(sb + 2 * A16_BYTES, &tmB16, full_bar(slot), kb * BK, p.b_row0 + n * BN);
            load_b(sb + 2 * A16_BYTES + B16_BYTES, &tmB16, full_bar(slot), kb * BK,
                   p.b_plane_rows + p.b_row0 + n * BN);
            advance();
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      int slot = 0;
      uint32_t phase = 0;
      auto advance = [&]() { if (++slot == kSlots) { slot = 0; phase ^= 1u; } };
      int it = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
        const int as = it % kAcc;
        mbar_wait(tempty_bar(as), ((it / kAcc) & 1u) ^ 1u);  // epilogue drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(as * MH * BN);
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(full_bar(slot), phase);
          tc_fence_after();
          const uint32_t a0 = slot_addr(slot), b0 = a0 + A_BYTES;
          if (MODE == kTf32 || MODE == kTf32c) {
#pragma unroll
            for (int k = 0; k < BK / 8; ++k)  // 32 bytes of K per tf32 MMA inside the 128 B atom
#pragma unroll
              for (int h = 0; h < MH; ++h)
                tc_mma_tf32<1>(tmem_d + h * BN, make_desc_sw128(a0 + h * (UM * 128) + k * 32),
                            make_desc_sw128(b0 + k * 32), idesc_tf32(UM), (kb | k) ? 1u : 0u);
            tc_commit(empty_bar(slot));  // frees the slot when these MMAs retire
            advance();
            if (MODE == kTf32c) {
              mbar_wait(full_bar(slot), phase);
              tc_fence_after();
              const uint32_t ahi = slot_addr(slot), alo = ahi + A16_BYTES;
              const uint32_t bhi = ahi + 2 * A16_BYTES, blo = bhi + B16_BYTES;
#pragma unroll
              for (int k = 0; k < BK / 16; ++k) {  // 32 bytes of K per bf16 MMA inside the 64 B atom
#pragma unroll
                for (int h = 0; h < MH; ++h) {
                  tc_mma_bf16<1>(tmem_d + h * BN, make_desc_sw64(alo + h * (UM * 64) + k * 32),
                              make_desc_sw64(bhi + k * 32), idesc_bf16(UM), 1u);
                  tc_mma_bf16<1>(tmem_d + h * BN, make_desc_sw64(ahi + h * (UM * 64) + k * 32),
                              make_desc_sw64(blo + k * 32), idesc_bf16(UM), 1u);
                }
              }
              tc_commit(empty_bar(slot));
              advance();
            }
          } else {  // kTf32x3: both slots are needed by the cross terms
            const int slot_hi = slot;
            advance();
            mbar_wait(full_bar(slot), phase);
            tc_fence_after();
            const uint32_t a1 = slot_addr(slot), b1 = a1 + A_BYTES;
#pragma unroll
            for (int k = 0; k < BK / 8; ++k) {
#pragma unroll
              for (int h = 0; h < MH; ++h) {
                const uint32_t ho = h * (UM * 128) + k * 32;
                tc_mma_tf32<1>(tmem_d + h * BN, make_desc_sw128(a1 + ho), make_desc_sw128(b0 + k * 32),
                            idesc_tf32(UM), (kb | k) ? 1u : 0u);                      // lo * hi
                tc_mma_tf32<1>(tmem_d + h * BN, make_desc_sw128(a0 + ho), make_desc_sw128(b1 + k * 32),
                            idesc_tf32(UM), 1u);                                       // hi * lo
                tc_mma_tf32<1>(tmem_d + h * BN, make_desc_sw128(a0 + ho), make_desc_sw128(b0 + k * 32),
                            idesc_tf32(UM), 1u);                                       // hi * hi
              }
            }
            tc_commit(empty_bar(slot_hi));
            tc_commit(empty_bar(slot));
            advance();
          }
        }
        tc_commit(tfull_bar(as));  // accumulator(s) complete
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
      const int as = it % kAcc;
      mbar_wait(tfull_bar(as), (it / kAcc) & 1u);
      tc_fence_after();
      const int64_t colt = static_cast<int64_t>(n) * BN;
#pragma unroll 1
      for (int h = 0; h < MH; ++h) {
        const int64_t row0 = static_cast<int64_t>(m) * BM + h * UM + quarter * 32;
        if (row0 >= p.Q) break;  // this half holds padding rows only (warp-uniform)
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t v[32];
          const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                                 static_cast<uint32_t>((as * MH + h) * BN + c * 32);
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
              const float val = metric_epilogue(stg[rr * EPI_LD + lane], p.metric, qs, gs);
              if (p.stream_out) __stcs(p.out + r * p.ld + col, val); else p.out[r * p.ld + col] = val;
            }
          }
          __syncwarp();
        }
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


template <int MODE, int MH>
int launch_t(dali_ctx *ctx, const CUtensorMap &tmA, const CUtensorMap &tmB, const CUtensorMap &tmA16,
             const CUtensorMap &tmB16, const UmmaParams &p) {
  using C = Cfg<MH>;
  if (int rc = ensure_dyn_smem(ctx, reinterpret_cast<const void *>(&distmat_umma_kernel<MODE, MH>), C::kSmemBytes))
    return rc;
  const int tiles = p.num_m_tiles * p.num_n_tiles;
  const int grid = tiles < ctx->num_sms ? tiles : ctx->num_sms;
  KTimer t(ctx, DALI_K_DISTMAT);
  distmat_umma_kernel<MODE, MH><<<grid, kThreads, C::kSmemBytes, ctx->stream>>>(tmA, tmB, tmA16,
                                                                               tmB16, p);
  DALI_CUDA_OK(ctx, cudaGetLastError());
  return DALI_OK;
}

}  // namespace

// q32/g32: [npl32][rows_pad][Dp] fp32 planes (plane 0 = TF32-rounded operand, plane 1 = residual,
// TF32X3 only); q16/g16: [2][rows_pad][Dp] bf16 planes (hi16, lo16; TF32C only).
// rows_pad multiples of 256, Dp multiple of 32.  mode: DALI_PREC_TF32 / TF32X3 / TF32C.
int launch_distmat_umma1(dali_ctx *ctx, const float *q32, const float *g32, const void *q16,
                        const void *g16, int64_t Q, int64_t G, int64_t Dp, int64_t q_rows_pad,
                        int64_t g_rows_pad, int64_t g_row0, int precision, int metric,
                        const float *qsq, const float *gsq, float *out, int64_t ld) {
  if (Q == 0 || G == 0) return DALI_OK;
  const int npl32 = precision == DALI_PREC_TF32X3 ? 2 : 1;
  if (Dp % BK != 0 || q_rows_pad % 256 != 0 || g_rows_pad % BN != 0 || g_row0 % BN != 0)
    return set_err(ctx, DALI_ERR_INVALID, "umma operands must be padded (rows 128/256, D 32)");
  if (q_rows_pad * 2 > INT32_MAX || g_rows_pad * 2 > INT32_MAX)
    return set_err(ctx, DALI_ERR_UNSUPPORTED, "operand too tall for one tensor map");
  // 256-row CTA tiles unless the query block is a single 128-row tile (or forced by the env)
  static const char *force_mh = getenv("DALI_UMMA_MH");
  int mh = 1;  // the 256-row tile (MH=2) measured slower in round 1 (shallower ring, exposed epilogue)
  if (force_mh) mh = atoi(force_mh) == 2 ? 2 : 1;
  const int BM = UM * mh;
  CUtensorMap tmA, tmB, tmA16, tmB16;
  int rc = make_map(ctx, &tmA, q32, q_rows_pad * npl32, Dp, BM, false);
  if (rc) return rc;
  rc = make_map(ctx, &tmB, g32, g_rows_pad * npl32, Dp, BN, false);
  if (rc) return rc;
  if (precision == DALI_PREC_TF32C) {
    rc = make_map(ctx, &tmA16, q16, q_rows_pad * 2, Dp, BM, true);
    if (rc) return rc;
    rc = make_map(ctx, &tmB16, g16, g_rows_pad * 2, Dp, BN, true);
    if (rc) return rc;
  } else {
    tmA16 = tmA;
    tmB16 = tmB;
  }
  UmmaParams p;
  p.Q = Q; p.G = G;
  p.num_m_tiles = static_cast<int>((Q + BM - 1) / BM);
  p.num_n_tiles = static_cast<int>((G + BN - 1) / BN);
  p.num_kb = static_cast<int>(Dp / BK);
  p.a_plane_rows = static_cast<int32_t>(q_rows_pad);
  p.b_plane_rows = static_cast<int32_t>(g_rows_pad);
  p.b_row0 = static_cast<int32_t>(g_row0);
  p.metric = metric; p.qsq = qsq; p.gsq = gsq; p.out = out; p.ld = ld;
  // L2 policies (CUTLASS TMA::CacheHintSm90 encodings): EVICT_FIRST 0x12F0.., EVICT_LAST 0x14F0..
  static const char *env_hint = getenv("DALI_UMMA_HINTS");
  const int hints = env_hint ? atoi(env_hint) : 0;
  p.hint_a = (hints & 1) ? 0x14F0000000000000ull : 0ull;  // queries: reused by every gallery tile
  p.hint_b = (hints & 2) ? 0x12F0000000000000ull : 0ull;  // gallery tile: short-lived
  p.stream_out = (hints & 4) ? 1 : 0;
  if (static_cast<int64_t>(p.num_m_tiles) * p.num_n_tiles > INT32_MAX)
    return set_err(ctx, DALI_ERR_UNSUPPORTED, "too many tiles (chunk the queries)");
  if (mh == 2) {
    switch (precision) {
      case DALI_PREC_TF32: return launch_t<kTf32, 2>(ctx, tmA, tmB, tmA16, tmB16, p);
      case DALI_PREC_TF32X3: return launch_t<kTf32x3, 2>(ctx, tmA, tmB, tmA16, tmB16, p);
      case DALI_PREC_TF32C: return launch_t<kTf32c, 2>(ctx, tmA, tmB, tmA16, tmB16, p);
      default: break;
    }
  } else {
    switch (precision) {
      case DALI_PREC_TF32: return launch_t<kTf32, 1>(ctx, tmA, tmB, tmA16, tmB16, p);
      case DALI_PREC_TF32X3: return launch_t<kTf32x3, 1>(ctx, tmA, tmB, tmA16, tmB16, p);
      case DALI_PREC_TF32C: return launch_t<kTf32c, 1>(ctx, tmA, tmB, tmA16, tmB16, p);
      default: break;
    }
  }
  return set_err(ctx, DALI_ERR_INVALID, "not a tensor-core precision");
}

}  // namespace dali
