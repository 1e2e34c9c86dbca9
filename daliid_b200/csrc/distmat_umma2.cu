// Q x G x D distance contraction on CTA pairs (tcgen05 cta_group::2), the default tensor-core
// path (SURVEY 8a rows a2/a2', a7 fused; precisions DALI_PREC_TF32, TF32X3, TF32C).
//
//   out[i,j] = epilogue( sum_k A[i,k] * B[j,k] )        A = prepared queries, B = gallery
//
// Why pairs: with one CTA per 128 x 256 tile every k-block moves 48 KiB of fp32 operands
// through L2 -> shared memory -> tensor core for 0.38 us of TF32 MMA time: 125 GB/s per SM in
// each direction, which is the shared-memory port (128 B/clk) and ~15 TB/s of L2 fabric chip
// wide; round 1 measured the tensor pipe 72-76 % active because of it.  Two CTAs of one TPC
// computing a 256 x 256 tile together each stage only HALF of the gallery tile (the MMA reads
// both halves), so operand traffic per FLOP drops by a third and the B-side shared-memory
// reads by half.
//
// Structure (persistent, one CTA pair per TPC, 192 threads per CTA):
//   warp 0    TMA producer (both CTAs): its 128 query rows + its 128 gallery rows of the k-block
//             into a ring of six 32 KiB slots; transaction bytes of BOTH CTAs are reported to the
//             leader's `full` barrier (cp.async.bulk.tensor ... .cta_group::2)
//   warp 1    allocates TMEM (both CTAs, 512 columns = two 128x256 fp32 accumulators each); in the
//             leader CTA one lane issues tcgen05.mma.cta_group::2 (M=256, N=256) and releases
//             slots / publishes accumulators with multicast tcgen05.commit to both CTAs
//   warps 2-5 epilogue (both CTAs, each drains its own 128 accumulator rows), overlapped with
//             the next tile's MMAs through the second accumulator:
//               kStore   metric -> transpose through shared memory -> coalesced row stores
//               kFilter  metric -> compare with the row's running k-th best distance -> append
//                        the rare survivors to the row's candidate list (fused top-k, the
//                        distance matrix is never written: BASELINE config 5)
//
// replaces  1.0 - torch.mm(q, g.T)   validateModels.py:47, evaluate.py:260-267,291,
//           evaluate_ensembled_models.py:281,300, evaluateCleanATModels.py:109,121,124
//           torch.argsort(distmat, dim=1)[:, :20]   validateModels.py:93 (kFilter)
#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "umma_common.cuh"

namespace dali {

int launch_distmat_umma1(dali_ctx *ctx, const float *q32, const float *g32, const void *q16,
                         const void *g16, int64_t Q, int64_t G, int64_t Dp, int64_t q_rows_pad,
                         int64_t g_rows_pad, int64_t g_row0, int precision, int metric,
                         const float *qsq, const float *gsq, float *out, int64_t ld);

namespace {

using namespace umma;

constexpr int kThreads = 192;
// Warp roles.  The scheduler of an SM sub-partition prefers the warp with the HIGHEST id, and the
// counting epilogue (kCount) issues ALU work almost every cycle: with the TMA and MMA threads in
// warps 0 / 1 they were starved by epilogue warps 4 / 5 on their sub-partitions and the tensor pipe
// idled (45 % active, ncu r2_fused_b).  They are warps 4 / 5 now: they issue the moment they wake
// up, and sleep inside mbarrier.try_wait otherwise.  Epilogue warps 0-3 own TMEM lane quarters 0-3.
constexpr int kTmaWarp = 4, kMmaWarp = 5;
constexpr int PM = 2 * UM;                 // rows of a pair tile
constexpr int HB = BN / 2;                 // gallery rows staged by each CTA
constexpr int A_BYTES = UM * BK * 4;       // 16 KiB fp32 (rows of 128 B, SWIZZLE_128B)
constexpr int B_BYTES = HB * BK * 4;       // 16 KiB
constexpr int A16_BYTES = UM * BK * 2;     // 8 KiB bf16 (rows of 64 B, SWIZZLE_64B)
constexpr int B16_BYTES = HB * BK * 2;     // 8 KiB
constexpr int kSlotBytes = A_BYTES + B_BYTES;  // 32 KiB == 2*A16 + 2*B16
constexpr int kSlots = 6;
constexpr int kBarBytes = 256;
constexpr int kStageBytes = 4 * 2 * 4096;  // four epilogue warps x two 32x32 fp32 store tiles
constexpr int kSmemBytes = kSlots * kSlotBytes + kStageBytes + kBarBytes + 1024;
static_assert(8 * (2 * kSlots + 4) + 8 <= kBarBytes, "barrier block too small");

// kBand + kCount: fused distance + positive-rank counting (the Q x G matrix is never written):
//   kBand   the few tiles that hold the same-identity (query, gallery) pairs: their distances go
//           to a small scratch, the matches' order keys to keys[M]
//   kCount  every other tile straight from TMEM, and the band tiles from the scratch: each
//           epilogue thread owns a query row, holds that query's thresholds (the distances of its
//           valid positives) in registers and counts the columns below each of them
enum Epi { kStore = 0, kFilter = 1, kBand = 2, kCount = 3, kAccum = 4 };  // kAccum: kStore + running mean

// mean accumulator loss per MMA in units of 2^-24 * s with the fixed-point hi plane, fitted to
// tests/probes/trunc_probe.py on B200 (constant over D = 512 .. 4096 to +-0.01)
constexpr double kKappaInterleaved = 0.24, kKappaTwoPass = 0.22;

struct Umma2Params {
  int64_t Q, G;
  int num_m_pairs, num_n_tiles, num_kb;
  int nband;            // n-tiles per L2 band (1 = plain m-fastest order)
  int32_t a_plane_rows, b_plane_rows;  // row offset of plane 1 inside the tensor maps
  int32_t b_row0;                      // first gallery row of this slab inside the B planes
  int metric;
  const float *qsq, *gsq;
  // kStore
  float *out;
  int64_t ld;
  int tma_out;          // 1: rows of `out` are 16-byte aligned, tiles leave through TMA stores
  // kStore, mean fusion over several launches (fuse.cu's operation order, ((d0+d1)+d2)/n):
  float *acc;           // running sum matrix (16-byte aligned rows, written through tmAcc)
  int64_t ld_acc;
  int acc_mode;         // 0 none; 1 acc = d; 2 acc = acc + d; 3 acc = (acc + d) / acc_div
  float acc_div;
  int store_out;        // 0: only `acc` is written
  // kFilter
  const float *thr;     // [Q] running k-th best distance of the row (+inf / -inf: accept all)
  int32_t *cand_cnt;    // [Q] entries appended to the row's list
  uint64_t *cand;       // [Q][cap] composites (order key << 32 | gallery id)
  int cap;
  int largest;
  int direct;           // 1: every column j is written to slot j of the list (first chunk)
  int32_t id_base;      // gallery id of column 0
  float acc_scale;      // 2^-24 for kF16x3 (operands carry a factor 2^12 each), else 1; times the
                        // truncation compensation (see f16x3_schedule)
  int two_pass;         // kF16x3: all correction MMAs (lo*hi, hi*lo) of a tile first, then hi*hi
  FusedArgs f;          // kBand / kCount
};

// ---- kStore epilogue helpers -----------------------------------------------------------------
__device__ __forceinline__ void sts128(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}
__device__ __forceinline__ void sts32(uint32_t addr, float a) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(a) : "memory");
}
__device__ __forceinline__ float lds32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, uint32_t src, int32_t x, int32_t y) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(map), "r"(src), "r"(x), "r"(y)
               : "memory");
}

// The metric on 32 accumulator columns of one query row, in registers (same expressions as the
// filter epilogue and umma::metric_epilogue).  KIND 0: alpha * acc + beta (cosine, dot);
// 1: |q|^2 + |g|^2 - 2 acc; 2: its square root.
template <int KIND>
__device__ __forceinline__ void apply_metric(uint32_t (&v)[32], float alpha, float beta, float scale,
                                             float qs, const float *__restrict__ gs, int lim) {
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const float acc = __uint_as_float(v[j]);
    float d;
    if (KIND == 0) {
      d = fmaf(acc, alpha, beta);
    } else {
      d = fmaf(-2.0f, acc * scale, qs + __ldg(gs + (j < lim ? j : 0)));
      if (KIND == 2) d = sqrtf(fmaxf(d, 1e-30f));
    }
    v[j] = __float_as_uint(d);
  }
}

// ---- kFilter epilogue: 32 accumulator columns of one query row -------------------------------
struct FilterRow {
  float thr, qs, alpha, beta, scale;  // alpha already carries the accumulator scale
  uint32_t flip;
  uint64_t *list;
  int32_t *cnt;
  int cap;
  bool row_ok;
};

// v[j] for a run-time j without spilling the accumulator registers: a 5-level select tree.
__device__ __forceinline__ uint32_t pick32(const uint32_t (&v)[32], int j) {
  uint32_t a[16], b[8], c[4], d[2];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = (j & 1) ? v[2 * i + 1] : v[2 * i];
#pragma unroll
  for (int i = 0; i < 8; ++i) b[i] = (j & 2) ? a[2 * i + 1] : a[2 * i];
#pragma unroll
  for (int i = 0; i < 4; ++i) c[i] = (j & 4) ? b[2 * i + 1] : b[2 * i];
#pragma unroll
  for (int i = 0; i < 2; ++i) d[i] = (j & 8) ? c[2 * i + 1] : c[2 * i];
  return (j & 16) ? d[1] : d[0];
}

// KIND 0: d = alpha * acc + beta (cosine: 1 - acc, dot: acc); 1: |q|^2 + |g|^2 - 2 acc; 2: its
// square root (the expressions of umma::metric_epilogue, bit for bit).  SEL 0: keep d <= thr,
// 1: keep d >= thr, 2: keep every column (first chunk, column j goes to slot j).  NaN always
// passes, as it sorts last / first like everywhere else.
// The hot loop is branch free (FFMA, FSETP, predicated OR into a survivor mask): survivors are
// about one element in a thousand, and a branch per element paced the whole kernel (measured:
// 11.8 ms against 4.9 ms of MMA time at 16k x 262k, D = 512).
template <int KIND, int SEL>
__device__ __forceinline__ void filter_cols(const uint32_t (&v)[32], const FilterRow &fr,
                                            const float *__restrict__ gs, int lim, uint32_t id0,
                                            int slot0) {
  auto dist_of = [&](uint32_t bits, int j) {
    const float acc = __uint_as_float(bits);
    float d;
    if (KIND == 0) {
      d = fmaf(acc, fr.alpha, fr.beta);
    } else {
      d = fmaf(-2.0f, acc * fr.scale, fr.qs + __ldg(gs + (j < lim ? j : 0)));
      if (KIND == 2) d = sqrtf(fmaxf(d, 1e-30f));
    }
    return d;
  };
  if (SEL == 2) {
    if (!fr.row_ok) return;
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j < lim) fr.list[slot0 + j] = composite(dist_key(dist_of(v[j], j)) ^ fr.flip, id0 + j);
    return;
  }
  uint32_t mask = 0;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const float d = dist_of(v[j], j);
    const bool pass = SEL == 1 ? !(d < fr.thr) : !(d > fr.thr);
    mask |= pass ? (1u << j) : 0u;
  }
  if (lim < 32) mask &= (1u << lim) - 1u;
  if (!fr.row_ok) mask = 0;
  if (mask) {  // rare, divergent: ONE counter round trip for all survivors of these 32 columns (a
               // returning atomic per survivor paced the short early chunks, where a row keeps ~10
               // of a tile's 256 columns)
    int pos = atomicAdd(fr.cnt, __popc(mask));
    while (mask) {
      const int j = __ffs(mask) - 1;
      mask &= mask - 1;
      const float d = dist_of(pick32(v, j), j);
      if (pos < fr.cap) fr.list[pos] = composite(dist_key(d) ^ fr.flip, id0 + j);
      ++pos;
    }
  }
  // the next tcgen05.ld is .sync.aligned: lanes that took the loop above must have rejoined
  // (without this the lanes with no survivor ran ahead and the load landed in registers the
  // others were still reading)
  __syncwarp();
}


// ---- kCount epilogue --------------------------------------------------------------------------
constexpr uint32_t kNanBits = 0x7FC00000u;

struct CountRow {
  int64_t o;       // offset of the query's match list
  int nvq;         // its valid positives
  bool row_ok;
};

constexpr int kPassThr = 31;  // thresholds per pass: slot 31 of the warp's table is always +inf

__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// One tile for one warp (lane = query row): the tile's columns against the lane's thresholds -- the
// distances of its valid positives, SORTED ascending by the threshold-sort kernel (rank.cu).  Each
// column finds, by a 5-level branch-free binary search, how many thresholds are <= it, and bumps the
// counter of that bucket; count_below(T_k) is then the prefix sum of the buckets (prefix kernel).
// Thresholds and counters of a pass (31 thresholds + a +inf sentinel, 32 buckets) live in the warp's
// 8 KB of shared memory, [slot][lane]: every access is conflict free and lane private.
//   per column: 5 x (LDS, FSETP, predicated add) + tie check (LDS, FSETP) + counter (LDS, IADD, STS)
//   = ~20 instructions and 8 shared-memory wavefronts -- independent of the number of positives.
// (A direct compare of every column with every threshold held in registers was tried first: 3
// instructions per PAIR on the 16-lane ALU pipe, 25 k instructions per tile and warp, 50 us per tile
// against 37 us of MMAs at D = 2048 -- slower than writing the matrix and reading it back.)
// A column that EQUALS a threshold (always the positive's own column; a genuine tie about once per
// hundred positives) raises `eq`; the rare path then decides by gallery id.  NaN columns (padding,
// zero-norm rows) pass every threshold and land in the last bucket, which is never read.
// SRC 0: accumulators from TMEM; 1: distances of a band tile from the scratch.
template <int KIND, int SRC>
__device__ __forceinline__ void count_tile(const Umma2Params &p, const CountRow &cr, int wmax, uint32_t tbase,
                                           const float *__restrict__ srow, int64_t colt, float alpha,
                                           float beta, float qs, uint32_t sT, uint32_t sC, int lane) {
  // sT, sC: shared-space byte addresses of this lane's column of the two [32][32] tables
  if (p.f.dbg & 2) return;
  for (int t0 = 0; t0 < wmax; t0 += kPassThr) {  // warp-uniform
    const int nmine = min(max(cr.nvq - t0, 0), kPassThr);
    {
      // all 32 loads in flight together (each lane reads its own query's list: L2 latency, once)
      float Tl[32];
#pragma unroll
      for (int i = 0; i < 32; ++i)  // slots beyond the lane's positives: +inf, nothing reaches them
        Tl[i] = i < nmine ? __ldg(p.f.sorted_thr + cr.o + t0 + i) : INFINITY;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        sts_u32(sT + i * 128, __float_as_uint(Tl[i]));
        sts_u32(sC + i * 128, 0u);
      }
    }
    __syncwarp();
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      const int64_t col0 = colt + c * 32;
      if (col0 >= p.G) break;
      uint32_t v[32];
      const int lim = p.G - col0 < 32 ? static_cast<int>(p.G - col0) : 32;
      if (SRC == 0) {
        tc_ld_32x32(tbase + c * 32, v);
        tc_wait_ld();
        apply_metric<KIND>(v, alpha, beta, p.acc_scale, qs, p.gsq ? p.gsq + col0 : nullptr, lim);
      } else {
        const uint4 *src = reinterpret_cast<const uint4 *>(srow + c * 32);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint4 x = __ldg(src + k);
          v[4 * k] = x.x; v[4 * k + 1] = x.y; v[4 * k + 2] = x.z; v[4 * k + 3] = x.w;
        }
      }
      if (p.f.dbg & 1) { if (v[0] == 0x12345678u) *p.f.flag = 2; continue; }
      if (lim < 32) {  // columns beyond the gallery (zero padding rows): NaN -> last bucket
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j >= lim) v[j] = kNanBits;
      }
      // level by level over all 32 columns: the 32 loads of a level are independent and in flight
      // together (one column after the other would wait ~30 cycles for each of its seven loads)
      uint32_t off[32];  // 128 * #{thresholds <= d_j} (bytes into the lane's column of the table)
#pragma unroll
      for (int j = 0; j < 32; ++j) off[j] = 0;
      if (!(p.f.dbg & 16))
#pragma unroll
      for (int step = 16; step >= 1; step >>= 1) {
        uint32_t t[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) t[j] = lds_u32(sT + off[j] + (step - 1) * 128);
#pragma unroll
        for (int j = 0; j < 32; ++j)
          off[j] += !(__uint_as_float(v[j]) < __uint_as_float(t[j])) ? step * 128 : 0;  // t <= d, or d NaN
      }
      bool eq = false;
      if (!(p.f.dbg & 8)) {
        uint32_t t[32];  // the threshold just below the bucket: does it tie with d?
#pragma unroll
        for (int j = 0; j < 32; ++j) t[j] = lds_u32(sT + (off[j] ? off[j] - 128 : 0));
#pragma unroll
        for (int j = 0; j < 32; ++j) eq |= __uint_as_float(t[j]) == __uint_as_float(v[j]);
      }
      if (!(p.f.dbg & 4))
#pragma unroll
      for (int j = 0; j < 32; ++j)  // off / 128 in [0, 31]: bucket 31 (NaN, beyond every threshold) is never read
        asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(sC + off[j]) : "memory");
      if (eq && (p.f.dbg & 64)) atomicAdd(p.f.flag + 1, 1);
      if (eq && !(p.f.dbg & 32)) {
        // Rare path.  All global loads are issued up front in three independent batches (gallery ids
        // of the 32 columns, match-list slots and gallery ids of the lane's thresholds): one load
        // chain per threshold made a band tile -- where every positive ties with its own column --
        // cost ~30 us per column group.
        int32_t ord[32], slots[32], gps[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) ord[j] = p.f.gid_base + __ldg(p.f.order + (col0 + j < p.G ? col0 + j : p.G - 1));
#pragma unroll
        for (int i = 0; i < 32; ++i) slots[i] = i < nmine ? __ldg(p.f.sorted_slot + cr.o + t0 + i) : 0;
#pragma unroll
        for (int i = 0; i < 32; ++i) gps[i] = i < nmine ? __ldg(p.f.gid + cr.o + slots[i]) : 0;
#pragma unroll 1
        for (int i = 0; i < nmine; ++i) {  // slots[] / gps[] are indexed at run time: local memory
          const float Ti = __uint_as_float(lds_u32(sT + i * 128));
          const int32_t gp = gps[i];
          int fix = 0;
#pragma unroll
          for (int j = 0; j < 32; ++j) fix += (__uint_as_float(v[j]) == Ti && ord[j] < gp) ? 1 : 0;
          if (fix) atomicAdd(p.f.counts + cr.o + slots[i], fix);
        }
      }
      __syncwarp();  // the next tcgen05.ld is .sync.aligned
    }
    // buckets -> global histogram (sorted positions); bucket b holds the columns with exactly b
    // thresholds of this pass <= them
    for (int b = 0; b < nmine; ++b) {
      const uint32_t ci = lds_u32(sC + b * 128);
      if (ci) atomicAdd(p.f.hist + cr.o + t0 + b, static_cast<int32_t>(ci));
    }
    __syncwarp();
  }
}

template <int MODE, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
distmat_umma2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const __grid_constant__ CUtensorMap tmA16,
                     const __grid_constant__ CUtensorMap tmB16,
                     const __grid_constant__ CUtensorMap tmOut,
                     const __grid_constant__ CUtensorMap tmAcc, const Umma2Params p) {
  extern __shared__ uint8_t smem_raw[];
  // the dynamic window starts at the same offset in both CTAs, so this rounding is identical
  // in the pair (the MMA applies the leader's operand offsets to the peer's shared memory)
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                              ~uintptr_t(1023));
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + kSlots * kSlotBytes + kStageBytes);
  // bars: [0,S) full (leader's are used), [S,2S) empty, [2S,2S+2) tmem_full,
  //       [2S+2,2S+4) tmem_empty (leader's are used), then the TMEM base pointer
  uint32_t *tmem_ptr_s = reinterpret_cast<uint32_t *>(bars + 2 * kSlots + 4);
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (kSlots + s); };
  auto tfull_bar = [&](int s) { return bar0 + 8u * (2 * kSlots + s); };
  auto tempty_bar = [&](int s) { return bar0 + 8u * (2 * kSlots + 2 + s); };
  auto slot_addr = [&](int s) { return smem_u32(smem + s * kSlotBytes); };

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();  // 0 = leader (issues the MMAs)
  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;
  const int num_tiles = p.f.tiles ? p.f.num_list : p.num_m_pairs * p.num_n_tiles;
  // Tile order.  Bands of `nband` gallery tiles: inside a band the tiles of one query pair-tile
  // are consecutive, so the ~74 tiles in flight share a few query tiles and the band's gallery
  // tiles stay in L2, while the query planes stream from HBM once per band.  nband = 1 is the
  // m-fastest order (queries resident in L2, gallery streamed once) used when the query planes
  // are small.
  const int band_tiles = p.num_m_pairs * p.nband;
  auto tile_mn = [&](int t, int &m, int &n) {
    if (p.f.tiles) {  // explicit list (fused counting: band tiles / all the others)
      const uint32_t mn = __ldg(p.f.tiles + t);
      m = static_cast<int>(mn >> 16);
      n = static_cast<int>(mn & 0xFFFFu);
      return;
    }
    const int band = t / band_tiles;
    const int r = t - band * band_tiles;
    const int n0 = band * p.nband;
    const int nb = min(p.nband, p.num_n_tiles - n0);
    if (p.nband == 1) { m = r; n = n0; return; }
    m = r / nb;
    n = n0 + (r - m * nb);
  };

  if (warp == kTmaWarp && lane == 0) {
    if (MODE != kF16x3 && MODE != kF16) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    }
    if (MODE == kTf32c || MODE == kF16x3 || MODE == kF16) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA16) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB16) : "memory");
    }
    for (int s = 0; s < kSlots; ++s) {
      mbar_init(full_bar(s), 1);   // the leader's arrive.expect_tx
      mbar_init(empty_bar(s), 1);  // one multicast commit
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);   // one multicast commit
      mbar_init(tempty_bar(s), 8);  // four epilogue warps in each CTA of the pair
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;"
                 ::"r"(smem_u32(tmem_ptr_s))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();  // barriers of both CTAs are initialised before any remote signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;

  if (warp == kTmaWarp) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int slot = 0;
      uint32_t phase = 0;
      auto advance = [&]() { if (++slot == kSlots) { slot = 0; phase ^= 1u; } };
      const uint32_t full0 = mapa_u32(full_bar(0), 0);  // the leader's full barriers
      for (int t = pair; t < num_tiles; t += num_pairs) {
        int m, n;
        tile_mn(t, m, n);
        const int arow = m * PM + static_cast<int>(rank) * UM;
        const int brow = p.b_row0 + n * BN + static_cast<int>(rank) * HB;
        if (MODE == kF16x3 && p.two_pass) {
          // pass 1: the four planes of a k-block per slot (corrections); pass 2: the hi planes of
          // TWO k-blocks per slot (same bytes per slot, half the barrier traffic per byte)
          for (int kb = 0; kb < p.num_kb; ++kb) {
            mbar_wait(empty_bar(slot), phase ^ 1u);
            if (rank == 0) mbar_expect_tx(full_bar(slot), 2 * kSlotBytes);
            const uint32_t fbh = full0 + 8u * slot;
            const uint32_t sb = slot_addr(slot);
            tma_load_2d_pair(sb, &tmA16, fbh, kb * BK, arow);
            tma_load_2d_pair(sb + A16_BYTES, &tmA16, fbh, kb * BK, p.a_plane_rows + arow);
            tma_load_2d_pair(sb + 2 * A16_BYTES, &tmB16, fbh, kb * BK, brow);
            tma_load_2d_pair(sb + 2 * A16_BYTES + B16_BYTES, &tmB16, fbh, kb * BK, p.b_plane_rows + brow);
            advance();
          }
          for (int kb = 0; kb < p.num_kb; kb += 2) {
            const bool two = kb + 1 < p.num_kb;
            mbar_wait(empty_bar(slot), phase ^ 1u);
            if (rank == 0) mbar_expect_tx(full_bar(slot), two ? 2 * kSlotBytes : kSlotBytes);
            const uint32_t fbh = full0 + 8u * slot;
            const uint32_t sb = slot_addr(slot);
            tma_load_2d_pair(sb, &tmA16, fbh, kb * BK, arow);
            tma_load_2d_pair(sb + 2 * A16_BYTES, &tmB16, fbh, kb * BK, brow);
            if (two) {
              tma_load_2d_pair(sb + A16_BYTES, &tmA16, fbh, (kb + 1) * BK, arow);
              tma_load_2d_pair(sb + 2 * A16_BYTES + B16_BYTES, &tmB16, fbh, (kb + 1) * BK, brow);
            }
            advance();
          }
          continue;
        }
        for (int kb = 0; kb < p.num_kb; ++kb) {
          if (MODE == kF16x3 || MODE == kF16) {  // one slot: fp16 hi (and residual) planes of both operands
            mbar_wait(empty_bar(slot), phase ^ 1u);
            if (rank == 0) mbar_expect_tx(full_bar(slot), MODE == kF16 ? kSlotBytes : 2 * kSlotBytes);
            const uint32_t fbh = full0 + 8u * slot;
            const uint32_t sb = slot_addr(slot);
            tma_load_2d_pair(sb, &tmA16, fbh, kb * BK, arow);
            tma_load_2d_pair(sb + 2 * A16_BYTES, &tmB16, fbh, kb * BK, brow);
            if (MODE == kF16x3) {
              tma_load_2d_pair(sb + A16_BYTES, &tmA16, fbh, kb * BK, p.a_plane_rows + arow);
              tma_load_2d_pair(sb + 2 * A16_BYTES + B16_BYTES, &tmB16, fbh, kb * BK,
                               p.b_plane_rows + brow);
            }
            advance();
            continue;
          }
          mbar_wait(empty_bar(slot), phase ^ 1u);
          if (rank == 0) mbar_expect_tx(full_bar(slot), 2 * kSlotBytes);
          const uint32_t fb = full0 + 8u * slot;
          tma_load_2d_pair(slot_addr(slot), &tmA, fb, kb * BK, arow);
          tma_load_2d_pair(slot_addr(slot) + A_BYTES, &tmB, fb, kb * BK, brow);
          advance();
          if (MODE == kTf32x3) {  // residual planes (fp32, TF32-rounded)
            mbar_wait(empty_bar(slot), phase ^ 1u);
            if (rank == 0) mbar_expect_tx(full_bar(slot), 2 * kSlotBytes);
            const uint32_t fb1 = full0 + 8u * slot;
            tma_load_2d_pair(slot_addr(slot), &tmA, fb1, kb * BK, p.a_plane_rows + arow);
            tma_load_2d_pair(slot_addr(slot) + A_BYTES, &tmB, fb1, kb * BK, p.b_plane_rows + brow);
            advance();
          } else if (MODE == kTf32c) {  // bf16 hi and residual planes
            mbar_wait(empty_bar(slot), phase ^ 1u);
            if (rank == 0) mbar_expect_tx(full_bar(slot), 2 * kSlotBytes);
            const uint32_t fb1 = full0 + 8u * slot;
            const uint32_t sb = slot_addr(slot);
            tma_load_2d_pair(sb, &tmA16, fb1, kb * BK, arow);
            tma_load_2d_pair(sb + A16_BYTES, &tmA16, fb1, kb * BK, p.a_plane_rows + arow);
            tma_load_2d_pair(sb + 2 * A16_BYTES, &tmB16, fb1, kb * BK, brow);
            tma_load_2d_pair(sb + 2 * A16_BYTES + B16_BYTES, &tmB16, fb1, kb * BK,
                             p.b_plane_rows + brow);
            advance();
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (rank == 0 && lane == 0) {
      constexpr uint32_t kIdT = idesc_tf32(PM), kIdB = idesc_bf16(PM);
      int slot = 0;
      uint32_t phase = 0;
      auto advance = [&]() { if (++slot == kSlots) { slot = 0; phase ^= 1u; } };
      int it = 0;
      for (int t = pair; t < num_tiles; t += num_pairs, ++it) {
        const int as = it & 1;
        // both CTAs' epilogues drained this accumulator
        mbar_wait_cluster(tempty_bar(as), ((it >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(as * BN);
        if (MODE == kF16x3 && p.two_pass) {
          // The tensor core truncates its fp32 accumulator after every MMA; the loss is a fraction
          // of an ulp of the RUNNING SUM per instruction.  The correction products are tiny, so they
          // are accumulated first, while the sum is small and nothing is lost; only the D/16 hi*hi
          // instructions then run on a large accumulator -- a third of the interleaved schedule's.
          constexpr uint32_t kIdH = idesc_f16(PM);
          for (int kb = 0; kb < p.num_kb; ++kb) {
            mbar_wait(full_bar(slot), phase);
            tc_fence_after();
            const uint32_t ahi = slot_addr(slot), alo = ahi + A16_BYTES;
            const uint32_t bhi = ahi + 2 * A16_BYTES, blo = bhi + B16_BYTES;
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              tc_mma_bf16<2>(tmem_d, make_desc_sw64(alo + k * 32), make_desc_sw64(bhi + k * 32), kIdH,
                             (kb | k) ? 1u : 0u);                                   // lo * hi
              tc_mma_bf16<2>(tmem_d, make_desc_sw64(ahi + k * 32), make_desc_sw64(blo + k * 32), kIdH, 1u);  // hi * lo
            }
            tc_commit_pair(empty_bar(slot), 3);
            advance();
          }
          for (int kb = 0; kb < p.num_kb; kb += 2) {
            mbar_wait(full_bar(slot), phase);
            tc_fence_after();
            const uint32_t a0h = slot_addr(slot), b0h = a0h + 2 * A16_BYTES;
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              tc_mma_bf16<2>(tmem_d, make_desc_sw64(a0h + k * 32), make_desc_sw64(b0h + k * 32), kIdH, 1u);
            if (kb + 1 < p.num_kb) {
#pragma unroll
              for (int k = 0; k < BK / 16; ++k)
                tc_mma_bf16<2>(tmem_d, make_desc_sw64(a0h + A16_BYTES + k * 32),
                               make_desc_sw64(b0h + B16_BYTES + k * 32), kIdH, 1u);
            }
            tc_commit_pair(empty_bar(slot), 3);
            advance();
          }
          tc_commit_pair(tfull_bar(as), 3);
          continue;
        }
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(full_bar(slot), phase);
          tc_fence_after();
          const uint32_t a0 = slot_addr(slot), b0 = a0 + A_BYTES;
          if (MODE == kF16) {
            constexpr uint32_t kIdH = idesc_f16(PM);
            const uint32_t ahi = a0, bhi = ahi + 2 * A16_BYTES;
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              tc_mma_bf16<2>(tmem_d, make_desc_sw64(ahi + k * 32), make_desc_sw64(bhi + k * 32),
                             kIdH, (kb | k) ? 1u : 0u);                            // hi * hi
            tc_commit_pair(empty_bar(slot), 3);
            advance();
          } else if (MODE == kF16x3) {
            constexpr uint32_t kIdH = idesc_f16(PM);
            const uint32_t ahi = a0, alo = ahi + A16_BYTES;
            const uint32_t bhi = ahi + 2 * A16_BYTES, blo = bhi + B16_BYTES;
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {  // 32 bytes of K per fp16 MMA inside the 64 B atom
              tc_mma_bf16<2>(tmem_d, make_desc_sw64(alo + k * 32), make_desc_sw64(bhi + k * 32),
                             kIdH, (kb | k) ? 1u : 0u);                            // lo * hi
              tc_mma_bf16<2>(tmem_d, make_desc_sw64(ahi + k * 32), make_desc_sw64(blo + k * 32),
                             kIdH, 1u);                                             // hi * lo
              tc_mma_bf16<2>(tmem_d, make_desc_sw64(ahi + k * 32), make_desc_sw64(bhi + k * 32),
                             kIdH, 1u);                                             // hi * hi
            }
            tc_commit_pair(empty_bar(slot), 3);
            advance();
          } else if (MODE == kTf32 || MODE == kTf32c) {
#pragma unroll
            for (int k = 0; k < BK / 8; ++k)  // 32 bytes of K per tf32 MMA inside the 128 B atom
              tc_mma_tf32<2>(tmem_d, make_desc_sw128(a0 + k * 32), make_desc_sw128(b0 + k * 32),
                             kIdT, (kb | k) ? 1u : 0u);
            tc_commit_pair(empty_bar(slot), 3);  // frees the slot in both CTAs
            advance();
            if (MODE == kTf32c) {
              mbar_wait(full_bar(slot), phase);
              tc_fence_after();
              const uint32_t ahi = slot_addr(slot), alo = ahi + A16_BYTES;
              const uint32_t bhi = ahi + 2 * A16_BYTES, blo = bhi + B16_BYTES;
#pragma unroll
              for (int k = 0; k < BK / 16; ++k) {  // 32 bytes of K per bf16 MMA inside the 64 B atom
                tc_mma_bf16<2>(tmem_d, make_desc_sw64(alo + k * 32), make_desc_sw64(bhi + k * 32),
                               kIdB, 1u);
                tc_mma_bf16<2>(tmem_d, make_desc_sw64(ahi + k * 32), make_desc_sw64(blo + k * 32),
                               kIdB, 1u);
              }
              tc_commit_pair(empty_bar(slot), 3);
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
              tc_mma_tf32<2>(tmem_d, make_desc_sw128(a1 + k * 32), make_desc_sw128(b0 + k * 32),
                             kIdT, (kb | k) ? 1u : 0u);                           // lo * hi
              tc_mma_tf32<2>(tmem_d, make_desc_sw128(a0 + k * 32), make_desc_sw128(b1 + k * 32),
                             kIdT, 1u);                                            // hi * lo
              tc_mma_tf32<2>(tmem_d, make_desc_sw128(a0 + k * 32), make_desc_sw128(b0 + k * 32),
                             kIdT, 1u);                                            // hi * hi
            }
            tc_commit_pair(empty_bar(slot_hi), 3);
            tc_commit_pair(empty_bar(slot), 3);
            advance();
          }
        }
        tc_commit_pair(tfull_bar(as), 3);  // accumulator complete, both CTAs
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue warps (both CTAs) =====================
    const int quarter = warp;  // warps 0-3: TMEM lanes [32*quarter, 32*quarter+32)
    const uint32_t stg_base = smem_u32(smem + kSlots * kSlotBytes) + static_cast<uint32_t>(warp * 8192);
    uint32_t nstore = 0;
    // kCount: thresholds and bucket counters of this warp, [32][32] words each, in its staging
    // memory; the addresses are this lane's column
    const uint32_t sT = stg_base + static_cast<uint32_t>(lane * 4);
    const uint32_t sC = sT + 32 * 32 * 4;
    if (EPI == kCount) {
      // band tiles: their distances were written by the kBand launch; counted here, while the
      // first accumulator of this pair is still being computed
      const float alpha = 0.f, beta = 0.f;
      for (int b = pair; b < p.f.num_band; b += num_pairs) {
        const uint32_t mn = __ldg(p.f.band + b);
        const int m = static_cast<int>(mn >> 16), n = static_cast<int>(mn & 0xFFFFu);
        const int64_t row0 = static_cast<int64_t>(m) * PM + rank * UM + quarter * 32;
        if (row0 >= p.Q) continue;
        const int64_t r = row0 + lane;
        CountRow cr;
        cr.row_ok = r < p.Q;
        const int32_t q = cr.row_ok ? __ldg(p.f.qorder + r) : 0;
        cr.nvq = cr.row_ok ? __ldg(p.f.nv + q) : 0;
        cr.o = __ldg(p.f.off + q);
        const int wmax = __reduce_max_sync(0xffffffffu, cr.nvq);
        const float *srow = p.f.scratch + (static_cast<int64_t>(b) * PM + rank * UM + quarter * 32 + lane) * BN;
        count_tile<0, 1>(p, cr, wmax, 0u, srow, static_cast<int64_t>(n) * BN, alpha, beta, 0.f, sT, sC, lane);
      }
    }
    int it = 0;
    for (int t = pair; t < num_tiles; t += num_pairs, ++it) {
      int m, n;
      tile_mn(t, m, n);
      const int as = it & 1;
      if (EPI == kAccum && p.acc_mode >= 2 && t + num_pairs < num_tiles) {
        // mean fusion: this thread's row of the NEXT tile's running sums into L2, a whole tile ahead
        int m2, n2;
        tile_mn(t + num_pairs, m2, n2);
        const int64_t r2 = static_cast<int64_t>(m2) * PM + rank * UM + quarter * 32 + lane;
        const int64_t c2 = static_cast<int64_t>(n2) * BN;
        if (r2 < p.Q) {
          const float *src = p.acc + r2 * p.ld_acc + c2;
#pragma unroll
          for (int j = 0; j < BN / 32; ++j)
            if (c2 + 32 * j < p.G) asm volatile("prefetch.global.L2 [%0];" ::"l"(src + 32 * j));
        }
      }
      mbar_wait(tfull_bar(as), (it >> 1) & 1u);
      tc_fence_after();
      const int64_t colt = static_cast<int64_t>(n) * BN;
      const int64_t row0 = static_cast<int64_t>(m) * PM + rank * UM + quarter * 32;
      const uint32_t tbase = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                             static_cast<uint32_t>(as * BN);
      if (row0 < p.Q) {  // otherwise this warp's rows are padding (warp-uniform)
        if (EPI == kCount) {
          const int64_t r = row0 + lane;
          CountRow cr;
          cr.row_ok = r < p.Q;
          const int32_t q = cr.row_ok ? __ldg(p.f.qorder + r) : 0;
          cr.nvq = cr.row_ok ? __ldg(p.f.nv + q) : 0;
          cr.o = __ldg(p.f.off + q);
          const int wmax = __reduce_max_sync(0xffffffffu, cr.nvq);
          const float qs = (p.qsq && cr.row_ok) ? __ldg(p.qsq + r) : 0.f;
          const float alpha = (p.metric == DALI_METRIC_COSINE ? -1.0f : 1.0f) * p.acc_scale;
          const float beta = p.metric == DALI_METRIC_COSINE ? 1.0f : 0.0f;
          if (p.metric == DALI_METRIC_SQEUCLIDEAN)
            count_tile<1, 0>(p, cr, wmax, tbase, nullptr, colt, alpha, beta, qs, sT, sC, lane);
          else if (p.metric == DALI_METRIC_EUCLIDEAN)
            count_tile<2, 0>(p, cr, wmax, tbase, nullptr, colt, alpha, beta, qs, sT, sC, lane);
          else
            count_tile<0, 0>(p, cr, wmax, tbase, nullptr, colt, alpha, beta, qs, sT, sC, lane);
        } else if (EPI == kStore || EPI == kBand || EPI == kAccum) {
          // one thread per query row: metric in registers, then either
          //   - the 32 x 32 block goes to a 128B-swizzled staging tile (conflict-free 16-byte
          //     stores) and leaves through one TMA store (rows / columns beyond Q / G clipped by
          //     the tensor map); two staging tiles per warp, so a store drains while the next
          //     block is prepared; or
          //   - (rows of `out` not 16-byte aligned) a padded transpose through the same staging
          //     memory and coalesced 128-byte row stores.
          const int64_t r_own = row0 + lane;
          const float qs = (p.qsq && r_own < p.Q) ? __ldg(p.qsq + r_own) : 0.f;
          const float alpha = (p.metric == DALI_METRIC_COSINE ? -1.0f : 1.0f) * p.acc_scale;
          const float beta = p.metric == DALI_METRIC_COSINE ? 1.0f : 0.0f;
          const int kind = p.metric == DALI_METRIC_SQEUCLIDEAN ? 1 : p.metric == DALI_METRIC_EUCLIDEAN ? 2 : 0;
          // mean fusion: 32 columns of this row of the running sum (rows are 16-byte aligned)
          float nxt[32];
          const bool acc32 = (p.ld_acc & 7) == 0 && (reinterpret_cast<uintptr_t>(p.acc) & 31) == 0;
          const float acc_rdiv = __frcp_rn(p.acc_div);
          auto load_prev = [&](int64_t col, float (&dst)[32]) {
            const float *src = p.acc + (r_own < p.Q ? r_own : 0) * p.ld_acc + col;
            if (p.G - col >= 32 && acc32) {  // whole 32-byte sectors per request
#pragma unroll
              for (int j = 0; j < 4; ++j)
                asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                             : "=f"(dst[8 * j]), "=f"(dst[8 * j + 1]), "=f"(dst[8 * j + 2]), "=f"(dst[8 * j + 3]),
                               "=f"(dst[8 * j + 4]), "=f"(dst[8 * j + 5]), "=f"(dst[8 * j + 6]), "=f"(dst[8 * j + 7])
                             : "l"(src + 8 * j) : "memory");
            } else if (p.G - col >= 32) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 x = *reinterpret_cast<const float4 *>(src + 4 * j);
                dst[4 * j] = x.x; dst[4 * j + 1] = x.y; dst[4 * j + 2] = x.z; dst[4 * j + 3] = x.w;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) dst[j] = col + j < p.G ? src[j] : 0.f;
            }
          };
          // kBand: this row's identity segment [seg_lo, seg_lo + seg_m) in sorted-gallery columns
          int64_t seg_lo = 0, seg_o = 0;
          int seg_m = 0;
          if (EPI == kBand && r_own < p.Q) {
            const int32_t q = __ldg(p.f.qorder + r_own);
            seg_o = __ldg(p.f.off + q);
            seg_m = static_cast<int>(__ldg(p.f.off + q + 1) - seg_o);
            seg_lo = __ldg(p.f.lo + q);
          }
          if (EPI == kAccum && p.acc_mode >= 2 && colt < p.G) load_prev(colt, nxt);
#pragma unroll 1
          for (int c = 0; c < BN / 32; ++c) {
            const int64_t col0 = colt + c * 32;
            if (col0 >= p.G) break;
            uint32_t v[32];
            const int lim = p.G - col0 < 32 ? static_cast<int>(p.G - col0) : 32;
            // mean fusion: the running sum of this row's 32 columns was requested one block ago;
            // the next block's is requested now, a block of epilogue work ahead of its use
            float prev[32];
            if (EPI == kAccum && p.acc_mode >= 2) {
#pragma unroll
              for (int j = 0; j < 32; ++j) prev[j] = nxt[j];
              if (c + 1 < BN / 32 && col0 + 32 < p.G) load_prev(col0 + 32, nxt);
            }
            tc_ld_32x32(tbase + c * 32, v);
            tc_wait_ld();
            const float *gs = p.gsq ? p.gsq + col0 : nullptr;
            if (kind == 0) apply_metric<0>(v, alpha, beta, p.acc_scale, qs, gs, lim);
            else if (kind == 1) apply_metric<1>(v, alpha, beta, p.acc_scale, qs, gs, lim);
            else apply_metric<2>(v, alpha, beta, p.acc_scale, qs, gs, lim);
            if (EPI == kBand) {
              // order keys of the matches among these 32 columns, to their slots of the match list
              const int64_t rel0 = col0 - seg_lo;
              if (rel0 < seg_m && rel0 + 32 > 0) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                  const int64_t rel = rel0 + j;
                  if (rel >= 0 && rel < seg_m && j < lim)
                    p.f.keys[seg_o + __ldg(p.f.slot_of_seg + seg_o + rel)] = dist_key(__uint_as_float(v[j]));
                }
              }
              __syncwarp();
            }
#ifdef DALI_UMMA_NO_STORE  // probe build (tests/probes/build_variants.sh): the tile is computed, not stored
            if (v[0] != 0x7fc12345u) continue;
#endif
            // What the stores cost (round 2, tests/probes/contraction_probe.py): without them the
            // launch is 21-24 us shorter at D = 768 (0.168 -> 0.147 ms) and at D = 2048 (0.418 ->
            // 0.394 ms) alike -- the staging writes and the TMA store's reads add 10 % / 4 % to the
            // shared-memory traffic of a tile, which the operand reads of the MMAs already saturate.
            // Not the L2 (an evict-first policy on the stores changed nothing), not bursts (pacing the
            // eight blocks of a tile over its MMA time: -1 %), and registers -> global memory without
            // staging (32 rows x 128 B per warp instruction) was slower: 0.190 / 0.452 ms.
            if (p.tma_out) {
              if (EPI != kAccum || p.store_out) {
                const uint32_t buf = stg_base + static_cast<uint32_t>(((nstore++) & 1) * 4096);
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                __syncwarp();
                const uint32_t rowaddr = buf + static_cast<uint32_t>(lane * 128);
#pragma unroll
                for (int j = 0; j < 8; ++j)
                  sts128(rowaddr + static_cast<uint32_t>(((j ^ (lane & 7)) << 4)),
                         __uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                         __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                  if (EPI == kBand)  // tile t of the list -> rows [256 t, 256 t + 256) of the scratch
                    tma_store_2d(&tmOut, buf, c * 32, t * PM + static_cast<int>(rank) * UM + quarter * 32);
                  else
                    tma_store_2d(&tmOut, buf, static_cast<int32_t>(col0), static_cast<int32_t>(row0));
                  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
              }
              if (EPI == kAccum) {
                // the running sum in fuse.cu's order and rounding: acc + d, the last launch divides
                if (p.acc_mode >= 2) {
#pragma unroll
                  for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__fadd_rn(prev[j], __uint_as_float(v[j])));
                }
                if (p.acc_mode == 3) {
                  // one range test for the 32 columns, then the branch-free refinement (a test and a
                  // branch per element serialised the block: 5 us per tile at D = 768)
                  bool fast = true;
#pragma unroll
                  for (int j = 0; j < 32; ++j) {
                    const float ax = fabsf(__uint_as_float(v[j]));
                    fast = fast && ax >= 7.888609052210118e-31f && ax < 1.2676506002282294e30f;
                  }
                  if (fast) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                      const float x = __uint_as_float(v[j]);
                      const float q0 = __fmul_rn(x, acc_rdiv);
                      v[j] = __float_as_uint(__fmaf_rn(__fmaf_rn(-q0, p.acc_div, x), acc_rdiv, q0));
                    }
                  } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(div_small_int(__uint_as_float(v[j]), p.acc_div, acc_rdiv));
                  }
                }
                const uint32_t buf = stg_base + static_cast<uint32_t>(((nstore++) & 1) * 4096);
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                __syncwarp();
                const uint32_t rowaddr = buf + static_cast<uint32_t>(lane * 128);
#pragma unroll
                for (int j = 0; j < 8; ++j)
                  sts128(rowaddr + static_cast<uint32_t>(((j ^ (lane & 7)) << 4)),
                         __uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                         __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                  tma_store_2d(&tmAcc, buf, static_cast<int32_t>(col0), static_cast<int32_t>(row0));
                  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                sts32(stg_base + static_cast<uint32_t>((lane * EPI_LD + j) * 4), __uint_as_float(v[j]));
              __syncwarp();
              const int64_t col = col0 + lane;
              float *dst = p.out + row0 * p.ld + col;
              const int rows = p.Q - row0 < 32 ? static_cast<int>(p.Q - row0) : 32;
              if (col < p.G) {
#pragma unroll 8
                for (int rr = 0; rr < 32; ++rr) {
                  const float val = lds32(stg_base + static_cast<uint32_t>((rr * EPI_LD + lane) * 4));
                  if (rr < rows) dst[rr * p.ld] = val;
                }
              }
              __syncwarp();
            }
          }
        } else {
          // one thread per query row: its 256 distances of this tile against the row's
          // running threshold; the survivors (about k ln(G/k) per row over the whole
          // gallery) go to the row's candidate list
          const int64_t r = row0 + lane;
          FilterRow fr;
          fr.row_ok = r < p.Q;
          fr.thr = fr.row_ok ? __ldg(p.thr + r) : 0.f;
          fr.qs = (p.qsq && fr.row_ok) ? __ldg(p.qsq + r) : 0.f;
          fr.scale = p.acc_scale;
          fr.alpha = (p.metric == DALI_METRIC_COSINE ? -1.0f : 1.0f) * p.acc_scale;
          fr.beta = p.metric == DALI_METRIC_COSINE ? 1.0f : 0.0f;
          fr.flip = p.largest ? 0xFFFFFFFFu : 0u;
          fr.list = p.cand + (fr.row_ok ? r : 0) * p.cap;
          fr.cnt = p.cand_cnt + (fr.row_ok ? r : 0);
          fr.cap = p.cap;
          const int kind = p.metric == DALI_METRIC_SQEUCLIDEAN ? 1 : p.metric == DALI_METRIC_EUCLIDEAN ? 2 : 0;
          const int sel = p.direct ? 2 : (p.largest ? 1 : 0);
          const int variant = kind * 3 + sel;  // warp-uniform
#pragma unroll 1
          for (int c = 0; c < BN / 32; ++c) {
            const int64_t col0 = colt + c * 32;
            if (col0 >= p.G) break;
            uint32_t v[32];
            tc_ld_32x32(tbase + c * 32, v);
            tc_wait_ld();
            const int lim = p.G - col0 < 32 ? static_cast<int>(p.G - col0) : 32;
            const float *gs = p.gsq ? p.gsq + col0 : nullptr;
            const uint32_t id0 = static_cast<uint32_t>(p.id_base + col0);
            const int slot0 = static_cast<int>(col0);
            switch (variant) {
              case 0: filter_cols<0, 0>(v, fr, gs, lim, id0, slot0); break;
              case 1: filter_cols<0, 1>(v, fr, gs, lim, id0, slot0); break;
              case 2: filter_cols<0, 2>(v, fr, gs, lim, id0, slot0); break;
              case 3: filter_cols<1, 0>(v, fr, gs, lim, id0, slot0); break;
              case 4: filter_cols<1, 1>(v, fr, gs, lim, id0, slot0); break;
              case 5: filter_cols<1, 2>(v, fr, gs, lim, id0, slot0); break;
              case 6: filter_cols<2, 0>(v, fr, gs, lim, id0, slot0); break;
              case 7: filter_cols<2, 1>(v, fr, gs, lim, id0, slot0); break;
              default: filter_cols<2, 2>(v, fr, gs, lim, id0, slot0); break;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(tempty_bar(as), 0);
    }
    if ((EPI == kStore || EPI == kBand || EPI == kAccum) && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  tc_fence_before();
  cluster_sync_all();  // both CTAs are done with TMEM and with each other's barriers
  if (warp == kMmaWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem_base)
                 : "memory");
  }
}

template <int MODE, int EPI>
int launch_t(dali_ctx *ctx, const CUtensorMap &tmA, const CUtensorMap &tmB, const CUtensorMap &tmA16,
             const CUtensorMap &tmB16, const CUtensorMap &tmOut, const Umma2Params &p,
             const CUtensorMap *tmAcc = nullptr) {
  if (int rc = ensure_dyn_smem(ctx, reinterpret_cast<const void *>(&distmat_umma2_kernel<MODE, EPI>), kSmemBytes))
    return rc;
  const int tiles = p.f.tiles ? std::max(p.f.num_list, p.f.num_band) : p.num_m_pairs * p.num_n_tiles;
  if (tiles <= 0) return DALI_OK;
  const int max_pairs = ctx->num_sms / 2;
  const int pairs = tiles < max_pairs ? tiles : max_pairs;
  KTimer t(ctx, DALI_K_DISTMAT);
  distmat_umma2_kernel<MODE, EPI><<<2 * pairs, kThreads, kSmemBytes, ctx->stream>>>(tmA, tmB, tmA16,
                                                                                   tmB16, tmOut,
                                                                                   tmAcc ? *tmAcc : tmOut, p);
  DALI_CUDA_OK(ctx, cudaGetLastError());
  return DALI_OK;
}

template <int EPI>
int launch_prec(dali_ctx *ctx, int precision, const CUtensorMap &tmA, const CUtensorMap &tmB,
                const CUtensorMap &tmA16, const CUtensorMap &tmB16, const CUtensorMap &tmOut,
                const Umma2Params &p, const CUtensorMap *tmAcc = nullptr) {
  switch (precision) {
    case DALI_PREC_TF32: return launch_t<kTf32, EPI>(ctx, tmA, tmB, tmA16, tmB16, tmOut, p, tmAcc);
    case DALI_PREC_TF32X3: return launch_t<kTf32x3, EPI>(ctx, tmA, tmB, tmA16, tmB16, tmOut, p, tmAcc);
    case DALI_PREC_TF32C: return launch_t<kTf32c, EPI>(ctx, tmA, tmB, tmA16, tmB16, tmOut, p, tmAcc);
    case DALI_PREC_F16X3: return launch_t<kF16x3, EPI>(ctx, tmA, tmB, tmA16, tmB16, tmOut, p, tmAcc);
    case DALI_PREC_F16: return launch_t<kF16, EPI>(ctx, tmA, tmB, tmA16, tmB16, tmOut, p, tmAcc);
    default: return set_err(ctx, DALI_ERR_INVALID, "not a tensor-core precision");
  }
}

// F16X3 arithmetic: operand format, MMA schedule and truncation compensation.
//
// What the tensor core does (measured on B200, tests/probes/trunc_probe.py): the 16 products of a
// kind::f16 MMA are aligned to the exponent of the fp32 accumulator and TRUNCATED (toward zero)
// about two bits below its ulp, and the sum is truncated to fp32 again.  With plain fp16 hi
// planes (22-bit products) every hi*hi MMA adding into a running sum s lost ~2.2 ulp(s), every
// correction MMA ~0.3 ulp(s): invisible for ordinary pairs (|s| << 1, error ~1e-7), but a
// near-duplicate pair (s -> 1) came out low by 1.5e-5 at D = 2048 and 2.8e-5 at D = 3840 --
// beyond the 1e-5 parity bar.  Three measures, cheapest first:
//   1. fixed-point hi plane (normalize.cu, hi_quant): hi = multiple of 0.5 of the 2^12-scaled
//      value, so every hi*hi product is a multiple of 0.25 -- nothing is lost at the alignment
//      step, and only the fractional part of the sum (0.22 ulp on average) at the final one.
//      The residual plane then holds |lo| <= 0.25 and the dropped lo*lo term is D/48 * 2^-24 for
//      an exact duplicate (2.5e-6 at D = 2048), ~0 otherwise.
//   2. the mean loss that remains is given back: the accumulator is multiplied by
//      1 + kappa * N * 2^-24 (N = MMAs issued on the large accumulator, kappa fitted: 0.24 per MMA
//      interleaved, 0.22 corrections-first), folded into the 2^-24 that undoes the operand scaling
//      -- no extra instruction, pairs with s ~ 0 unchanged.
//   3. D > 2048: all correction MMAs of a tile first, hi*hi last (two_pass): the 2 D / 16 correction
//      instructions then run on a small accumulator and lose nothing.  Costs 1.5x the operand
//      traffic (13 % of the kernel at D = 2048), hence only where 1 + 2 are not enough.
// Measured after 1-3 (float64 reference): exact duplicates <= 4.5e-6 at D = 2048 and <= 5.5e-6 at
// D = 3840 / 4096; adversarial partial-sum trajectories (all energy in the first / last k-block)
// <= 8e-6; ordinary pairs <= 1e-6.  DALI_F16X3_TWO_PASS=0/1, DALI_F16X3_COMP=<kappa> (0 = off)
// and DALI_F16X3_GRID=<grid> (0 = plain fp16) override for calibration runs.
void f16x3_schedule(int64_t Dp, int *two_pass, float *acc_scale) {
  static const char *env_tp = getenv("DALI_F16X3_TWO_PASS");
  static const char *env_k = getenv("DALI_F16X3_COMP");
  const bool tp = env_tp ? atoi(env_tp) != 0 : Dp > 2048;
  double kappa = tp ? kKappaTwoPass : kKappaInterleaved;
  if (f16x3_hi_grid(Dp) != 0.5f) kappa = 0.0;  // the constants belong to the default operand format
  if (env_k) kappa = atof(env_k);
  const double n = tp ? static_cast<double>(Dp) / 16.0 : 3.0 * static_cast<double>(Dp) / 16.0;
  *two_pass = tp ? 1 : 0;
  *acc_scale = static_cast<float>(5.9604644775390625e-08 * (1.0 + kappa * n * 5.9604644775390625e-08));
}

// Output map: fp32 [Q][G] with row pitch ld, 32 x 32 boxes in the 128-byte swizzle the epilogue
// writes its staging tiles in.
int make_out_map(dali_ctx *ctx, CUtensorMap *map, float *out, int64_t Q, int64_t G, int64_t ld) {
  // ctx->encode_tiled was resolved by make_map() in setup()
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(G), static_cast<cuuint64_t>(Q)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 4};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = reinterpret_cast<EncodeTiledFn>(ctx->encode_tiled)(
      map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, out, dims, strides, box, estr,
      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_err(ctx, DALI_ERR_CUDA, "cuTensorMapEncodeTiled (output) failed with CUresult " + std::to_string(r));
  return DALI_OK;
}

int setup(dali_ctx *ctx, const float *q32, const float *g32, const void *q16, const void *g16,
          int64_t Q, int64_t G, int64_t Dp, int64_t q_rows_pad, int64_t g_rows_pad, int64_t g_row0,
          int precision, int metric, const float *qsq, const float *gsq, CUtensorMap *tmA,
          CUtensorMap *tmB, CUtensorMap *tmA16, CUtensorMap *tmB16, Umma2Params *p) {
  const int npl32 = precision == DALI_PREC_TF32X3 ? 2 : 1;
  if (Dp % BK != 0 || q_rows_pad % PM != 0 || g_rows_pad % BN != 0 || g_row0 % BN != 0)
    return set_err(ctx, DALI_ERR_INVALID, "umma operands must be padded (rows 256, D 32)");
  if (q_rows_pad * 2 > INT32_MAX || g_rows_pad * 2 > INT32_MAX)
    return set_err(ctx, DALI_ERR_UNSUPPORTED, "operand too tall for one tensor map");
  const bool f16 = precision == DALI_PREC_F16X3 || precision == DALI_PREC_F16;
  int rc = DALI_OK;
  if (!f16) {
    rc = make_map(ctx, tmA, q32, q_rows_pad * npl32, Dp, UM, false);
    if (rc) return rc;
    rc = make_map(ctx, tmB, g32, g_rows_pad * npl32, Dp, HB, false);
    if (rc) return rc;
  }
  if (precision == DALI_PREC_TF32C || f16) {
    rc = make_map(ctx, tmA16, q16, q_rows_pad * 2, Dp, UM, true, f16);
    if (rc) return rc;
    rc = make_map(ctx, tmB16, g16, g_rows_pad * 2, Dp, HB, true, f16);
    if (rc) return rc;
    if (f16) {
      *tmA = *tmA16;
      *tmB = *tmB16;
    }
  } else {
    *tmA16 = *tmA;
    *tmB16 = *tmB;
  }
  *p = Umma2Params{};
  p->Q = Q; p->G = G;
  p->num_m_pairs = static_cast<int>((Q + PM - 1) / PM);
  p->num_n_tiles = static_cast<int>((G + BN - 1) / BN);
  p->num_kb = static_cast<int>(Dp / BK);
  p->a_plane_rows = static_cast<int32_t>(q_rows_pad);
  p->b_plane_rows = static_cast<int32_t>(g_rows_pad);
  p->b_row0 = static_cast<int32_t>(g_row0);
  p->metric = metric; p->qsq = qsq; p->gsq = gsq;
  p->acc_scale = f16 ? 5.9604644775390625e-08f /* 2^-24 */ : 1.0f;
  if (precision == DALI_PREC_F16X3) f16x3_schedule(Dp, &p->two_pass, &p->acc_scale);
  {
    // bytes of operand planes one tile row of 256 rows reads per pass over K
    const int64_t bytes_per_row = Dp * (precision == DALI_PREC_F16 ? 2 : precision == DALI_PREC_TF32 ? 4 : precision == DALI_PREC_F16X3 ? 4 : 8);
    const int64_t a_bytes = static_cast<int64_t>(p->num_m_pairs) * PM * bytes_per_row;
    static const char *env_band = getenv("DALI_UMMA_NBAND");
    int nband = 1;
    if (a_bytes > (48ll << 20))
      nband = static_cast<int>(std::max<int64_t>(1, (24ll << 20) / (BN * bytes_per_row)));
    if (env_band) nband = atoi(env_band);
    p->nband = std::max(1, std::min(nband, p->num_n_tiles));
  }
  if (static_cast<int64_t>(p->num_m_pairs) * p->num_n_tiles > INT32_MAX)
    return set_err(ctx, DALI_ERR_UNSUPPORTED, "too many tiles (chunk the queries)");
  return DALI_OK;
}

}  // namespace

// q32/g32: [npl32][rows_pad][Dp] fp32 planes (plane 0 = TF32-rounded operand, plane 1 = residual,
// TF32X3 only); q16/g16: [2][rows_pad][Dp] bf16 planes (hi16, lo16; TF32C only).
// rows_pad multiples of 256, Dp multiple of 32.  mode: DALI_PREC_TF32 / TF32X3 / TF32C.
// DALI_UMMA_2CTA=0 selects the one-CTA-per-tile kernel of distmat_umma.cu (cross-check).
int launch_distmat_umma(dali_ctx *ctx, const float *q32, const float *g32, const void *q16,
                        const void *g16, int64_t Q, int64_t G, int64_t Dp, int64_t q_rows_pad,
                        int64_t g_rows_pad, int64_t g_row0, int precision, int metric,
                        const float *qsq, const float *gsq, float *out, int64_t ld, float *acc,
                        int64_t ld_acc, int acc_mode, float acc_div) {
  if (Q == 0 || G == 0) return DALI_OK;
  static const char *env = getenv("DALI_UMMA_2CTA");
  if (env && atoi(env) == 0 && precision != DALI_PREC_F16X3 && precision != DALI_PREC_F16 && !acc_mode)
    return launch_distmat_umma1(ctx, q32, g32, q16, g16, Q, G, Dp, q_rows_pad, g_rows_pad, g_row0,
                                precision, metric, qsq, gsq, out, ld);
  CUtensorMap tmA, tmB, tmA16, tmB16;
  Umma2Params p;
  int rc = setup(ctx, q32, g32, q16, g16, Q, G, Dp, q_rows_pad, g_rows_pad, g_row0, precision, metric,
                 qsq, gsq, &tmA, &tmB, &tmA16, &tmB16, &p);
  if (rc) return rc;
  p.out = out; p.ld = ld;
  // TMA stores need 16-byte aligned rows (always true for the library's own matrices, whose
  // leading dimension is a multiple of 4; a caller's contiguous [Q, G] with odd G is not)
  static const char *env_tma = getenv("DALI_UMMA_TMA_STORE");
  CUtensorMap tmOut = tmA;
  p.tma_out = out && (ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0) &&
              !(env_tma && atoi(env_tma) == 0) && Q <= INT32_MAX && G <= INT32_MAX;
  if (p.tma_out) {
    rc = make_out_map(ctx, &tmOut, out, Q, G, ld);
    if (rc) return rc;
  }
  p.store_out = 1;
  if (acc_mode) {
    // the running sum leaves (and its previous value is read) in 16-byte pieces
    if (!acc || ld_acc % 4 != 0 || (reinterpret_cast<uintptr_t>(acc) & 15) != 0 || Q > INT32_MAX || G > INT32_MAX)
      return set_err(ctx, DALI_ERR_UNSUPPORTED, "fused mean needs a sum matrix with 16-byte aligned rows");
    CUtensorMap tmAcc;
    rc = make_out_map(ctx, &tmAcc, acc, Q, G, ld_acc);
    if (rc) return rc;
    p.acc = acc; p.ld_acc = ld_acc; p.acc_mode = acc_mode; p.acc_div = acc_div;
    p.store_out = out != nullptr && p.tma_out;
    if (out && !p.tma_out)
      return set_err(ctx, DALI_ERR_UNSUPPORTED, "fused mean with an individual matrix needs 16-byte aligned rows there too");
    p.tma_out = 1;
    return launch_prec<kAccum>(ctx, precision, tmA, tmB, tmA16, tmB16, p.store_out ? tmOut : tmAcc, p, &tmAcc);
  }
  return launch_prec<kStore>(ctx, precision, tmA, tmB, tmA16, tmB16, tmOut, p);
}

// Fused distance + top-k candidate filter: columns [0, G) of the slab starting at gallery row
// g_row0 are compared with thr[Q]; survivors are appended to cand[Q][cap] (cand_cnt[Q] counts
// every survivor, also those beyond cap: the caller detects overflow).  direct != 0 requires
// G <= cap and writes column j to slot j without counting.
int launch_distmat_filter_umma(dali_ctx *ctx, const float *q32, const float *g32, const void *q16,
                               const void *g16, int64_t Q, int64_t G, int64_t Dp,
                               int64_t q_rows_pad, int64_t g_rows_pad, int64_t g_row0,
                               int precision, int metric, const float *qsq, const float *gsq,
                               const float *thr, int32_t *cand_cnt, uint64_t *cand, int cap,
                               int largest, int direct, int32_t id_base) {
  if (Q == 0 || G == 0) return DALI_OK;
  if (direct && G > cap) return set_err(ctx, DALI_ERR_INVALID, "direct filter chunk wider than cap");
  CUtensorMap tmA, tmB, tmA16, tmB16;
  Umma2Params p;
  int rc = setup(ctx, q32, g32, q16, g16, Q, G, Dp, q_rows_pad, g_rows_pad, g_row0, precision, metric,
                 qsq, gsq, &tmA, &tmB, &tmA16, &tmB16, &p);
  if (rc) return rc;
  p.thr = thr; p.cand_cnt = cand_cnt; p.cand = cand; p.cap = cap;
  p.largest = largest; p.direct = direct; p.id_base = id_base;
  return launch_prec<kFilter>(ctx, precision, tmA, tmB, tmA16, tmB16, tmA, p);
}

// ---- fused distance + positive-rank counting ---------------------------------------------------
// Both launches run on identity-sorted operand planes (see FusedArgs in common.cuh).
template <int EPI>
static int launch_fused_prec(dali_ctx *ctx, int precision, const CUtensorMap &tmA, const CUtensorMap &tmB,
                             const CUtensorMap &tmA16, const CUtensorMap &tmB16, const CUtensorMap &tmOut,
                             const Umma2Params &p) {
  switch (precision) {
    case DALI_PREC_TF32: return launch_t<kTf32, EPI>(ctx, tmA, tmB, tmA16, tmB16, tmOut, p);
    case DALI_PREC_TF32C: return launch_t<kTf32c, EPI>(ctx, tmA, tmB, tmA16, tmB16, tmOut, p);
    case DALI_PREC_F16X3: return launch_t<kF16x3, EPI>(ctx, tmA, tmB, tmA16, tmB16, tmOut, p);
    case DALI_PREC_F16: return launch_t<kF16, EPI>(ctx, tmA, tmB, tmA16, tmB16, tmOut, p);
    default: return set_err(ctx, DALI_ERR_UNSUPPORTED, "fused counting: precision not supported");
  }
}

bool fused_count_supports(int precision) {
  return precision == DALI_PREC_TF32 || precision == DALI_PREC_TF32C || precision == DALI_PREC_F16X3 ||
         precision == DALI_PREC_F16;
}

// kBand: fa.tiles = the band tiles; their distances -> fa.scratch, the matches' keys -> fa.keys.
int launch_distmat_band_umma(dali_ctx *ctx, const float *q32, const float *g32, const void *q16,
                             const void *g16, int64_t Q, int64_t G, int64_t Dp, int64_t q_rows_pad,
                             int64_t g_rows_pad, int64_t g_row0, int precision, int metric,
                             const float *qsq, const float *gsq, const FusedArgs &fa) {
  if (Q == 0 || G == 0 || fa.num_list == 0) return DALI_OK;
  CUtensorMap tmA, tmB, tmA16, tmB16, tmOut;
  Umma2Params p;
  int rc = setup(ctx, q32, g32, q16, g16, Q, G, Dp, q_rows_pad, g_rows_pad, g_row0, precision, metric,
                 qsq, gsq, &tmA, &tmB, &tmA16, &tmB16, &p);
  if (rc) return rc;
  p.f = fa;
  p.tma_out = 1;
  rc = make_out_map(ctx, &tmOut, fa.scratch, static_cast<int64_t>(fa.num_list) * PM, BN, BN);
  if (rc) return rc;
  return launch_fused_prec<kBand>(ctx, precision, tmA, tmB, tmA16, tmB16, tmOut, p);
}

// kCount: fa.tiles = every tile that is not a band tile (computed and counted from TMEM), fa.band =
// the band tiles (counted from fa.scratch); fa.counts must be zero on entry.
int launch_distmat_count_umma(dali_ctx *ctx, const float *q32, const float *g32, const void *q16,
                              const void *g16, int64_t Q, int64_t G, int64_t Dp, int64_t q_rows_pad,
                              int64_t g_rows_pad, int64_t g_row0, int precision, int metric,
                              const float *qsq, const float *gsq, const FusedArgs &fa) {
  if (Q == 0 || G == 0) return DALI_OK;
  CUtensorMap tmA, tmB, tmA16, tmB16;
  Umma2Params p;
  int rc = setup(ctx, q32, g32, q16, g16, Q, G, Dp, q_rows_pad, g_rows_pad, g_row0, precision, metric,
                 qsq, gsq, &tmA, &tmB, &tmA16, &tmB16, &p);
  if (rc) return rc;
  p.f = fa;
  static const char *env_dbg = getenv("DALI_FUSED_DBG");
  if (env_dbg) p.f.dbg = atoi(env_dbg);
  return launch_fused_prec<kCount>(ctx, precision, tmA, tmB, tmA16, tmB16, tmA, p);
}

}  // namespace dali
