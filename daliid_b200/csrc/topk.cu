// Per-row top-k selection without a full sort.
//
// replaces  torch.argsort(distmat, dim=1)[:, :20]          validateModels.py:93
//           torch.topk(S, k=5, dim=1, largest=True)        validateModels.py:180
//           torch.argsort(sim, descending=True)[:topK]     getFeatures.py:303 (prefix only)
// Order: value ascending (descending when `largest`), ties by ascending id, NaN last
// (first when `largest`, as torch.topk treats NaN as the largest value).
//
// One CTA streams one row (4 B / pair).  A running threshold (the current k-th best
// composite key) filters the stream; survivors are appended to a shared-memory candidate
// list which is bitonic-sorted and pruned back to k whenever it could overflow.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace dali {

namespace {

constexpr int kTopkThreads = 256;
constexpr int kCap = 2048;           // candidate capacity (uint64 composites, 16 KiB)
constexpr int kTile = kTopkThreads * 4;

__device__ __forceinline__ float4 ld_stream_f4(const float *p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}

// sort cand[0..n) ascending, n a power of two (entries >= cnt hold UINT64_MAX)
__device__ void bitonic_sort(uint64_t *cand, int n) {
  for (int size = 2; size <= n; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int i = threadIdx.x; i < (n >> 1); i += blockDim.x) {
        const int pos = 2 * i - (i & (stride - 1));
        const bool asc = (pos & size) == 0;
        const uint64_t a = cand[pos], b = cand[pos + stride];
        if ((a > b) == asc) {
          cand[pos] = b;
          cand[pos + stride] = a;
        }
      }
    }
  }
  __syncthreads();
}

__device__ __forceinline__ int next_pow2(int v) {
  int n = 2;
  while (n < v) n <<= 1;
  return n;
}

struct TopkShared {
  uint64_t cand[kCap];
  uint64_t thr;
  int cnt;
};

__device__ void prune(TopkShared &s, int k) {
  __syncthreads();
  const int cnt = s.cnt;
  const int n = next_pow2(cnt);
  for (int i = cnt + threadIdx.x; i < n; i += kTopkThreads) s.cand[i] = ~0ull;
  bitonic_sort(s.cand, n);
  if (threadIdx.x == 0) {
    if (cnt >= k) {
      s.cnt = k;
      s.thr = s.cand[k - 1];
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kTopkThreads)
topk_kernel(const float *__restrict__ dist, int64_t G, int64_t ld, int k, int largest,
            const int32_t *__restrict__ col_ids, int32_t id_base, float *__restrict__ d_out,
            int32_t *__restrict__ i_out, const int32_t *__restrict__ only_rows) {
  __shared__ TopkShared s;
  const int64_t q = blockIdx.x;
  if (only_rows && only_rows[q] == 0) return;  // repair pass: rows the streaming path completed
  const float *row = dist + q * ld;
  const int32_t *ids = col_ids ? col_ids + q * ld : nullptr;
  const int tid = threadIdx.x;
  const uint32_t flip = largest ? 0xFFFFFFFFu : 0u;
  if (tid == 0) {
    s.cnt = 0;
    s.thr = ~0ull;
  }
  __syncthreads();

  auto offer = [&](float d, int64_t col) {
    const uint32_t id = static_cast<uint32_t>(ids ? __ldg(ids + col) : id_base + static_cast<int32_t>(col));
    const uint64_t c = composite(dist_key(d) ^ flip, id);
    if (c < s.thr) {
      const int pos = atomicAdd(&s.cnt, 1);
      s.cand[pos] = c;
    }
  };

  // warm-up: the first 256 columns set an initial threshold cheaply
  const int64_t warm = G < kTopkThreads ? G : kTopkThreads;
  if (tid < warm) offer(__ldg(row + tid), tid);
  prune(s, k);

  // main stream: scalar head to reach 16-byte alignment, then float4 tiles
  int64_t c0 = warm;
  const int mis = static_cast<int>((reinterpret_cast<uintptr_t>(row + c0) >> 2) & 3);
  int64_t head = (4 - mis) & 3;
  if (head > G - c0) head = G - c0;
  if (tid < head) offer(__ldg(row + c0 + tid), c0 + tid);
  c0 += head;
  int need = 0;
  for (; c0 < G; c0 += kTile) {
    // the list can grow by at most kTile per iteration; `need` is block-uniform
    if (need) prune(s, k);
    const int64_t c = c0 + 4 * tid;
    if (c + 3 < G) {
      const float4 x = ld_stream_f4(row + c);
      offer(x.x, c); offer(x.y, c + 1); offer(x.z, c + 2); offer(x.w, c + 3);
    } else {
      for (int64_t e = c; e < G; ++e) offer(__ldg(row + e), e);
    }
    // barrier + OR: the last thread to arrive has seen every append of this iteration
    need = __syncthreads_or(s.cnt > kCap - kTile - 4);
  }
  prune(s, k);

  const int cnt = s.cnt < k ? s.cnt : k;
  for (int i = tid; i < k; i += kTopkThreads) {
    float dv;
    int32_t iv;
    if (i < cnt) {
      const uint64_t c = s.cand[i];
      dv = key_to_dist(static_cast<uint32_t>(c >> 32) ^ flip);
      iv = static_cast<int32_t>(static_cast<uint32_t>(c));
    } else {
      dv = largest ? -INFINITY : INFINITY;
      iv = -1;
    }
    d_out[q * k + i] = dv;
    i_out[q * k + i] = iv;
  }
}

// Fused top-k (distmat_umma2.cu, kFilter epilogue): after every gallery chunk the row's candidate
// list [kept best so far | survivors of the chunk] is sorted, cut back to the k best, and the
// row's threshold becomes its k-th best distance.  cnt[row] > cap means the chunk produced more
// survivors than the list holds: the overflow flag tells the host to redo the call unfused.
constexpr int kCompactThreads = 128;
constexpr int kCompactCap = 1024;

// Four rows per CTA.  A row whose list holds <= 256 entries (all but pathological chunks: k kept +
// about f*k survivors) is sorted by ONE warp in its own 2 KiB of shared memory with warp-level
// barriers only; longer lists (up to cap) are then sorted by the whole CTA, one row after the other.
constexpr int kWarpSort = 256;

__device__ __forceinline__ void warp_bitonic(uint64_t *a, int n, int lane) {
  for (int size = 2; size <= n; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncwarp();
      for (int i = lane; i < (n >> 1); i += 32) {
        const int pos = 2 * i - (i & (stride - 1));
        const bool asc = (pos & size) == 0;
        const uint64_t x = a[pos], y = a[pos + stride];
        if ((x > y) == asc) {
          a[pos] = y;
          a[pos + stride] = x;
        }
      }
    }
  }
  __syncwarp();
}

__global__ void __launch_bounds__(kCompactThreads)
topk_compact_kernel(uint64_t *__restrict__ cand, int32_t *__restrict__ cnt, float *__restrict__ thr,
                    int64_t Q, int cap, int k, int largest, int fixed_cnt, int32_t *__restrict__ overflow,
                    float *__restrict__ d_out, int32_t *__restrict__ i_out,
                    int32_t *__restrict__ row_flags) {
  __shared__ uint64_t s_small[4][kWarpSort];
  __shared__ uint64_t s_big[kCompactCap];
  __shared__ int s_n[4];
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  const uint32_t flip = largest ? 0xFFFFFFFFu : 0u;
  const int64_t q_w = static_cast<int64_t>(blockIdx.x) * 4 + w;

  // writes the outcome of one sorted row; `t`/`nt`: the participating thread and their number
  auto emit = [&](int64_t q, const uint64_t *srt, int n, int t, int nt) {
    uint64_t *list = cand + q * cap;
    const int m = n < k ? n : k;
    for (int i = t; i < m; i += nt) list[i] = srt[i];
    if (t == 0) {
      cnt[q] = m;
      thr[q] = n >= k ? key_to_dist(static_cast<uint32_t>(srt[k - 1] >> 32) ^ flip)
                      : (largest ? -INFINITY : INFINITY);
    }
    if (d_out) {
      for (int i = t; i < k; i += nt) {
        float dv = largest ? -INFINITY : INFINITY;
        int32_t iv = -1;
        if (i < m) {
          dv = key_to_dist(static_cast<uint32_t>(srt[i] >> 32) ^ flip);
          iv = static_cast<int32_t>(static_cast<uint32_t>(srt[i]));
        }
        d_out[q * k + i] = dv;
        i_out[q * k + i] = iv;
      }
    }
  };

  int n = 0;
  if (q_w < Q) {
    n = fixed_cnt >= 0 ? fixed_cnt : cnt[q_w];
    if (n > cap) {
      if (lane == 0) {
        atomicOr(overflow, 1);
        if (row_flags) row_flags[q_w] = 1;
      }
      n = cap;
    }
    if (n <= kWarpSort) {
      uint64_t *a = s_small[w];
      const uint64_t *list = cand + q_w * cap;
      const int np2 = next_pow2(n);
      for (int i = lane; i < np2; i += 32) a[i] = i < n ? list[i] : ~0ull;
      warp_bitonic(a, np2, lane);
      emit(q_w, a, n, lane, 32);
      n = 0;  // done
    }
  }
  if (lane == 0) s_n[w] = n;
  __syncthreads();
  for (int r = 0; r < 4; ++r) {  // long lists: the whole CTA, block-uniform
    const int nr = s_n[r];
    if (nr == 0) continue;
    const int64_t q = static_cast<int64_t>(blockIdx.x) * 4 + r;
    const uint64_t *list = cand + q * cap;
    const int np2 = next_pow2(nr);
    for (int i = tid; i < np2; i += kCompactThreads) s_big[i] = i < nr ? list[i] : ~0ull;
    bitonic_sort(s_big, np2);
    emit(q, s_big, nr, tid, kCompactThreads);
    __syncthreads();
  }
}

// k <= 32: one warp per row, the k best composites live in registers, one per lane and sorted
// ascending; the list is streamed 32 entries at a time, a ballot picks the entries that beat the
// running k-th best and each is inserted with two shuffles.  The survivors of a chunk all passed
// the OLD threshold, but only ~k ln((k + n) / k) of n beat the tightening one (40 of 128 for k = 20):
// ~400 warp instructions per row, no shared memory, no barriers -- the bitonic sort above needs
// ~2000 plus 36 warp barriers for the same row, and made the compaction 4.3 of the 29.8 ms of a
// 100k x 125k fused top-20 (three to five passes over 100k lists).
__global__ void __launch_bounds__(256)
topk_compact_warp_kernel(uint64_t *__restrict__ cand, int32_t *__restrict__ cnt, float *__restrict__ thr,
                         int64_t Q, int cap, int k, int largest, int fixed_cnt, int32_t *__restrict__ overflow,
                         float *__restrict__ d_out, int32_t *__restrict__ i_out,
                         int32_t *__restrict__ row_flags) {
  const int lane = threadIdx.x & 31;
  const int64_t q = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (q >= Q) return;
  const uint32_t flip = largest ? 0xFFFFFFFFu : 0u;
  int n = fixed_cnt >= 0 ? fixed_cnt : cnt[q];
  if (n > cap) {
    if (lane == 0) {
      atomicOr(overflow, 1);
      if (row_flags) row_flags[q] = 1;
    }
    n = cap;
  }
  uint64_t *list = cand + q * cap;
  constexpr uint64_t kNone = ~0ull;
  uint64_t best = kNone;  // lane l: the l-th best so far (lanes >= k stay kNone)
  uint64_t kth = kNone;   // the k-th best: only entries below it matter
  for (int base = 0; base < n; base += 32) {
    const uint64_t x = base + lane < n ? list[base + lane] : kNone;
    unsigned mask = __ballot_sync(0xffffffffu, x < kth);
    while (mask) {  // warp-uniform
      const int src = __ffs(mask) - 1;
      mask &= mask - 1;
      const uint64_t v = __shfl_sync(0xffffffffu, x, src);
      if (v < kth) {
        const int pos = __popc(__ballot_sync(0xffffffffu, best <= v));  // entries that stay in front of v
        const uint64_t up = __shfl_up_sync(0xffffffffu, best, 1);
        if (lane == pos) best = v;
        else if (lane > pos && lane < k) best = up;
        kth = __shfl_sync(0xffffffffu, best, k - 1);
      }
    }
  }
  const int m = n < k ? n : k;
  __syncwarp();
  if (lane < m) list[lane] = best;
  if (lane == 0) {
    cnt[q] = m;
    thr[q] = n >= k ? key_to_dist(static_cast<uint32_t>(kth >> 32) ^ flip) : (largest ? -INFINITY : INFINITY);
  }
  if (d_out && lane < k) {
    float dv = largest ? -INFINITY : INFINITY;
    int32_t iv = -1;
    if (lane < m) {
      dv = key_to_dist(static_cast<uint32_t>(best >> 32) ^ flip);
      iv = static_cast<int32_t>(static_cast<uint32_t>(best));
    }
    d_out[q * k + lane] = dv;
    i_out[q * k + lane] = iv;
  }
}

// Merge of per-shard top-k lists (gallery-sharded 1:N identification): vals / ids are
// [parts][Q][k] as torch.distributed.all_gather_into_tensor leaves them; one warp per query row
// streams its parts * k candidates through the same register-resident insertion as above.  Padded
// entries (id -1) lose against every real candidate.
__global__ void __launch_bounds__(256)
topk_merge_warp_kernel(const float *__restrict__ vals, const int32_t *__restrict__ ids, int parts, int64_t Q,
                       int k, int largest, float *__restrict__ d_out, int32_t *__restrict__ i_out) {
  const int lane = threadIdx.x & 31;
  const int64_t q = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (q >= Q) return;
  const uint32_t flip = largest ? 0xFFFFFFFFu : 0u;
  constexpr uint64_t kNone = ~0ull;
  uint64_t best = kNone, kth = kNone;
  const int n = parts * k;
  for (int base = 0; base < n; base += 32) {
    uint64_t x = kNone;
    if (base + lane < n) {
      const int part = (base + lane) / k, j = (base + lane) - part * k;
      const int64_t at = (static_cast<int64_t>(part) * Q + q) * k + j;
      const int32_t id = __ldg(ids + at);
      if (id >= 0) x = composite(dist_key(__ldg(vals + at)) ^ flip, static_cast<uint32_t>(id));
    }
    unsigned mask = __ballot_sync(0xffffffffu, x < kth);
    while (mask) {
      const int src = __ffs(mask) - 1;
      mask &= mask - 1;
      const uint64_t v = __shfl_sync(0xffffffffu, x, src);
      if (v < kth) {
        const int pos = __popc(__ballot_sync(0xffffffffu, best <= v));
        const uint64_t up = __shfl_up_sync(0xffffffffu, best, 1);
        if (lane == pos) best = v;
        else if (lane > pos && lane < k) best = up;
        kth = __shfl_sync(0xffffffffu, best, k - 1);
      }
    }
  }
  if (lane < k) {
    float dv = largest ? -INFINITY : INFINITY;
    int32_t iv = -1;
    if (best != kNone) {
      dv = key_to_dist(static_cast<uint32_t>(best >> 32) ^ flip);
      iv = static_cast<int32_t>(static_cast<uint32_t>(best));
    }
    d_out[q * k + lane] = dv;
    i_out[q * k + lane] = iv;
  }
}

}  // namespace

int launch_topk_merge(dali_ctx *ctx, const float *vals, const int32_t *ids, int parts, int64_t Q, int k,
                      int largest, float *d_out, int32_t *i_out) {
  if (Q == 0) return DALI_OK;
  if (k < 1 || k > 32 || parts < 1) return set_err(ctx, DALI_ERR_INVALID, "top-k merge: 1 <= k <= 32");
  KTimer t(ctx, DALI_K_TOPK);
  topk_merge_warp_kernel<<<static_cast<unsigned>((Q + 7) / 8), 256, 0, ctx->stream>>>(vals, ids, parts, Q, k, largest,
                                                                                      d_out, i_out);
  DALI_CUDA_OK(ctx, cudaGetLastError());
  return DALI_OK;
}

int launch_topk_compact(dali_ctx *ctx, uint64_t *cand, int32_t *cand_cnt, float *thr, int64_t Q,
                        int cap, int k, int largest, int fixed_cnt, int32_t *overflow, float *d_out,
                        int32_t *i_out, int32_t *row_flags) {
  if (Q == 0) return DALI_OK;
  if (cap > kCompactCap || k > cap)
    return set_err(ctx, DALI_ERR_INVALID, "top-k compaction: cap <= 1024 and k <= cap");
  KTimer t(ctx, DALI_K_TOPK);
  static const char *env_sort = getenv("DALI_TOPK_COMPACT_SORT");  // 1: the bitonic kernel also for k <= 32
  if (k <= 32 && !(env_sort && atoi(env_sort)))
    topk_compact_warp_kernel<<<static_cast<unsigned>((Q + 7) / 8), 256, 0, ctx->stream>>>(
        cand, cand_cnt, thr, Q, cap, k, largest, fixed_cnt, overflow, d_out, i_out, row_flags);
  else
    topk_compact_kernel<<<static_cast<unsigned>((Q + 3) / 4), kCompactThreads, 0, ctx->stream>>>(
        cand, cand_cnt, thr, Q, cap, k, largest, fixed_cnt, overflow, d_out, i_out, row_flags);
  DALI_CUDA_OK(ctx, cudaGetLastError());
  return DALI_OK;
}

namespace {

// Streaming selection from a materialised matrix, for long rows: the same threshold-filter +
// compaction scheme as the fused epilogue (distmat_umma2.cu), reading the matrix once at HBM
// speed instead of one latency-bound CTA per row.  Column chunks grow geometrically; a CTA filters
// one row segment against the row's running k-th best distance and appends the rare survivors.
constexpr int kFilterThreads = 256;

__global__ void __launch_bounds__(kFilterThreads)
topk_filter_kernel(const float *__restrict__ dist, int64_t ld, int64_t c_begin, int64_t c_end,
                   int64_t per_split, const float *__restrict__ thr, int32_t *__restrict__ cnt,
                   uint64_t *__restrict__ cand, int cap, int largest, int direct, int32_t id_base) {
  const int64_t q = blockIdx.x;
  const float *row = dist + q * ld;
  const int64_t c0 = c_begin + static_cast<int64_t>(blockIdx.y) * per_split;
  const int64_t c1 = c0 + per_split < c_end ? c0 + per_split : c_end;
  if (c0 >= c1) return;
  const float t = direct ? 0.f : __ldg(thr + q);
  const uint32_t flip = largest ? 0xFFFFFFFFu : 0u;
  uint64_t *list = cand + q * cap;
  auto offer = [&](float d, int64_t c) {
    const bool pass = direct ? true : (largest ? !(d < t) : !(d > t));
    if (pass) {
      const uint64_t cmp = composite(dist_key(d) ^ flip, static_cast<uint32_t>(id_base + c));
      if (direct) {
        list[c - c_begin] = cmp;
      } else {
        const int pos = atomicAdd(cnt + q, 1);
        if (pos < cap) list[pos] = cmp;
      }
    }
  };
  const int tid = threadIdx.x;
  const int mis = static_cast<int>((reinterpret_cast<uintptr_t>(row + c0) >> 2) & 3);
  int64_t head = (4 - mis) & 3;
  if (head > c1 - c0) head = c1 - c0;
  if (tid < head) offer(__ldg(row + c0 + tid), c0 + tid);
  const int64_t cv0 = c0 + head;
  const int64_t nvec = (c1 - cv0) >> 2;
  for (int64_t v = tid; v < nvec; v += kFilterThreads) {
    const float4 x = ld_stream_f4(row + cv0 + 4 * v);
    const int64_t c = cv0 + 4 * v;
    bool any;
    if (direct) any = true;
    else any = largest ? (!(x.x < t) | !(x.y < t) | !(x.z < t) | !(x.w < t))
                       : (!(x.x > t) | !(x.y > t) | !(x.z > t) | !(x.w > t));
    if (any) { offer(x.x, c); offer(x.y, c + 1); offer(x.z, c + 2); offer(x.w, c + 3); }
  }
  const int64_t ct0 = cv0 + 4 * nvec;
  if (tid < c1 - ct0) offer(__ldg(row + ct0 + tid), ct0 + tid);
}

// ---- rows of up to 16384 columns: one CTA per row, one pass (round 2) ---------------------------
// The row is staged in shared memory with cp.async (64 KB at most: three CTAs per SM keep ~190 KB of
// loads in flight), then
//   1. a strided sample of up to 2048 of its keys is histogrammed over 1024 bins of the sample's key
//      range; the bin that holds the sample's r-th smallest key (r ~ 5 k S / G + 4) gives an inclusive
//      threshold under which about 5 k + a few of the row's keys fall;
//   2. one pass over the staged row appends the composites (key, id) at or below the threshold to a
//      1024-entry list;
//   3. the list is bitonic-sorted (256 entries typically) and its first k leave.
// Exact: every key <= threshold is a candidate, so ties at the threshold are all in the list, and
// the composite order (key, id) decides.  A row whose list comes out shorter than k or longer than
// the capacity (massive ties, a sample that misses the tail) is flagged and redone by topk_kernel.
// Market matrix (3368 x 15913, k = 20): 0.22 ms for the chunked filter / compaction launches above
// -> 0.104 ms (k = 5: 0.089, k = 128: 0.167), tests/probes/topk_probe.py.
constexpr int kSelThreads = 256;
constexpr int kSelMaxG = 16384;
constexpr int kSelSample = 2048;
constexpr int kSelBins = 1024;
constexpr int kSelCap = 1024;  // candidate list: 512 entries for k <= 64 (three CTAs per SM at G = 16k), else 1024

__global__ void __launch_bounds__(kSelThreads)
topk_select_kernel(const float *__restrict__ dist, int64_t G, int64_t ld, int k, int cap, int largest, int32_t id_base,
                   float *__restrict__ d_out, int32_t *__restrict__ i_out, int32_t *__restrict__ row_flags) {
  extern __shared__ __align__(16) uint8_t sel_smem[];
  float *srow = reinterpret_cast<float *>(sel_smem);                 // [Gpad] the row (raw fp32)
  const int Gi = static_cast<int>(G);
  const int Gpad = (Gi + 3 + 3) & ~3;                                // + up to 3 leading columns of misalignment
  uint64_t *cand = reinterpret_cast<uint64_t *>(srow + Gpad);        // [cap]
  uint32_t *hist = reinterpret_cast<uint32_t *>(cand + cap);         // [kSelBins]
  __shared__ uint32_t s_min, s_max, s_thr;
  __shared__ int s_cnt;
  __shared__ uint32_t s_warp[kSelThreads / 32];
  const int64_t q = blockIdx.x;
  const float *row = dist + q * ld;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t flip = largest ? 0xFFFFFFFFu : 0u;

  // 0. stage the row: 16-byte pieces from the aligned address at or below the row start
  const int mis = static_cast<int>((reinterpret_cast<uintptr_t>(row) >> 2) & 3);  // columns before the row in its first piece
  const float *base = row - mis;
  const int nvec = (mis + Gi + 3) >> 2;
  const uint32_t srow_s = static_cast<uint32_t>(__cvta_generic_to_shared(srow));
  for (int v = tid; v < nvec; v += kSelThreads) {
    // (the last piece may reach past the row's end, i.e. possibly past the allocation: element-wise)
    if (4 * v + 4 <= mis + Gi)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(srow_s + 16u * v), "l"(base + 4 * v) : "memory");
    else
      for (int e = 4 * v; e < mis + Gi; ++e) srow[e] = base[e];
  }
  if (tid == 0) {
    s_min = 0xFFFFFFFFu;
    s_max = 0u;
    s_cnt = 0;
  }
  for (int i = tid; i < kSelBins; i += kSelThreads) hist[i] = 0u;
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  const float *r0 = srow + mis;  // r0[c] = column c

  // 1. sample: S keys at a fixed stride, their range, their histogram
  const int S = Gi < kSelSample ? Gi : kSelSample;
  const int stride = Gi / S;  // >= 1
  uint32_t sk[kSelSample / kSelThreads];
  uint32_t mn = 0xFFFFFFFFu, mx = 0u;
#pragma unroll
  for (int j = 0; j < kSelSample / kSelThreads; ++j) {
    const int sidx = tid + j * kSelThreads;
    sk[j] = sidx < S ? (dist_key(r0[sidx * stride]) ^ flip) : 0xFFFFFFFFu;
    if (sidx < S) {
      mn = min(mn, sk[j]);
      mx = max(mx, sk[j]);
    }
  }
  mn = __reduce_min_sync(0xffffffffu, mn);
  mx = __reduce_max_sync(0xffffffffu, mx);
  if (lane == 0) {
    atomicMin(&s_min, mn);
    atomicMax(&s_max, mx);
  }
  __syncthreads();
  const uint32_t kmin = s_min;
  const uint32_t span = s_max - kmin;
  const int bits = 32 - __clz(span | 1u);
  const int shift = bits > 10 ? bits - 10 : 0;  // bins of 2^shift keys: at most 1024 of them
#pragma unroll
  for (int j = 0; j < kSelSample / kSelThreads; ++j)
    if (tid + j * kSelThreads < S) atomicAdd(&hist[(sk[j] - kmin) >> shift], 1u);
  __syncthreads();
  // rank r of the sample whose bin bounds the candidates: about 5 k of the row below it, + a margin
  const int r = min(S, static_cast<int>(((k <= 64 ? 5ll : 3ll) * k * S + Gi - 1) / Gi) + 4);
  {
    // exclusive prefix over the bins, four per thread; the thread whose range holds rank r publishes
    const uint32_t h0 = hist[4 * tid], h1 = hist[4 * tid + 1], h2 = hist[4 * tid + 2], h3 = hist[4 * tid + 3];
    const uint32_t mine = h0 + h1 + h2 + h3;
    uint32_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += y;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t before = incl - mine;
    for (int w2 = 0; w2 < warp; ++w2) before += s_warp[w2];
    const uint32_t rr = static_cast<uint32_t>(r);
    if (before < rr && rr <= before + mine) {
      int b = 4 * tid;
      uint32_t acc = before + h0;
      if (acc < rr) { ++b; acc += h1; }
      if (acc < rr) { ++b; acc += h2; }
      if (acc < rr) { ++b; }
      // inclusive upper edge of bin b (saturating)
      const uint64_t edge = static_cast<uint64_t>(kmin) + ((static_cast<uint64_t>(b) + 1) << shift) - 1;
      s_thr = edge > 0xFFFFFFFFull ? 0xFFFFFFFFu : static_cast<uint32_t>(edge);
    }
  }
  __syncthreads();
  const uint32_t thr = s_thr;

  // 2. candidates: every key at or below the threshold
  auto offer = [&](float d, int c) {
    const uint32_t key = dist_key(d) ^ flip;
    if (key <= thr) {
      const int pos = atomicAdd(&s_cnt, 1);
      if (pos < cap) cand[pos] = composite(key, static_cast<uint32_t>(id_base + c));
    }
  };
  {
    // aligned float4 pieces of the staged row; the first / last piece masked to the row
    for (int v = tid; v < nvec; v += kSelThreads) {
      const float4 x = *reinterpret_cast<const float4 *>(srow + 4 * v);
      const int c = 4 * v - mis;
      if (c >= 0 && c + 3 < Gi) {
        offer(x.x, c); offer(x.y, c + 1); offer(x.z, c + 2); offer(x.w, c + 3);
      } else {
        if (c >= 0 && c < Gi) offer(x.x, c);
        if (c + 1 >= 0 && c + 1 < Gi) offer(x.y, c + 1);
        if (c + 2 >= 0 && c + 2 < Gi) offer(x.z, c + 2);
        if (c + 3 >= 0 && c + 3 < Gi) offer(x.w, c + 3);
      }
    }
  }
  __syncthreads();
  const int cnt = s_cnt;
  const int want = k < Gi ? k : Gi;
  if (cnt < want || cnt > cap) {  // block-uniform: the one-CTA-per-row kernel redoes this row
    if (tid == 0) row_flags[q] = 1;
    return;
  }
  // 3. short lists: rank by counting (the composites are distinct), the k best written straight to
  // their places -- ~cnt broadcast loads per candidate and no barrier (a bitonic sort of 256 entries
  // is 36 barriers: k = 20 at the Market shape 0.117 -> 0.103 ms); long lists (large k): sort
  if (cnt > 256) {  // block-uniform
    const int np2 = next_pow2(cnt);
    for (int i = cnt + tid; i < np2; i += kSelThreads) cand[i] = ~0ull;
    bitonic_sort(cand, np2);
    for (int i = tid; i < k; i += kSelThreads) {
      const uint64_t c = cand[i];  // cnt >= k here
      d_out[q * k + i] = key_to_dist(static_cast<uint32_t>(c >> 32) ^ flip);
      i_out[q * k + i] = static_cast<int32_t>(static_cast<uint32_t>(c));
    }
    return;
  }
  for (int i = tid; i < cnt; i += kSelThreads) {
    const uint64_t c = cand[i];
    int rank = 0;
    for (int j = 0; j < cnt; ++j) rank += cand[j] < c ? 1 : 0;
    if (rank < k) {
      d_out[q * k + rank] = key_to_dist(static_cast<uint32_t>(c >> 32) ^ flip);
      i_out[q * k + rank] = static_cast<int32_t>(static_cast<uint32_t>(c));
    }
  }
  for (int i = cnt + tid; i < k; i += kSelThreads) {  // rows shorter than k
    d_out[q * k + i] = largest ? -INFINITY : INFINITY;
    i_out[q * k + i] = -1;
  }
}

// ---- rows of up to 16384 columns, second form: the row stays in registers ------------------------
// 512 threads, thread t holds the 16-byte pieces t, t + 512, ... (eight at most: 32 order keys in
// registers, all eight loads in flight at once, no staging in shared memory, two CTAs per SM).
//   1. every group of `gs` adjacent threads takes the minimum of its keys: M = 512 / gs >= 2 k group
//      minima.  The k smallest of them are k different elements of the row, so the k-th smallest
//      minimum T bounds the row's k-th smallest key from above -- and closely: about k (1 + k / M)
//      keys lie at or below it (Market row, k = 20: ~25 candidates, against ~130 from the sample
//      histogram of topk_select_kernel);
//   2. T = the minimum whose (stable) rank among the M minima is k - 1, found by counting;
//   3. every thread offers its keys at or below T to the candidate list; the list is ranked by
//      counting and its k best leave.
// Exact for the same reason as above (every key <= T is a candidate, the composite order decides);
// a row with more candidates than the list holds (massive ties) is flagged for topk_kernel.
constexpr int kMinPieces = 8;     // 16-byte pieces per thread
constexpr int kMinCap = 512;      // candidate list
constexpr int kMinMaxGroups = 256;

// 16 bytes at p + OFF (a compile-time byte offset: one address register for all of a thread's pieces)
// if `on`, `fill` otherwise; a predicated-off load touches nothing
template <int OFF>
__device__ __forceinline__ float4 ld_piece_if(const float4 *p, bool on, float fill) {
  float4 x;
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %5, 0;\n\t"
      "mov.f32 %0, %7;\n\tmov.f32 %1, %7;\n\tmov.f32 %2, %7;\n\tmov.f32 %3, %7;\n\t"
      "@p ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4 + %6];\n\t}"
      : "=&f"(x.x), "=&f"(x.y), "=&f"(x.z), "=&f"(x.w)
      : "l"(p), "r"(on ? 1 : 0), "n"(OFF), "f"(fill));
  return x;
}

// THREADS = 512 (two CTAs per SM, rows of up to 16378 columns) or 1024 (one CTA per SM, up to 32762)
template <int THREADS>
__global__ void __launch_bounds__(THREADS, 1024 / THREADS)
topk_minima_kernel(const float *__restrict__ dist, int64_t G, int64_t ld, int k, int log2gs, int largest,
                   int32_t id_base, float *__restrict__ d_out, int32_t *__restrict__ i_out,
                   int32_t *__restrict__ row_flags) {
  __shared__ float gm[kMinMaxGroups];
  __shared__ uint64_t cand[kMinCap];
  __shared__ float s_thr;
  __shared__ int s_cnt, s_nan;
  const int64_t q = blockIdx.x;
  const float *row = dist + q * ld;
  const int tid = threadIdx.x;
  const int Gi = static_cast<int>(G);
  const uint32_t flip = largest ? 0xFFFFFFFFu : 0u;
  // 0. the row: [a0, a0 + 4 nv) is its 16-byte aligned part, thread t holds the pieces t, t + 512, ...
  // (all loads unconditional or predicated -- no branch between them, so the eight are in flight
  // together); the <= 3 columns before and after it go to the threads 0 .. 2 and 32 .. 34 as one
  // extra key each.  Slots without a column get the largest key and are never offered.
  const int mis = static_cast<int>((reinterpret_cast<uintptr_t>(row) >> 2) & 3);
  const int a0 = min((4 - mis) & 3, Gi);
  const int nv = (Gi - a0) >> 2;
  const int ct0 = a0 + 4 * nv;
  if (tid == 0) {
    s_cnt = 0;
    s_nan = 0;
    s_thr = INFINITY;
  }
  const float4 *vrow = reinterpret_cast<const float4 *>(row + a0);
  float4 x[kMinPieces];
  {
    const float4 *p0 = vrow + tid;
    const float fill = largest ? -INFINITY : INFINITY;  // times sgn below: +inf, never under a threshold
    x[0] = ld_piece_if<0 * THREADS * 16>(p0, tid + 0 * THREADS < nv, fill);
    x[1] = ld_piece_if<1 * THREADS * 16>(p0, tid + 1 * THREADS < nv, fill);
    x[2] = ld_piece_if<2 * THREADS * 16>(p0, tid + 2 * THREADS < nv, fill);
    x[3] = ld_piece_if<3 * THREADS * 16>(p0, tid + 3 * THREADS < nv, fill);
    x[4] = ld_piece_if<4 * THREADS * 16>(p0, tid + 4 * THREADS < nv, fill);
    x[5] = ld_piece_if<5 * THREADS * 16>(p0, tid + 5 * THREADS < nv, fill);
    x[6] = ld_piece_if<6 * THREADS * 16>(p0, tid + 6 * THREADS < nv, fill);
    x[7] = ld_piece_if<7 * THREADS * 16>(p0, tid + 7 * THREADS < nv, fill);
    static_assert(kMinPieces == 8, "eight pieces per thread");
  }
  // (the values are not touched before all eight requests have been issued)
#pragma unroll
  for (int j = 0; j < kMinPieces; ++j)
    asm volatile("" : "+f"(x[j].x), "+f"(x[j].y), "+f"(x[j].z), "+f"(x[j].w));
  // everything up to the candidate list works on the distances themselves (times -1 for `largest`):
  // for numbers the float order is the key order (-0 == +0 in both); NaN compares false everywhere,
  // so a NaN is never a candidate -- right for the smallest k unless fewer than k numbers exist (the
  // row is then flagged below), wrong for the largest k, where NaN ranks first: such rows are flagged
  const float sgn = largest ? -1.f : 1.f;
  int ec = -1;  // column of the extra element
  if (tid < a0) ec = tid;
  if (tid >= 32 && tid - 32 < Gi - ct0) ec = ct0 + tid - 32;
  const float ex = ec >= 0 ? __ldg(row + ec) * sgn : INFINITY;
  float y[kMinPieces][4];
#pragma unroll
  for (int j = 0; j < kMinPieces; ++j) {
    y[j][0] = x[j].x * sgn; y[j][1] = x[j].y * sgn; y[j][2] = x[j].z * sgn; y[j][3] = x[j].w * sgn;
  }
  if (largest) {  // uniform
    bool nan = ex != ex;
#pragma unroll
    for (int j = 0; j < kMinPieces; ++j)
#pragma unroll
      for (int u = 0; u < 4; ++u) nan |= y[j][u] != y[j][u];
    if (nan) s_nan = 1;
  }
  // 1. group minima (fminf skips NaN; slots without a column hold +inf)
  float mn = ex;
#pragma unroll
  for (int j = 0; j < kMinPieces; ++j)
#pragma unroll
    for (int u = 0; u < 4; ++u) mn = fminf(mn, y[j][u]);
  for (int o = 1; o < (1 << log2gs); o <<= 1) mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
  const int M = THREADS >> log2gs;
  if ((tid & ((1 << log2gs) - 1)) == 0) gm[tid >> log2gs] = mn;
  __syncthreads();
  // 2. the minimum of rank want - 1 (ties by group index).  Whatever value comes out, the result is
  // exact as long as at least `want` elements lie at or below it.
  const int want = k < Gi ? k : Gi;
  if (tid < M) {
    const float mine = gm[tid];
    int rank = 0;
    for (int j = 0; j < M; ++j) {
      const float o = gm[j];
      rank += (o < mine || (o == mine && j < tid)) ? 1 : 0;
    }
    if (rank == want - 1) s_thr = mine;
  }
  __syncthreads();
  const float thr = s_thr;
  // 3. candidates: every element of the row at or below the threshold
  auto offer = [&](float v, int c) {
    const int pos = atomicAdd(&s_cnt, 1);
    if (pos < kMinCap) cand[pos] = composite(dist_key(v * sgn) ^ flip, static_cast<uint32_t>(id_base + c));
  };
#pragma unroll
  for (int j = 0; j < kMinPieces; ++j) {
    const int v = tid + j * THREADS;
    // (one test per piece first: a piece with a candidate is rare)
    if (v < nv && fminf(fminf(y[j][0], y[j][1]), fminf(y[j][2], y[j][3])) <= thr) {
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (y[j][u] <= thr) offer(y[j][u], a0 + 4 * v + u);
    }
  }
  if (ec >= 0 && ex <= thr) offer(ex, ec);
  __syncthreads();
  const int cnt = s_cnt;
  if (cnt < want || cnt > kMinCap || s_nan) {  // block-uniform: the one-CTA-per-row kernel redoes this row
    if (tid == 0) row_flags[q] = 1;
    return;
  }
  for (int i = tid; i < cnt; i += THREADS) {
    const uint64_t c = cand[i];
    int rank = 0;
    for (int j = 0; j < cnt; ++j) rank += cand[j] < c ? 1 : 0;
    if (rank < k) {
      d_out[q * k + rank] = key_to_dist(static_cast<uint32_t>(c >> 32) ^ flip);
      i_out[q * k + rank] = static_cast<int32_t>(static_cast<uint32_t>(c));
    }
  }
  for (int i = cnt + tid; i < k; i += THREADS) {  // rows shorter than k
    d_out[q * k + i] = largest ? -INFINITY : INFINITY;
    i_out[q * k + i] = -1;
  }
}

}  // namespace

static int launch_topk_classic(dali_ctx *ctx, const float *dist, int64_t Q, int64_t G, int64_t ld, int k,
                               int largest, const int32_t *col_ids, int32_t id_base, float *d_out,
                               int32_t *i_out, const int32_t *only_rows) {
  KTimer t(ctx, DALI_K_TOPK);
  topk_kernel<<<static_cast<unsigned>(Q), kTopkThreads, 0, ctx->stream>>>(
      dist, G, ld, k, largest, col_ids, id_base, d_out, i_out, only_rows);
  DALI_CUDA_OK(ctx, cudaGetLastError());
  return DALI_OK;
}

int launch_topk(dali_ctx *ctx, const float *dist, int64_t Q, int64_t G, int64_t ld, int k,
                int largest, const int32_t *col_ids, int32_t id_base, float *d_out,
                int32_t *i_out) {
  if (Q == 0) return DALI_OK;
  if (k < 1 || k > 128) return set_err(ctx, DALI_ERR_INVALID, "top-k needs 1 <= k <= 128");
  // rows that fit the registers of one CTA: one pass, one launch (+ the repair launch for flagged rows);
  // threshold from group minima: THREADS / gs >= 2 k groups
  static const char *env_min = getenv("DALI_TOPK_MINIMA");
  if (!col_ids && G >= 256 && (G + 6) / 4 <= 1024ll * kMinPieces && !(env_min && atoi(env_min) == 0)) {
    void *flags_v;
    int rc = ws_ensure(ctx, WS_CAND_CNT, sizeof(int32_t) * 2 * Q, &flags_v);
    if (rc) return rc;
    int32_t *row_flags = static_cast<int32_t *>(flags_v) + Q;
    DALI_CUDA_OK(ctx, cudaMemsetAsync(row_flags, 0, sizeof(int32_t) * Q, ctx->stream));
    const int log2gs = k <= 32 ? 3 : k <= 64 ? 2 : 1;
    {
      KTimer t(ctx, DALI_K_TOPK);
      if ((G + 6) / 4 <= 512ll * kMinPieces)
        topk_minima_kernel<512><<<static_cast<unsigned>(Q), 512, 0, ctx->stream>>>(
            dist, G, ld, k, log2gs, largest, id_base, d_out, i_out, row_flags);
      else
        topk_minima_kernel<1024><<<static_cast<unsigned>(Q), 1024, 0, ctx->stream>>>(
            dist, G, ld, k, log2gs + 1, largest, id_base, d_out, i_out, row_flags);
      DALI_CUDA_OK(ctx, cudaGetLastError());
    }
    return launch_topk_classic(ctx, dist, Q, G, ld, k, largest, nullptr, id_base, d_out, i_out, row_flags);
  }
  // rows that fit shared memory: one pass, one launch (+ the repair launch for flagged rows)
  static const char *env_sel = getenv("DALI_TOPK_SELECT");
  if (!col_ids && G >= 256 && G <= kSelMaxG && !(env_sel && atoi(env_sel) == 0)) {
    void *flags_v;
    int rc = ws_ensure(ctx, WS_CAND_CNT, sizeof(int32_t) * 2 * Q, &flags_v);
    if (rc) return rc;
    int32_t *row_flags = static_cast<int32_t *>(flags_v) + Q;
    DALI_CUDA_OK(ctx, cudaMemsetAsync(row_flags, 0, sizeof(int32_t) * Q, ctx->stream));
    const int cap = k <= 64 ? kSelCap / 2 : kSelCap;
    const size_t smem = sizeof(float) * ((G + 6) & ~int64_t(3)) + sizeof(uint64_t) * cap + sizeof(uint32_t) * kSelBins;
    if ((rc = ensure_dyn_smem(ctx, reinterpret_cast<const void *>(&topk_select_kernel), smem))) return rc;
    {
      KTimer t(ctx, DALI_K_TOPK);
      topk_select_kernel<<<static_cast<unsigned>(Q), kSelThreads, smem, ctx->stream>>>(dist, G, ld, k, cap, largest,
                                                                                   id_base, d_out, i_out, row_flags);
      DALI_CUDA_OK(ctx, cudaGetLastError());
    }
    return launch_topk_classic(ctx, dist, Q, G, ld, k, largest, nullptr, id_base, d_out, i_out, row_flags);
  }
  static const char *env = getenv("DALI_TOPK_STREAM");
  // measured (B200): the streaming path wins for long rows (17.5k x 63k: 1.39 vs 2.44 ms) and for
  // few rows (3368 x 15913: 0.22 vs 0.37 ms, the classic kernel has too few CTAs in flight); with
  // many medium rows (19281 x 19281, re-ranking) its per-chunk launches cost more than they save
  const bool stream_ok = !col_ids && G >= 4096 && (G >= 32768 || Q < 8192) && !(env && atoi(env) == 0);
  if (!stream_ok)
    return launch_topk_classic(ctx, dist, Q, G, ld, k, largest, col_ids, id_base, d_out, i_out, nullptr);
  constexpr int kCapList = kCompactCap;
  void *cand_v, *cnt_v, *thr_v, *flag_v;
  int rc = ws_ensure(ctx, WS_CAND, sizeof(uint64_t) * Q * kCapList, &cand_v);
  if (rc) return rc;
  if ((rc = ws_ensure(ctx, WS_CAND_CNT, sizeof(int32_t) * 2 * Q, &cnt_v))) return rc;  // counts | row flags
  if ((rc = ws_ensure(ctx, WS_THR, sizeof(float) * Q, &thr_v))) return rc;
  if ((rc = ws_ensure(ctx, WS_FLAG, 256, &flag_v))) return rc;
  uint64_t *cand = static_cast<uint64_t *>(cand_v);
  int32_t *cnt = static_cast<int32_t *>(cnt_v);
  int32_t *row_flags = cnt + Q;
  float *thr = static_cast<float *>(thr_v);
  int32_t *flag = static_cast<int32_t *>(flag_v);
  DALI_CUDA_OK(ctx, cudaMemsetAsync(row_flags, 0, sizeof(int32_t) * Q, ctx->stream));
  DALI_CUDA_OK(ctx, cudaMemsetAsync(flag, 0, sizeof(int32_t), ctx->stream));
  const int64_t growth = std::max<int64_t>(1, std::min<int64_t>(8, kCapList / (8 * k)));
  const int64_t first_cols = std::min<int64_t>(kCapList, std::max<int64_t>(2 * k, 256));
  int64_t seen = 0;
  while (seen < G) {
    const bool first = seen == 0;
    const int64_t chunk = first ? std::min<int64_t>(G, first_cols)
                                : std::min<int64_t>(G - seen, std::max<int64_t>(256, growth * seen));
    // split long segments so that small Q still fills the machine; >= 2048 columns per CTA
    int64_t nsplit = std::max<int64_t>(1, std::min<int64_t>((chunk + 8191) / 8192,
                                                            (8ll * ctx->num_sms + Q - 1) / Q));
    const int64_t per_split = ((chunk + nsplit - 1) / nsplit + 3) & ~int64_t(3);
    nsplit = (chunk + per_split - 1) / per_split;
    {
      KTimer t(ctx, DALI_K_TOPK);
      const dim3 grid(static_cast<unsigned>(Q), static_cast<unsigned>(nsplit));
      topk_filter_kernel<<<grid, kFilterThreads, 0, ctx->stream>>>(dist, ld, seen, seen + chunk, per_split, thr,
                                                                  cnt, cand, kCapList, largest, first ? 1 : 0,
                                                                  id_base);
      DALI_CUDA_OK(ctx, cudaGetLastError());
    }
    seen += chunk;
    const bool last = seen >= G;
    rc = launch_topk_compact(ctx, cand, cnt, thr, Q, kCapList, k, largest, first ? static_cast<int>(chunk) : -1,
                             flag, last ? d_out : nullptr, last ? i_out : nullptr, row_flags);
    if (rc) return rc;
  }
  // repair pass, device side: rows whose list overflowed in some chunk (massive ties, adversarial
  // column order) are redone by the one-CTA-per-row kernel; every other CTA exits at once
  return launch_topk_classic(ctx, dist, Q, G, ld, k, largest, nullptr, id_base, d_out, i_out, row_flags);
}

}  // namespace dali
