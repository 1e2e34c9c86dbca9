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
            int32_t *__restrict__ i_out) {
  __shared__ TopkShared s;
  const int64_t q = blockIdx.x;
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

__global__ void __launch_bounds__(kCompactThreads)
topk_compact_kernel(uint64_t *__restrict__ cand, int32_t *__restrict__ cnt, float *__restrict__ thr,
                    int cap, int k, int largest, int fixed_cnt, int32_t *__restrict__ overflow,
                    float *__restrict__ d_out, int32_t *__restrict__ i_out) {
  __shared__ uint64_t s[kCompactCap];
  const int64_t q = blockIdx.x;
  const int tid = threadIdx.x;
  const uint32_t flip = largest ? 0xFFFFFFFFu : 0u;
  int n = fixed_cnt >= 0 ? fixed_cnt : cnt[q];
  if (n > cap) {
    if (tid == 0) atomicOr(overflow, 1);
    n = cap;
  }
  uint64_t *list = cand + q * cap;
  const int np2 = next_pow2(n);
  for (int i = tid; i < np2; i += kCompactThreads) s[i] = i < n ? list[i] : ~0ull;
  bitonic_sort(s, np2);
  const int m = n < k ? n : k;
  for (int i = tid; i < m; i += kCompactThreads) list[i] = s[i];
  if (tid == 0) {
    cnt[q] = m;
    thr[q] = n >= k ? key_to_dist(static_cast<uint32_t>(s[k - 1] >> 32) ^ flip)
                    : (largest ? -INFINITY : INFINITY);
  }
  if (d_out) {
    for (int i = tid; i < k; i += kCompactThreads) {
      float dv = largest ? -INFINITY : INFINITY;
      int32_t iv = -1;
      if (i < m) {
        dv = key_to_dist(static_cast<uint32_t>(s[i] >> 32) ^ flip);
        iv = static_cast<int32_t>(static_cast<uint32_t>(s[i]));
      }
      d_out[q * k + i] = dv;
      i_out[q * k + i] = iv;
    }
  }
}

}  // namespace

int launch_topk_compact(dali_ctx *ctx, uint64_t *cand, int32_t *cand_cnt, float *thr, int64_t Q,
                        int cap, int k, int largest, int fixed_cnt, int32_t *overflow, float *d_out,
                        int32_t *i_out) {
  if (Q == 0) return DALI_OK;
  if (cap > kCompactCap || k > cap)
    return set_err(ctx, DALI_ERR_INVALID, "top-k compaction: cap <= 1024 and k <= cap");
  KTimer t(ctx, DALI_K_TOPK);
  topk_compact_kernel<<<static_cast<unsigned>(Q), kCompactThreads, 0, ctx->stream>>>(
      cand, cand_cnt, thr, cap, k, largest, fixed_cnt, overflow, d_out, i_out);
  DALI_CUDA_OK(ctx, cudaGetLastError());
  return DALI_OK;
}

int launch_topk(dali_ctx *ctx, const float *dist, int64_t Q, int64_t G, int64_t ld, int k,
                int largest, const int32_t *col_ids, int32_t id_base, float *d_out,
                int32_t *i_out) {
  if (Q == 0) return DALI_OK;
  if (k < 1 || k > 128) return set_err(ctx, DALI_ERR_INVALID, "top-k needs 1 <= k <= 128");
  KTimer t(ctx, DALI_K_TOPK);
  topk_kernel<<<static_cast<unsigned>(Q), kTopkThreads, 0, ctx->stream>>>(
      dist, G, ld, k, largest, col_ids, id_base, d_out, i_out);
  DALI_CUDA_OK(ctx, cudaGetLastError());
  return DALI_OK;
}

}  // namespace dali
