// Sort-free ranking: positive-rank counting + CMC/AP epilogue.
//
// Replaces the arithmetic behind torchreid.metrics.evaluate_rank as called at
// validateModels.py:68-69, evaluate.py:312-313, evaluate_ensembled_models.py:324-325,
// evaluateCleanATModels.py:266-267.  Instead of argsort + gather + per-query loop, the
// 1-based kept rank of every valid positive p of query q is
//     1 + #{ j : (key(d[q,j]), j) <lex (key(d[q,p]), p) }  -  #{ junk u : (key_u,u) < (key_p,p) }
// which needs ONE streaming pass over the distance row (4 B / pair, HBM bound) and no
// label reads in the hot loop (junk items are all inside the query's match list).
#include "common.cuh"

namespace dali {

namespace {

constexpr int kCountThreads = 256;
constexpr int kChunk = 32;  // thresholds held in registers per CTA

__device__ __forceinline__ float4 ld_stream_f4(const float *p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}

// ------------------------------ gather ------------------------------------
__global__ void rank_gather_kernel(const float *__restrict__ dist, int64_t ld, int64_t g0,
                                   int64_t Gs, const int64_t *__restrict__ off,
                                   const int32_t *__restrict__ gid, uint32_t *__restrict__ keys) {
  const int64_t q = blockIdx.x;
  const int64_t o = off[q];
  const int m = static_cast<int>(off[q + 1] - o);
  const float *row = dist + q * ld;
  for (int t = threadIdx.x; t < m; t += blockDim.x) {
    const int64_t local = static_cast<int64_t>(gid[o + t]) - g0;
    uint32_t k = 0u;
    if (local >= 0 && local < Gs) k = dist_key(row[local]);
    keys[o + t] = k;
  }
}

// ------------------------------ count --------------------------------------
template <int NT>
struct Thr {
  uint64_t c[NT];
};

template <int NT>
__device__ __forceinline__ void cmp_acc(const Thr<NT> &thr, int (&cnt)[NT], float d, uint32_t g) {
  const uint64_t c = composite(dist_key(d), g);
#pragma unroll
  for (int i = 0; i < NT; ++i) cnt[i] += (c < thr.c[i]) ? 1 : 0;
}

template <int NT>
__device__ __forceinline__ void count_row(const float *__restrict__ row, int64_t c0, int64_t c1,
                                          uint32_t gbase, const uint32_t *__restrict__ tkeys,
                                          const int32_t *__restrict__ tgids, int n,
                                          int32_t *__restrict__ out, bool use_atomic,
                                          int *s_acc) {
  Thr<NT> thr;
  int cnt[NT];
#pragma unroll
  for (int i = 0; i < NT; ++i) {
    // out-of-range slots get composite 0: nothing is smaller, they never count
    thr.c[i] = (i < n) ? composite(__ldg(tkeys + i), static_cast<uint32_t>(__ldg(tgids + i))) : 0ull;
    cnt[i] = 0;
  }
  const int tid = threadIdx.x;
  // head: scalar until the address is 16-byte aligned
  const int mis = static_cast<int>((reinterpret_cast<uintptr_t>(row + c0) >> 2) & 3);
  int64_t head = (4 - mis) & 3;
  if (head > c1 - c0) head = c1 - c0;
  if (tid < head) cmp_acc<NT>(thr, cnt, __ldg(row + c0 + tid), gbase + static_cast<uint32_t>(c0 + tid));
  const int64_t cv0 = c0 + head;
  const int64_t nvec = (c1 - cv0) >> 2;
  int64_t v = tid;
  for (; v + kCountThreads < nvec; v += 2 * kCountThreads) {
    const float4 x0 = ld_stream_f4(row + cv0 + 4 * v);
    const float4 x1 = ld_stream_f4(row + cv0 + 4 * (v + kCountThreads));
    const uint32_t g0 = gbase + static_cast<uint32_t>(cv0 + 4 * v);
    const uint32_t g1 = gbase + static_cast<uint32_t>(cv0 + 4 * (v + kCountThreads));
    cmp_acc<NT>(thr, cnt, x0.x, g0);
    cmp_acc<NT>(thr, cnt, x0.y, g0 + 1);
    cmp_acc<NT>(thr, cnt, x0.z, g0 + 2);
    cmp_acc<NT>(thr, cnt, x0.w, g0 + 3);
    cmp_acc<NT>(thr, cnt, x1.x, g1);
    cmp_acc<NT>(thr, cnt, x1.y, g1 + 1);
    cmp_acc<NT>(thr, cnt, x1.z, g1 + 2);
    cmp_acc<NT>(thr, cnt, x1.w, g1 + 3);
  }
  if (v < nvec) {
    const float4 x0 = ld_stream_f4(row + cv0 + 4 * v);
    const uint32_t g0 = gbase + static_cast<uint32_t>(cv0 + 4 * v);
    cmp_acc<NT>(thr, cnt, x0.x, g0);
    cmp_acc<NT>(thr, cnt, x0.y, g0 + 1);
    cmp_acc<NT>(thr, cnt, x0.z, g0 + 2);
    cmp_acc<NT>(thr, cnt, x0.w, g0 + 3);
  }
  const int64_t ct0 = cv0 + 4 * nvec;
  if (tid < c1 - ct0) cmp_acc<NT>(thr, cnt, __ldg(row + ct0 + tid), gbase + static_cast<uint32_t>(ct0 + tid));

  // block reduction: redux.sync per counter, one shared atomic per warp
#pragma unroll
  for (int i = 0; i < NT; ++i) {
    const int w = __reduce_add_sync(0xffffffffu, cnt[i]);
    if ((tid & 31) == 0 && w) atomicAdd(&s_acc[i], w);
  }
  __syncthreads();
  if (tid < n) {
    if (use_atomic) {
      if (s_acc[tid]) atomicAdd(out + tid, s_acc[tid]);
    } else {
      out[tid] = s_acc[tid];
    }
  }
}

__global__ void __launch_bounds__(kCountThreads)
rank_count_kernel(const float *__restrict__ dist, int64_t ld, int64_t g0, int64_t Gs,
                  const int64_t *__restrict__ off, const int32_t *__restrict__ nvalid,
                  const int32_t *__restrict__ gid, const uint32_t *__restrict__ keys,
                  int32_t *__restrict__ counts, int nsplit) {
  __shared__ int s_acc[kChunk];
  const int64_t q = blockIdx.x;
  const int chunk = blockIdx.y;
  const int nv = nvalid[q];
  if (chunk * kChunk >= nv) return;  // whole CTA exits (uniform)
  if (threadIdx.x < kChunk) s_acc[threadIdx.x] = 0;
  __syncthreads();
  const int n = min(kChunk, nv - chunk * kChunk);
  const int64_t o = off[q] + static_cast<int64_t>(chunk) * kChunk;
  // column range of this split, in multiples of 4 columns
  const int64_t per = (((Gs + nsplit - 1) / nsplit) + 3) & ~int64_t(3);
  const int64_t c0r = per * static_cast<int64_t>(blockIdx.z);
  const int64_t c0 = c0r < Gs ? c0r : Gs;
  const int64_t c1 = (c0 + per) < Gs ? (c0 + per) : Gs;
  const float *row = dist + q * ld;
  const uint32_t gbase = static_cast<uint32_t>(g0);
  const bool atom = nsplit > 1;
  if (n <= 8)
    count_row<8>(row, c0, c1, gbase, keys + o, gid + o, n, counts + o, atom, s_acc);
  else if (n <= 16)
    count_row<16>(row, c0, c1, gbase, keys + o, gid + o, n, counts + o, atom, s_acc);
  else if (n <= 24)
    count_row<24>(row, c0, c1, gbase, keys + o, gid + o, n, counts + o, atom, s_acc);
  else
    count_row<32>(row, c0, c1, gbase, keys + o, gid + o, n, counts + o, atom, s_acc);
}

// ------------------------------ finalize -----------------------------------
constexpr int kFinThreads = 128;
constexpr int kFinCap = 2048;  // matches staged in shared memory

__global__ void __launch_bounds__(kFinThreads)
rank_finalize_kernel(const int64_t *__restrict__ off, const int32_t *__restrict__ nvalid,
                     const int32_t *__restrict__ gid, const uint32_t *__restrict__ keys,
                     const int32_t *__restrict__ counts, int max_rank,
                     int32_t *__restrict__ ranks_sorted, float *__restrict__ ap,
                     int32_t *__restrict__ first_rank, int32_t *__restrict__ cmc_cnt) {
  __shared__ uint64_t s_c[kFinCap];
  __shared__ int32_t s_rank[kFinCap];
  const int64_t q = blockIdx.x;
  const int64_t o = off[q];
  const int m = static_cast<int>(off[q + 1] - o);
  const int nv = nvalid[q];
  const int tid = threadIdx.x;
  if (nv == 0) {
    if (tid == 0) {
      ap[q] = 0.f;
      first_rank[q] = -1;
    }
    return;
  }
  const bool staged = m <= kFinCap;
  if (staged) {
    for (int t = tid; t < m; t += kFinThreads)
      s_c[t] = composite(keys[o + t], static_cast<uint32_t>(gid[o + t]));
    __syncthreads();
  }
  for (int t = tid; t < nv; t += kFinThreads) {
    const uint64_t ck = staged ? s_c[t] : composite(keys[o + t], static_cast<uint32_t>(gid[o + t]));
    int below_valid = 0, below_junk = 0;
    for (int u = 0; u < m; ++u) {
      const uint64_t cu =
          staged ? s_c[u] : composite(keys[o + u], static_cast<uint32_t>(gid[o + u]));
      const int lt = cu < ck ? 1 : 0;
      if (u < nv) below_valid += lt; else below_junk += lt;
    }
    const int r = counts[o + t] - below_junk + 1;  // 1-based rank among kept gallery items
    ranks_sorted[o + below_valid] = r;             // below_valid = (k-1), a permutation of 0..nv-1
    if (staged) s_rank[below_valid] = r;
  }
  __syncthreads();
  if (tid == 0) {
    // torchreid Cython accumulation: float running sum, each term formed in double
    float s = 0.f;
    for (int k = 1; k <= nv; ++k) {
      const int r = staged ? s_rank[k - 1] : ranks_sorted[o + k - 1];
      s = static_cast<float>(static_cast<double>(s) + static_cast<double>(k) / static_cast<double>(r));
    }
    ap[q] = s / static_cast<float>(nv);
    const int fr = staged ? s_rank[0] : ranks_sorted[o];
    first_rank[q] = fr;
    if (fr <= max_rank) atomicAdd(cmc_cnt + (fr - 1), 1);
    atomicAdd(cmc_cnt + max_rank, 1);  // num_valid_q
  }
}

}  // namespace

int launch_rank_gather(dali_ctx *ctx, const dali_rank_plan *plan, const float *dist, int64_t ld,
                       int64_t g0, int64_t Gs, uint32_t *keys) {
  if (plan->M == 0 || plan->Q == 0) return DALI_OK;
  KTimer t(ctx, DALI_K_RANK_GATHER);
  rank_gather_kernel<<<static_cast<unsigned>(plan->Q), 128, 0, ctx->stream>>>(
      dist, ld, g0, Gs, plan->d_off, plan->d_gid, keys);
  DALI_CUDA_OK(ctx, cudaGetLastError());
  return DALI_OK;
}

int launch_rank_count(dali_ctx *ctx, const dali_rank_plan *plan, const float *dist, int64_t ld,
                      int64_t g0, int64_t Gs, const uint32_t *keys, int32_t *counts) {
  if (plan->M == 0 || plan->Q == 0) return DALI_OK;
  DALI_CUDA_OK(ctx, cudaMemsetAsync(counts, 0, sizeof(int32_t) * plan->M, ctx->stream));
  if (plan->max_nv == 0 || Gs == 0) return DALI_OK;
  const int nchunk = (plan->max_nv + kChunk - 1) / kChunk;
  if (nchunk > 65535) return set_err(ctx, DALI_ERR_UNSUPPORTED, "more than 2M positives for one query");
  // enough CTAs for >= 4 per SM; never split a row below 4096 columns
  int64_t want = (4ll * ctx->num_sms + plan->Q - 1) / (plan->Q > 0 ? plan->Q : 1);
  int64_t max_split = (Gs + 4095) / 4096;
  int nsplit = static_cast<int>(want < 1 ? 1 : (want > max_split ? max_split : want));
  if (nsplit < 1) nsplit = 1;
  if (nsplit > 1024) nsplit = 1024;
  dim3 grid(static_cast<unsigned>(plan->Q), nchunk, nsplit);
  KTimer t(ctx, DALI_K_RANK_COUNT);
  rank_count_kernel<<<grid, kCountThreads, 0, ctx->stream>>>(dist, ld, g0, Gs, plan->d_off,
                                                            plan->d_nv, plan->d_gid, keys, counts,
                                                            nsplit);
  DALI_CUDA_OK(ctx, cudaGetLastError());
  return DALI_OK;
}

int launch_rank_finalize(dali_ctx *ctx, const dali_rank_plan *plan, const uint32_t *keys,
                         const int32_t *counts, int max_rank, int32_t *ranks_sorted, float *ap,
                         int32_t *first_rank, int32_t *cmc_cnt) {
  DALI_CUDA_OK(ctx, cudaMemsetAsync(cmc_cnt, 0, sizeof(int32_t) * (max_rank + 1), ctx->stream));
  if (plan->Q == 0) return DALI_OK;
  KTimer t(ctx, DALI_K_RANK_FINALIZE);
  rank_finalize_kernel<<<static_cast<unsigned>(plan->Q), kFinThreads, 0, ctx->stream>>>(
      plan->d_off, plan->d_nv, plan->d_gid, keys, counts, max_rank, ranks_sorted, ap, first_rank,
      cmc_cnt);
  DALI_CUDA_OK(ctx, cudaGetLastError());
  return DALI_OK;
}

}  // namespace dali
