// Sort-free ranking: positive-rank counting + CMC/AP epilogue.
//
// Replaces the arithmetic behind torchreid.metrics.evaluate_rank as called at
// validateModels.py:68-69, evaluate.py:312-313, evaluate_ensembled_models.py:324-325,
// evaluateCleanATModels.py:266-267.  Instead of argsort + gather + per-query loop, the
// 1-based kept rank of every valid positive p of query q is
//     1 + #{ j : (key(d[q,j]), j) <lex (key(d[q,p]), p) }  -  #{ junk u : (key_u,u) < (key_p,p) }
// which needs ONE streaming pass over the distance row (4 B / pair, HBM bound) and no
// label reads in the hot loop (junk items are all inside the query's match list).
#include <algorithm>
#include <cstdlib>

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

// ------------------------------ plan expansion ------------------------------
// One warp per query: walks the query's identity range of the gallery CSR and writes its match
// list -- valid positives (different camera) first in ascending gallery id, junk (same identity,
// same camera) after them -- plus the two counts.  Replaces the host loop that used to dominate
// the plan construction.
__global__ void __launch_bounds__(128)
plan_expand_kernel(int64_t Q, const int64_t *__restrict__ off, const int64_t *__restrict__ lo,
                   const int32_t *__restrict__ qcam, const int32_t *__restrict__ order,
                   const int32_t *__restrict__ gcam, int32_t *__restrict__ gid,
                   int32_t *__restrict__ nv_out, int32_t *__restrict__ njunk_out,
                   int32_t *__restrict__ slot_out) {
  const int64_t q = static_cast<int64_t>(blockIdx.x) * 4 + (threadIdx.x >> 5);
  if (q >= Q) return;
  const int lane = threadIdx.x & 31;
  const int64_t o = off[q];
  const int m = static_cast<int>(off[q + 1] - o);
  const int64_t l = lo[q];
  const int32_t qc = qcam[q];
  int nv = 0, nj = 0;
  for (int i0 = 0; i0 < m; i0 += 32) {
    const int i = i0 + lane;
    int32_t g = 0;
    bool valid = false, junk = false;
    if (i < m) {
      g = order[l + i];
      valid = gcam[g] != qc;
      junk = !valid;
    }
    const unsigned bv = __ballot_sync(0xffffffffu, valid);
    const unsigned bj = __ballot_sync(0xffffffffu, junk);
    const unsigned below = (1u << lane) - 1u;
    if (valid) {
      const int slot = nv + __popc(bv & below);
      gid[o + slot] = g;
      slot_out[o + i] = slot;
    }
    if (junk) {  // filled from the back
      const int slot = (m - 1) - (nj + __popc(bj & below));
      gid[o + slot] = g;
      slot_out[o + i] = slot;
    }
    nv += __popc(bv);
    nj += __popc(bj);
  }
  if (lane == 0) {
    nv_out[q] = nv;
    njunk_out[q] = nj;
  }
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

// ------------------------------ count, v2 -----------------------------------
// LUT-bucketed counting: per-element work independent of the number of positives.
//   1. the CTA sorts its (<= 254) thresholds T[0..n) by composite (key, gallery id);
//   2. a lookup table over NB uniform bins of the key range [key(T[0]), key(T[n-1])] gives, for a
//      bin, how many thresholds lie in lower bins and how many lie in this bin (almost always 0);
//   3. each streamed element finds bucket b = #{i : T[i] <= element} with one table lookup
//      (+ exact 64-bit compares only inside a threshold-holding bin) and bumps a PRIVATE 16-bit
//      counter cnt[b][thread] in shared memory: no atomics, at most 2-way bank conflicts;
//   4. counters are reduced per bucket; count_below(T[i]) = sum_{b <= i} hist[b].
constexpr int kV2Chunk = 254;  // thresholds per pass; buckets 0..n fit 8 bits

// bucket of an element whose table bin holds `ni` thresholds starting at T[base]: exact compares
__device__ __noinline__ uint32_t bucket_exact(const uint64_t *T, uint32_t base, uint32_t ni, uint64_t c) {
  uint32_t b = base;
  for (uint32_t j = 0; j < ni; ++j) b += (T[base + j] <= c) ? 1u : 0u;
  return b;
}


// kV2Threads = 256 for few thresholds; 128 when a query has many positives (DeepChange: ~120),
// where zeroing and reducing n x threads private counters is a large share of the CTA's work.
// BYTEC: 8-bit private counters (a thread must then see <= 255 elements: the launcher splits the
// row accordingly), rows of counters skewed by 4 bytes so that lanes hitting different buckets
// spread over the banks.  Halves the shared memory of a many-threshold CTA again.
// FUSED (single GPU, one CTA per query: no threshold chunks, no column splits): the CTA reads its
// thresholds straight from the matrix row (no gather kernel) and, having the counts of all its
// positives, finishes the query itself -- junk subtraction, kept ranks, torchreid's sequential
// float AP, first-match histogram (what rank_finalize_kernel does) -- so the rank stage of an
// evaluation is ONE launch instead of three.
struct FusedOut {
  int32_t *ranks_sorted;  // [M] kept ranks in rank order (py_f64 accumulation on the host)
  float *ap;              // [Q]
  int32_t *first_rank;    // [Q]
  int32_t *cmc_cnt;       // [max_rank + 1]
  int max_rank;
};

template <int LOG2NB, int kV2Threads, bool BYTEC, bool FUSED = false>
__global__ void __launch_bounds__(kV2Threads)
rank_count_v2_kernel(const float *__restrict__ dist, int64_t ld, int64_t g0, int64_t Gs,
                     const int64_t *__restrict__ off, const int32_t *__restrict__ nvalid,
                     const int32_t *__restrict__ gid, const uint32_t *__restrict__ keys,
                     int32_t *__restrict__ counts, int nsplit, int tchunk, FusedOut fo) {
  constexpr int NB = 1 << LOG2NB;
  extern __shared__ __align__(16) uint8_t smem_v2[];
  uint64_t *Tu = reinterpret_cast<uint64_t *>(smem_v2);            // [256] unsorted
  uint64_t *T = Tu + 256;                                          // [256] sorted
  uint32_t *hist = reinterpret_cast<uint32_t *>(T + 256);          // [256]
  uint16_t *orig = reinterpret_cast<uint16_t *>(hist + 256);       // [256]
  uint16_t *lut = orig + 256;                                      // [NB + 8], entry NB = "above all"
  uint16_t *cnt = lut + NB + 8;                                    // packed private counters, 32-bit words [bucket group][thread]

  const int64_t q = blockIdx.x;
  const int chunk = blockIdx.y;
  const int nv = nvalid[q];
  if (FUSED && nv == 0) {  // query without a valid match: not counted (torchreid skips it)
    if (threadIdx.x == 0) {
      fo.ap[q] = 0.f;
      fo.first_rank[q] = -1;
    }
    return;
  }
  if (chunk * tchunk >= nv) return;  // uniform exit
  const int n = min(tchunk, nv - chunk * tchunk);
  const int64_t o = off[q] + static_cast<int64_t>(chunk) * tchunk;
  const int tid = threadIdx.x;

  // 1. sort the thresholds by counting (composites are distinct: gallery ids differ)
  for (int t = tid; t < n; t += kV2Threads) {
    const uint32_t g = static_cast<uint32_t>(__ldg(gid + o + t));
    const uint32_t k = FUSED ? dist_key(__ldg(dist + q * ld + g)) : __ldg(keys + o + t);
    Tu[t] = composite(k, g);
  }
  __syncthreads();
  for (int t = tid; t < n; t += kV2Threads) {
    const uint64_t c = Tu[t];
    int pos = 0;
    for (int u = 0; u < n; ++u) pos += (Tu[u] < c) ? 1 : 0;
    T[pos] = c;
    orig[pos] = static_cast<uint16_t>(t);
  }
  __syncthreads();
  const uint32_t klo = static_cast<uint32_t>(T[0] >> 32);
  const uint32_t khi = static_cast<uint32_t>(T[n - 1] >> 32);
  const int bits = 32 - __clz(khi - klo);  // 0 when khi == klo
  const int sh = bits > LOG2NB ? bits - LOG2NB : 0;

  // 2. lookup table: lut[b] = (#thresholds in lower bins) | (#thresholds in bin b) << 8
  {
    constexpr int BPT = NB / kV2Threads;
    const int b0 = tid * BPT;
    int lo = 0, hi = n;  // first threshold whose bin >= b0
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      const int bm = static_cast<int>((static_cast<uint32_t>(T[mid] >> 32) - klo) >> sh);
      if (bm < b0) lo = mid + 1; else hi = mid;
    }
    int i = lo;
#pragma unroll 1
    for (int b = b0; b < b0 + BPT; ++b) {
      const int base = i;
      while (i < n && static_cast<int>((static_cast<uint32_t>(T[i] >> 32) - klo) >> sh) == b) ++i;
      lut[b] = static_cast<uint16_t>(base | ((i - base) << 8));
    }
    if (tid == 0) lut[NB] = static_cast<uint16_t>(n);  // keys beyond the table: above every threshold
  }
  // 3. zero the private counters (bucket n = "above every threshold" is never counted)
  {
    uint32_t *w = reinterpret_cast<uint32_t *>(cnt);
    const int words = BYTEC ? ((n + 4) >> 2) * kV2Threads : (n + 1) * (kV2Threads / 2);
    for (int i = tid; i < words; i += kV2Threads) w[i] = 0u;
  }
  __syncthreads();

  // 4. stream the row segment of this split
  const int64_t per = (((Gs + nsplit - 1) / nsplit) + 3) & ~int64_t(3);
  const int64_t c0r = per * static_cast<int64_t>(blockIdx.z);
  const int64_t c0 = c0r < Gs ? c0r : Gs;
  const int64_t c1 = (c0 + per) < Gs ? (c0 + per) : Gs;
  const float *row = dist + q * ld;
  const uint32_t gbase = static_cast<uint32_t>(g0);
  const uint32_t cnt_base = static_cast<uint32_t>(__cvta_generic_to_shared(cnt)) + static_cast<uint32_t>(tid << 2);
  uint16_t *mycnt = cnt + tid;

  // ~17 instructions per element, no data-dependent branch except the rare "bin holds
  // thresholds" call (a branch for keys above every threshold, or a cheaper path for
  // non-negative values, was measured 20-45 % slower: it serialises the eight visits).
  auto visit = [&](float d, uint32_t g) {
    // dist_key() without selects: d + 0 turns -0 into +0 and any NaN into 0x7FFFFFFF (PTX:
    // canonical NaN), whose image 0xFFFFFFFF is clamped to dist_key's NaN key 0xFFFFFFFE
    const uint32_t u = __float_as_uint(d + 0.0f);
    const uint32_t key = min(u ^ (static_cast<uint32_t>(static_cast<int32_t>(u) >> 31) | 0x80000000u),
                             0xFFFFFFFEu);
    const uint32_t dk = max(key, klo) - klo;
    const uint32_t e = lut[min(dk >> sh, static_cast<uint32_t>(NB))];
    uint32_t b = e & 0xFFu;
    if (e >= 0x100u) b = bucket_exact(T, b, e >> 8, composite(key, g));  // rare
    if (BYTEC) {
      // 8-bit private counters packed four to a 32-bit word, [bucket group][thread], bumped with one
      // fire-and-forget shared-memory add: a thread's words sit in its own bank, so the add is
      // conflict free whatever the bucket (a byte read-modify-write was 12 % slower at the
      // DeepChange shape).  A field cannot carry into its neighbour: the launcher bounds the
      // elements a thread sees to 254.  For the 16-bit counters of the few-threshold case the same
      // packed add measured 5-10 % SLOWER than the plain load / add / store below (Market shapes:
      // 0.110 vs 0.103 ms), so they keep it.
      asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(cnt_base + (((b >> 2) * kV2Threads) << 2)),
                   "r"(1u << ((b & 3u) << 3)) : "memory");
    } else {
      mycnt[b * kV2Threads] += 1;  // row n ("above every threshold") is never read
    }
  };

  const int mis = static_cast<int>((reinterpret_cast<uintptr_t>(row + c0) >> 2) & 3);
  int64_t head = (4 - mis) & 3;
  if (head > c1 - c0) head = c1 - c0;
  if (tid < head) visit(__ldg(row + c0 + tid), gbase + static_cast<uint32_t>(c0 + tid));
  const int64_t cv0 = c0 + head;
  const int64_t nvec = (c1 - cv0) >> 2;
  int64_t v = tid;
  for (; v + kV2Threads < nvec; v += 2 * kV2Threads) {
    const float4 x0 = ld_stream_f4(row + cv0 + 4 * v);
    const float4 x1 = ld_stream_f4(row + cv0 + 4 * (v + kV2Threads));
    const uint32_t ga = gbase + static_cast<uint32_t>(cv0 + 4 * v);
    const uint32_t gb = gbase + static_cast<uint32_t>(cv0 + 4 * (v + kV2Threads));
    visit(x0.x, ga); visit(x0.y, ga + 1); visit(x0.z, ga + 2); visit(x0.w, ga + 3);
    visit(x1.x, gb); visit(x1.y, gb + 1); visit(x1.z, gb + 2); visit(x1.w, gb + 3);
  }
  if (v < nvec) {
    const float4 x0 = ld_stream_f4(row + cv0 + 4 * v);
    const uint32_t ga = gbase + static_cast<uint32_t>(cv0 + 4 * v);
    visit(x0.x, ga); visit(x0.y, ga + 1); visit(x0.z, ga + 2); visit(x0.w, ga + 3);
  }
  const int64_t ct0 = cv0 + 4 * nvec;
  if (tid < c1 - ct0) visit(__ldg(row + ct0 + tid), gbase + static_cast<uint32_t>(ct0 + tid));
  __syncthreads();

  // 5. reduce the private counters per bucket (one warp per bucket, 8 counters per lane)
  {
    const int w = tid >> 5, l = tid & 31;
    const uint32_t *cw = reinterpret_cast<const uint32_t *>(cnt);
    for (int b = w; b < n; b += kV2Threads / 32) {
      uint32_t sum = 0;
      if (BYTEC) {
        const int sh = (b & 3) << 3;
        for (int e = l; e < kV2Threads; e += 32) sum += (cw[(b >> 2) * kV2Threads + e] >> sh) & 0xFFu;
      } else if (kV2Threads == 256) {
        const uint4 x = *reinterpret_cast<const uint4 *>(cnt + b * kV2Threads + l * 8);
        sum = (x.x & 0xFFFFu) + (x.x >> 16) + (x.y & 0xFFFFu) + (x.y >> 16) +
              (x.z & 0xFFFFu) + (x.z >> 16) + (x.w & 0xFFFFu) + (x.w >> 16);
      } else {
        const uint2 x = *reinterpret_cast<const uint2 *>(cnt + b * kV2Threads + l * 4);
        sum = (x.x & 0xFFFFu) + (x.x >> 16) + (x.y & 0xFFFFu) + (x.y >> 16);
      }
      sum = __reduce_add_sync(0xffffffffu, sum);
      if (l == 0) hist[b] = sum;
    }
  }
  __syncthreads();
  // 6. count_below(T[i]) = sum_{b <= i} hist[b]; scatter back to plan order
  if (!FUSED) {
    for (int t = tid; t < n; t += kV2Threads) {
      uint32_t below = 0;
      for (int b = 0; b <= t; ++b) below += hist[b];
      int32_t *dst = counts + o + orig[t];
      if (nsplit > 1) {
        if (below) atomicAdd(dst, static_cast<int32_t>(below));
      } else {
        *dst = static_cast<int32_t>(below);
      }
    }
    return;
  }
  // 7. FUSED: finish the query.  T[] is sorted by (key, gallery id) = by rank, so sorted index i
  // is the positive with the i-th best rank; junk matches (same identity, same camera) ahead of
  // it are subtracted, exactly as rank_finalize_kernel does.
  int32_t *s_rank = reinterpret_cast<int32_t *>(Tu);  // Tu is free after the sort
  const int m = static_cast<int>(off[q + 1] - off[q]);
  __syncthreads();
  for (int t = tid; t < n; t += kV2Threads) {
    uint32_t below = 0;
    for (int b = 0; b <= t; ++b) below += hist[b];
    const uint64_t ck = T[t];
    int below_junk = 0;
    for (int u = nv; u < m; ++u) {
      const uint32_t g = static_cast<uint32_t>(__ldg(gid + o + u));
      below_junk += composite(dist_key(__ldg(dist + q * ld + g)), g) < ck ? 1 : 0;
    }
    const int r = static_cast<int>(below) - below_junk + 1;  // 1-based rank among kept items
    s_rank[t] = r;
    fo.ranks_sorted[o + t] = r;
  }
  __syncthreads();
  if (tid == 0) {
    float s = 0.f;  // torchreid Cython accumulation: float running sum, each term formed in double
    for (int k = 1; k <= n; ++k)
      s = static_cast<float>(static_cast<double>(s) + static_cast<double>(k) / static_cast<double>(s_rank[k - 1]));
    fo.ap[q] = s / static_cast<float>(n);
    const int fr = s_rank[0];
    fo.first_rank[q] = fr;
    if (fr <= fo.max_rank) atomicAdd(fo.cmc_cnt + (fr - 1), 1);
    atomicAdd(fo.cmc_cnt + fo.max_rank, 1);  // num_valid_q
  }
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

// ------------------- fused distance + counting: threshold sort, prefix ---------------------
// (the counting itself is the kCount epilogue of distmat_umma2.cu)
namespace {
constexpr int kFzThreads = 128;  // >= the most valid positives of a query on the fused path

// One CTA per query: its valid positives' distances sorted ascending by (key, gallery id) -- a
// rank sort, the composites are distinct -- with the match-list slot of each; a positive whose
// distance is NaN or infinite raises the flag (the caller then takes the matrix path).
__global__ void __launch_bounds__(kFzThreads)
fused_sort_thresholds_kernel(const int64_t *__restrict__ off, const int32_t *__restrict__ nvalid,
                             const int32_t *__restrict__ gid, const uint32_t *__restrict__ keys,
                             float *__restrict__ sorted_thr, int32_t *__restrict__ sorted_slot,
                             int32_t *__restrict__ flag) {
  __shared__ uint64_t s_c[kFzThreads];
  const int64_t q = blockIdx.x;
  const int64_t o = off[q];
  const int nv = min(nvalid[q], kFzThreads);
  const int t = threadIdx.x;
  uint64_t c = 0;
  if (t < nv) {
    const uint32_t k = keys[o + t];
    c = composite(k, static_cast<uint32_t>(gid[o + t]));
    s_c[t] = c;
    if (k >= 0xFF800000u || k <= 0x007FFFFFu) *flag = 1;  // +inf, NaN / -inf
  }
  __syncthreads();
  if (t < nv) {
    int r = 0;
    for (int u = 0; u < nv; ++u) r += s_c[u] < c ? 1 : 0;
    sorted_thr[o + r] = key_to_dist(static_cast<uint32_t>(c >> 32));
    sorted_slot[o + r] = t;
  }
}

// counts[slot of the k-th smallest threshold] += sum of the buckets of its pass up to k
__global__ void __launch_bounds__(kFzThreads)
fused_prefix_kernel(const int64_t *__restrict__ off, const int32_t *__restrict__ nvalid,
                    const int32_t *__restrict__ hist, const int32_t *__restrict__ sorted_slot,
                    int32_t *__restrict__ counts, int pass) {
  const int64_t q = blockIdx.x;
  const int64_t o = off[q];
  const int nv = min(nvalid[q], kFzThreads);
  const int k = threadIdx.x;
  if (k >= nv) return;
  const int b0 = (k / pass) * pass;
  int sum = 0;
  for (int b = b0; b <= k; ++b) sum += hist[o + b];
  counts[o + sorted_slot[o + k]] += sum;
}
}  // namespace

int launch_fused_sort_thresholds(dali_ctx *ctx, const dali_rank_plan *plan, const uint32_t *keys,
                                 float *sorted_thr, int32_t *sorted_slot, int32_t *flag) {
  if (plan->Q == 0) return DALI_OK;
  KTimer t(ctx, DALI_K_RANK_GATHER);
  fused_sort_thresholds_kernel<<<static_cast<unsigned>(plan->Q), kFzThreads, 0, ctx->stream>>>(
      plan->d_off, plan->d_nv, plan->d_gid, keys, sorted_thr, sorted_slot, flag);
  DALI_CUDA_OK(ctx, cudaGetLastError());
  return DALI_OK;
}

int launch_fused_prefix(dali_ctx *ctx, const dali_rank_plan *plan, const int32_t *hist,
                        const int32_t *sorted_slot, int32_t *counts, int pass) {
  if (plan->Q == 0) return DALI_OK;
  KTimer t(ctx, DALI_K_RANK_COUNT);
  fused_prefix_kernel<<<static_cast<unsigned>(plan->Q), kFzThreads, 0, ctx->stream>>>(
      plan->d_off, plan->d_nv, hist, sorted_slot, counts, pass);
  DALI_CUDA_OK(ctx, cudaGetLastError());
  return DALI_OK;
}

int launch_plan_expand(dali_ctx *ctx, const dali_rank_plan *plan) {
  if (plan->Q == 0) return DALI_OK;
  plan_expand_kernel<<<static_cast<unsigned>((plan->Q + 3) / 4), 128, 0, ctx->stream>>>(
      plan->Q, plan->d_off, plan->d_lo, plan->d_qcam, plan->d_order, plan->d_gcam, plan->d_gid,
      plan->d_nv, plan->d_njunk, plan->d_slot);
  DALI_CUDA_OK(ctx, cudaGetLastError());
  ctx->launches++;
  return DALI_OK;
}

int launch_rank_gather(dali_ctx *ctx, const dali_rank_plan *plan, const float *dist, int64_t ld,
                       int64_t g0, int64_t Gs, uint32_t *keys) {
  if (plan->M == 0 || plan->Q == 0) return DALI_OK;
  KTimer t(ctx, DALI_K_RANK_GATHER);
  rank_gather_kernel<<<static_cast<unsigned>(plan->Q), 128, 0, ctx->stream>>>(
      dist, ld, g0, Gs, plan->d_off, plan->d_gid, keys);
  DALI_CUDA_OK(ctx, cudaGetLastError());
  return DALI_OK;
}

static size_t v2_smem_bytes(int log2nb, int nbuckets, int threads, bool bytec = false) {
  // thresholds (2 x 256 x 8), hist, orig, table, then the packed private counters
  return 256 * 8 * 2 + 256 * 4 + 256 * 2 + ((size_t(1) << log2nb) + 8) * 2 +
         (bytec ? size_t((nbuckets + 4) >> 2) * threads * 4 : size_t(nbuckets + 1) * threads * 2) + 16;
}

template <int LOG2NB, int THREADS, bool BYTEC = false>
static int launch_v2(dali_ctx *ctx, dim3 grid, size_t smem, const dali_rank_plan *plan, const float *dist,
                     int64_t ld, int64_t g0, int64_t Gs, const uint32_t *keys, int32_t *counts,
                     int nsplit, int tchunk) {
  if (int rc = ensure_dyn_smem(ctx, reinterpret_cast<const void *>(&rank_count_v2_kernel<LOG2NB, THREADS, BYTEC>),
                               smem))
    return rc;
  rank_count_v2_kernel<LOG2NB, THREADS, BYTEC><<<grid, THREADS, smem, ctx->stream>>>(
      dist, ld, g0, Gs, plan->d_off, plan->d_nv, plan->d_gid, keys, counts, nsplit, tchunk, FusedOut{});
  return DALI_OK;
}

int launch_rank_count(dali_ctx *ctx, const dali_rank_plan *plan, const float *dist, int64_t ld,
                      int64_t g0, int64_t Gs, const uint32_t *keys, int32_t *counts) {
  if (plan->M == 0 || plan->Q == 0) return DALI_OK;
  DALI_CUDA_OK(ctx, cudaMemsetAsync(counts, 0, sizeof(int32_t) * plan->M, ctx->stream));
  if (plan->max_nv == 0 || Gs == 0) return DALI_OK;
  static const bool use_v1 = getenv("DALI_RANK_V1") != nullptr;  // debugging cross-check only
  static const char *env_c = getenv("DALI_RANK_CHUNK");
  static const char *env_t = getenv("DALI_RANK_THREADS");
  // thresholds per CTA: a query with many positives is split over several CTAs that each stream
  // the row (re-reads hit L2) -- the private counters of one CTA (n x threads x 2 B) are what
  // limits the number of resident warps
  int tchunk = kV2Chunk;
  if (env_c) tchunk = std::max(8, std::min(kV2Chunk, atoi(env_c)));
  const int per_cta = use_v1 ? kChunk : tchunk;
  const int nchunk = (plan->max_nv + per_cta - 1) / per_cta;
  if (nchunk > 65535) return set_err(ctx, DALI_ERR_UNSUPPORTED, "too many positives for one query");
  // enough CTAs for >= 4 per SM; never split a row below 4096 columns; a v2 CTA counts in
  // 16-bit private counters, so it may see at most 65535 * 128 columns
  int64_t want = (4ll * ctx->num_sms + plan->Q * nchunk - 1) / (plan->Q * nchunk);
  const int64_t max_split = (Gs + 4095) / 4096;
  const int64_t min_split = (Gs + (4ll << 20) - 1) / (4ll << 20);
  int64_t ns = std::max<int64_t>(1, std::min(want, max_split));
  ns = std::max(ns, min_split);
  // many thresholds per query: byte counters, so a thread may see at most 255 elements
  static const char *env_b = getenv("DALI_RANK_BYTE");
  const bool bytec = !use_v1 && std::min(plan->max_nv, tchunk) > 64 && !(env_b && atoi(env_b) == 0);
  if (bytec) ns = std::max<int64_t>(ns, (Gs + 250ll * 128 - 1) / (250ll * 128));
  if (ns > 65535) return set_err(ctx, DALI_ERR_UNSUPPORTED, "slab too wide for one launch");
  const int nsplit = static_cast<int>(ns);
  dim3 grid(static_cast<unsigned>(plan->Q), nchunk, nsplit);
  KTimer t(ctx, DALI_K_RANK_COUNT);
  if (use_v1) {
    rank_count_kernel<<<grid, kCountThreads, 0, ctx->stream>>>(dist, ld, g0, Gs, plan->d_off,
                                                              plan->d_nv, plan->d_gid, keys, counts,
                                                              nsplit);
  } else {
    const int nb = std::min(plan->max_nv, tchunk);
    // finer table and fewer private counter copies when a query has many thresholds
    int rc;
    const int force = env_t ? atoi(env_t) : 0;
    if (force == 1128) {
      rc = launch_v2<11, 128>(ctx, grid, v2_smem_bytes(11, nb, 128), plan, dist, ld, g0, Gs, keys, counts, nsplit, tchunk);
    } else if (bytec && force == 0) {
      rc = launch_v2<12, 128, true>(ctx, grid, v2_smem_bytes(12, nb, 128, true), plan, dist, ld, g0, Gs, keys, counts,
                                    nsplit, tchunk);
    } else if (force == 128 || (force == 0 && nb > 64)) {
      rc = launch_v2<12, 128>(ctx, grid, v2_smem_bytes(12, nb, 128), plan, dist, ld, g0, Gs, keys, counts, nsplit, tchunk);
    } else {
      rc = launch_v2<11, 256>(ctx, grid, v2_smem_bytes(11, nb, 256), plan, dist, ld, g0, Gs, keys, counts, nsplit, tchunk);
    }
    if (rc) return rc;
  }
  DALI_CUDA_OK(ctx, cudaGetLastError());
  return DALI_OK;
}

// Whole rank stage of a single-GPU evaluation in one launch; *done = 0 (nothing launched) when the
// shape needs threshold chunks, column splits or the many-threshold variants.
int launch_rank_fused(dali_ctx *ctx, const dali_rank_plan *plan, const float *dist, int64_t ld,
                      int max_rank, int32_t *ranks_sorted, float *ap, int32_t *first_rank,
                      int32_t *cmc_cnt, int *done) {
  *done = 0;
  static const char *env = getenv("DALI_RANK_FUSED");
  if (env && atoi(env) == 0) return DALI_OK;
  const int64_t G = plan->G;
  if (plan->Q == 0 || plan->M == 0 || G == 0 || plan->max_nv == 0 || plan->max_nv > kV2Chunk) return DALI_OK;
  if (getenv("DALI_RANK_V1") || getenv("DALI_RANK_CHUNK") || getenv("DALI_RANK_THREADS")) return DALI_OK;
  // the same split rule as launch_rank_count: only the one-CTA-per-query case is fused
  const int64_t want = (4ll * ctx->num_sms + plan->Q - 1) / plan->Q;
  const int64_t max_split = (G + 4095) / 4096;
  const int64_t min_split = (G + (4ll << 20) - 1) / (4ll << 20);
  if (std::max<int64_t>(std::max<int64_t>(1, std::min(want, max_split)), min_split) != 1) return DALI_OK;
  // many thresholds per query (DeepChange: ~120): byte counters, which limit a thread to 255
  // elements.  A thread visits ceil(nvec / 256) float4 vectors plus at most one head and one tail
  // element, so rows of up to 250 * 256 = 64000 columns are safe (63 vectors + 2 = 254 elements);
  // wider rows take the split launch, which bounds its segments the same way.
  const bool bytec = plan->max_nv > 64;
  static const char *env_b = getenv("DALI_RANK_FUSED_BYTE");
  if (bytec && (G > 250ll * 256 || (env_b && atoi(env_b) == 0))) return DALI_OK;
  DALI_CUDA_OK(ctx, cudaMemsetAsync(cmc_cnt, 0, sizeof(int32_t) * (max_rank + 1), ctx->stream));
  FusedOut fo{ranks_sorted, ap, first_rank, cmc_cnt, max_rank};
  KTimer t(ctx, DALI_K_RANK_COUNT);
  const dim3 grid(static_cast<unsigned>(plan->Q), 1, 1);
  if (bytec) {
    const size_t smem = v2_smem_bytes(12, plan->max_nv, 256, true);
    if (int rc = ensure_dyn_smem(ctx, reinterpret_cast<const void *>(&rank_count_v2_kernel<12, 256, true, true>), smem))
      return rc;
    rank_count_v2_kernel<12, 256, true, true><<<grid, 256, smem, ctx->stream>>>(
        dist, ld, 0, G, plan->d_off, plan->d_nv, plan->d_gid, nullptr, nullptr, 1, kV2Chunk, fo);
  } else {
    const size_t smem = v2_smem_bytes(11, plan->max_nv, 256);
    if (int rc = ensure_dyn_smem(ctx, reinterpret_cast<const void *>(&rank_count_v2_kernel<11, 256, false, true>), smem))
      return rc;
    rank_count_v2_kernel<11, 256, false, true><<<grid, 256, smem, ctx->stream>>>(
        dist, ld, 0, G, plan->d_off, plan->d_nv, plan->d_gid, nullptr, nullptr, 1, kV2Chunk, fo);
  }
  DALI_CUDA_OK(ctx, cudaGetLastError());
  *done = 1;
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
