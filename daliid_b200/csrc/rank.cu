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
#include <type_traits>

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

// ------------------------------ count, v3 -----------------------------------
// Register-resident counting for queries with few positives (<= 64 thresholds per CTA; the Market
// shapes).  What bounded v2 was the shared-memory pipe: a random 16-bit table lookup (~3.4
// wavefronts per warp and element) plus the read and the write of a private counter (2 + 2).  v3
// keeps the lookup and drops the counters:
//   * the table bin of an element is ONE fused multiply-add on the fp32 distance: t = d * (-scale)
//     + C lands in [2^23, 2^24), where the low mantissa bits ARE the rounded bin, two FMNMX clamp it
//     (and send NaN above every threshold), one LEA forms the shared-memory address.  The map is
//     monotone in d and the thresholds are binned by the same expression, so an element in a bin
//     without thresholds is ordered exactly against all of them; an element that shares its bin with
//     a threshold takes the exact path (64-bit composite compares, ties by gallery id);
//   * the table entry is base = #{thresholds <= element}; the element must be counted for the
//     thresholds base..n-1, i.e. it contributes the bit mask ~0 << base.  The masks of 8 elements are
//     added into BIT-SLICED per-thread counters (plane k holds bit k of 32 / 64 counters) with a
//     carry-save adder tree: 14 LOP3 + 2 per further plane for 8 elements and 32 thresholds, all in
//     registers, no shared-memory traffic;
//   * at the end a warp transposes its planes (lane i then holds the bits of threshold i of all 32
//     lanes: popc) and adds 32 / 64 totals to the CTA's histogram; hist[i] is count_below(T[i])
//     directly, no prefix sum.
// A CTA whose thresholds are not all finite (NaN rows, infinite distances) counts its row with the
// generic compare loop instead: same results, slow, rare.
constexpr int kV3Chunk = 64;
constexpr int kV3MaxChunks = 1;  // threshold chunks per query the split launch takes before v2's byte counters
#ifndef DALI_V3_DEPTH
#define DALI_V3_DEPTH 1  // ring slots (two float4 per thread each): 1 measured best (0.0641 / 0.0655 / 0.0696 ms for 1 / 2 / 3)
#endif
#ifndef DALI_V3B_LOG2NB
#define DALI_V3B_LOG2NB 12  // table bins of the many-threshold (CAP = 256) variant
#endif
#ifndef DALI_V3_SPREAD_TOTALS
#define DALI_V3_SPREAD_TOTALS 1  // totals: plane transposes spread over the CTA's warps (0: one warp per word)
#endif
#ifndef DALI_V3_MINB
#define DALI_V3_MINB 8  // resident 256-thread CTAs per SM the tight variant is compiled for (32 registers)
#endif
constexpr uint32_t kV3Magic = 0x4B000000u;  // 2^23 as fp32 bits

__device__ __forceinline__ uint32_t lop3_xor3(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ uint32_t lop3_maj(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ uint32_t shl_clamp(uint32_t a, uint32_t s) {  // s >= 32 -> 0
  uint32_t d;
  asm("shl.b32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(s));
  return d;
}

// 32 x 32 bit transpose across the lanes of a warp: on return bit j of lane i = bit i of lane j
__device__ __forceinline__ uint32_t warp_bit_transpose(uint32_t x, int lane) {
#pragma unroll
  for (int j = 16; j >= 1; j >>= 1) {
    const uint32_t m = j == 16 ? 0x0000FFFFu : j == 8 ? 0x00FF00FFu : j == 4 ? 0x0F0F0F0Fu
                     : j == 2 ? 0x33333333u : 0x55555555u;
    const uint32_t y = __shfl_xor_sync(0xffffffffu, x, j);
    x = (lane & j) ? ((x & ~m) | ((y >> j) & m)) : ((x & m) | ((y & m) << j));
  }
  return x;
}

template <int NW, int PL>
struct SlicedCounters {
  uint32_t p[PL][NW];
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int k = 0; k < PL; ++k)
#pragma unroll
      for (int w = 0; w < NW; ++w) p[k][w] = 0u;
  }
  // adds eight 0/1 vectors (one bit per threshold)
  __device__ __forceinline__ void add8(const uint32_t (&m)[8][NW]) {
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      uint32_t ones = p[0][w], twos = p[1][w], fours = p[2][w];
      const uint32_t tA = lop3_maj(ones, m[0][w], m[1][w]);
      ones = lop3_xor3(ones, m[0][w], m[1][w]);
      const uint32_t tB = lop3_maj(ones, m[2][w], m[3][w]);
      ones = lop3_xor3(ones, m[2][w], m[3][w]);
      const uint32_t fA = lop3_maj(twos, tA, tB);
      twos = lop3_xor3(twos, tA, tB);
      const uint32_t tC = lop3_maj(ones, m[4][w], m[5][w]);
      ones = lop3_xor3(ones, m[4][w], m[5][w]);
      const uint32_t tD = lop3_maj(ones, m[6][w], m[7][w]);
      ones = lop3_xor3(ones, m[6][w], m[7][w]);
      const uint32_t fB = lop3_maj(twos, tC, tD);
      twos = lop3_xor3(twos, tC, tD);
      uint32_t carry = lop3_maj(fours, fA, fB);
      fours = lop3_xor3(fours, fA, fB);
      p[0][w] = ones; p[1][w] = twos; p[2][w] = fours;
#pragma unroll
      for (int k = 3; k < PL; ++k) {
        const uint32_t t = p[k][w] & carry;
        p[k][w] ^= carry;
        carry = t;
      }
    }
  }
};

// CAP = thresholds one CTA takes: 64 -> the bit-sliced register counters described above; 256 (queries
// with many positives: DeepChange's ~120, up to 254) -> v3's front end (FMA bins, 16-byte table fill,
// float exact path, cp.async ring) in front of v2's 8-bit private counters, packed four to a word in
// DYNAMIC shared memory [bucket group][thread] and bumped with one conflict-free red.shared.add (a
// thread must then see at most 255 elements: the launcher checks).
template <int LOG2NB, int THREADS, int PL, bool FUSED, bool TIGHT = false, int CAP = 64>
__global__ void __launch_bounds__(THREADS, TIGHT ? DALI_V3_MINB * 256 / THREADS : 1)
rank_count_v3_kernel(const float *__restrict__ dist, int64_t ld, int64_t g0, int64_t Gs,
                     const int64_t *__restrict__ off, const int32_t *__restrict__ nvalid,
                     const int32_t *__restrict__ gid, const uint32_t *__restrict__ keys,
                     int32_t *__restrict__ counts, int nsplit, FusedOut fo) {
  constexpr int NB = 1 << LOG2NB;  // bins 0 .. NB; thresholds fall into 2 .. NB-2
  constexpr int NE = NB + 8;       // table entries (a multiple of 8)
  constexpr int kJunkCap = 64;
  constexpr int kCap = CAP;  // thresholds this CTA takes
  constexpr bool BYTES = CAP > 64;
  static_assert(THREADS >= CAP && THREADS % (CAP < THREADS ? CAP : THREADS) == 0, "one thread per threshold at least");
  __shared__ uint64_t Tu[kCap + kJunkCap];  // unsorted thresholds, then the junk composites
  __shared__ uint64_t T[kCap + 1];          // sorted, T[n] = sentinel above every composite
  __shared__ float Tf[kCap + 1];            // sorted thresholds as distances, Tf[n] = +inf
  __shared__ uint32_t Tg[kCap + 1];         // their gallery ids
  __shared__ uint32_t hist[kCap + 1];
  __shared__ uint16_t orig[kCap];
  __shared__ uint16_t tb[kCap + 2];
  __shared__ double s_term[FUSED ? kCap : 1];
  __shared__ __align__(16) uint16_t lut[NE];
  __shared__ uint32_t bh2[BYTES ? CAP + 4 : 1];  // byte-counter variant: elements per bucket
  // the row is streamed through a per-thread ring in shared memory with cp.async: the loads of the
  // first DEPTH iterations are in flight while the prologue runs, and no registers are held for them
  constexpr int DEPTH = DALI_V3_DEPTH;
  __shared__ float4 ring[DEPTH][2][THREADS];
  constexpr int kPartPlanes = PL + 3;
  __shared__ uint32_t part[BYTES ? 1 : 2 * kPartPlanes * 32];  // [threshold word][plane][partial sum]

  const int64_t q = blockIdx.x;
  const int chunk = FUSED ? 0 : blockIdx.y;  // FUSED: one CTA per query, the whole row, all thresholds
  const int tid = threadIdx.x;
  const float *row = dist + q * ld;

  // column range of this split, in multiples of 4 columns; [cv0, cv0 + 4 nvec) is its 16-byte
  // aligned part, dealt out as float4 vectors: iteration `it` gives a thread the vectors
  // tid + 2 it THREADS and that + THREADS.  Vectors beyond the range read as NaN, which sorts above
  // every threshold and so counts for none.
  const int64_t per = (FUSED || nsplit == 1) ? ((Gs + 3) & ~int64_t(3)) : ((((Gs + nsplit - 1) / nsplit) + 3) & ~int64_t(3));
  const int64_t c0r = FUSED ? 0 : per * static_cast<int64_t>(blockIdx.z);
  const int64_t c0 = c0r < Gs ? c0r : Gs;
  const int64_t c1 = (FUSED || (c0 + per) >= Gs) ? Gs : (c0 + per);
  const int mis = static_cast<int>((reinterpret_cast<uintptr_t>(row + c0) >> 2) & 3);
  int head = (4 - mis) & 3;
  if (head > c1 - c0) head = static_cast<int>(c1 - c0);
  const int64_t cv0 = c0 + head;
  const int nvec = static_cast<int>((c1 - cv0) >> 2);
  const int tail = static_cast<int>(c1 - cv0) & 3;
  const int niter = (nvec + 2 * THREADS - 1) / (2 * THREADS);
  const float qnan = __int_as_float(0x7FFFFFFF);
  const float *pv = row + cv0 + 4 * tid;  // this thread's next pair of vectors
  const int nfull = nvec / (2 * THREADS);  // iterations in which every thread has both its vectors
  const uint32_t ring_s = static_cast<uint32_t>(__cvta_generic_to_shared(&ring[0][0][tid]));
  // requests the thread's two vectors of iteration `itx` into ring[slot]; always one commit group
  auto issue = [&](int slot, int itx) {
    if (itx < nfull) {  // uniform
      const uint32_t dst = ring_s + static_cast<uint32_t>(slot * 2 * THREADS * 16);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(pv) : "memory");
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + THREADS * 16), "l"(pv + 4 * THREADS) : "memory");
    } else if (itx < niter) {
      const int rem = nvec - tid - itx * (2 * THREADS);  // vectors left from pv on, in steps of THREADS
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const uint32_t dst = ring_s + static_cast<uint32_t>((slot * 2 + j) * THREADS * 16);
        if (rem > j * THREADS)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(pv + 4 * j * THREADS) : "memory");
        else
          asm volatile("st.shared.v4.f32 [%0], {%1, %1, %1, %1};" ::"r"(dst), "f"(qnan) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    pv += 8 * THREADS;
  };
#pragma unroll
  for (int sl = 0; sl < DEPTH; ++sl) issue(sl, sl);

  const int nv = __ldg(nvalid + q);
  const int64_t off_q = __ldg(off + q), off_q1 = __ldg(off + q + 1);  // with nv: one latency, not two
  if (FUSED && nv == 0) {
    if (tid == 0) {
      fo.ap[q] = 0.f;
      fo.first_rank[q] = -1;
    }
    return;
  }
  if (chunk * kCap >= nv) return;  // uniform exit
  const int n = min(kCap, nv - chunk * kCap);
  const int64_t o = off_q + static_cast<int64_t>(chunk) * kCap;
  // FUSED: the junk matches (same identity, same camera) follow the valid ones in the match list
  const int m = FUSED ? static_cast<int>(off_q1 - off_q) : n;
  const int nj = min(m - n, kJunkCap);  // staged; more than that are read from global memory later

  // 1. thresholds, sorted by counting (composites are distinct: gallery ids differ)
  for (int t = tid; t < n + nj; t += THREADS) {
    const uint32_t g = static_cast<uint32_t>(__ldg(gid + o + t));
    const uint32_t k = FUSED ? dist_key(__ldg(row + g)) : __ldg(keys + o + t);
    Tu[t] = composite(k, g);
  }
  for (int t = tid; t <= kCap; t += THREADS) hist[t] = 0u;
  __syncthreads();
  {
    constexpr int TPT = THREADS / kCap;  // threads per threshold (4 or 2), adjacent lanes
    const int i = tid / TPT, part = tid % TPT;
    const uint64_t c = Tu[i < n ? i : 0];
    int pos = 0;
    if (i < n)  // (whole warps beyond the thresholds skip the loop)
      for (int u = part; u < n; u += TPT) pos += (Tu[u] < c) ? 1 : 0;
#pragma unroll
    for (int x = 1; x < TPT; x <<= 1) pos += __shfl_xor_sync(0xffffffffu, pos, x);
    if (i < n && part == 0) {
      T[pos] = c;
      Tf[pos] = key_to_dist(static_cast<uint32_t>(c >> 32));
      Tg[pos] = static_cast<uint32_t>(c);
      orig[pos] = static_cast<uint16_t>(i);
    }
  }
  if (tid == 0) {
    T[n] = ~0ull;
    Tf[n] = __int_as_float(0x7F800000);
    Tg[n] = 0xFFFFFFFFu;
  }
  __syncthreads();

  // 2. the bin map: u = sat(C - d s) in [0, 1] (NaN -> 0: above every threshold), bin = round(u NB)
  // read off the mantissa of u NB + 2^23.  hi -> bin 2, lo -> bin NB - 2; monotone in d whatever s
  // is, so only the bins' fineness depends on the approximate reciprocals.  s is capped so that C
  // keeps a few bits below one bin (thresholds that nearly coincide relative to their size then
  // share bins, which is slower, not wrong).
  const uint32_t klo = static_cast<uint32_t>(T[0] >> 32);
  const uint32_t khi = static_cast<uint32_t>(T[n - 1] >> 32);
  const float lo = Tf[0], hi = Tf[n - 1];
  float sc = fminf(__fdividef(static_cast<float>(NB - 4) / static_cast<float>(NB), hi - lo),
                   __fdividef(524288.0f / static_cast<float>(NB), fabsf(hi)));
  sc = fminf(fmaxf(sc, 1.0e-30f), 1.0e30f);
  const float nsc = -sc;
  const float C = fmaf(hi, sc, 2.0f / static_cast<float>(NB));
  // all thresholds finite (key images of -inf / +inf bound the finite range)
  const bool fast = klo > 0x007FFFFFu && khi < 0xFF800000u;
  auto bin_bits = [&](float d) -> uint32_t {  // fp32 bits of 2^23 + bin
    float u;
    asm("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(u) : "f"(d), "f"(nsc), "f"(C));
    return __float_as_uint(fmaf(u, static_cast<float>(NB), 8388608.0f));
  };
  const uint32_t gbase = static_cast<uint32_t>(g0);

  if (fast) {
    // 3. table: entry(b) = #{thresholds in higher bins (= smaller distances)}, | 0x8000 if bin b
    // holds thresholds itself.  tb[i] = bin of T[i], non-increasing; tb[-1] := NE, tb[n] := -1.
    uint16_t *tbs = tb + 1;
    if (tid < n) tbs[tid] = static_cast<uint16_t>(bin_bits(Tf[tid]) - kV3Magic);
    __syncthreads();
    // (a) every block of eight entries as if no threshold fell inside it: #{tb >= block start}
    for (int blk = tid; blk < NE / 8; blk += THREADS) {
      const int b0 = blk * 8;
      int l = 0, h = n;  // first index with tb < b0
      while (l < h) {
        const int mid = (l + h) >> 1;
        if (static_cast<int>(tbs[mid]) >= b0) l = mid + 1; else h = mid;
      }
      const uint32_t w = static_cast<uint32_t>(l) * 0x00010001u;
      *reinterpret_cast<uint4 *>(lut + b0) = make_uint4(w, w, w, w);
    }
    __syncthreads();
    // (b) the blocks that do hold thresholds, patched by the thresholds' threads: threshold i (the
    // first of a run sharing its bin) writes its bin and the pure bins above it up to the next
    // threshold's bin or the block's end; the last threshold of a block also the bins below it
    if (tid < n) {
      const int i = tid;
      const int b = tbs[i];
      const int up = i > 0 ? static_cast<int>(tbs[i - 1]) : NE;        // bin of the next smaller distance
      const int dn = i + 1 < n ? static_cast<int>(tbs[i + 1]) : -1;    // bin of the next larger distance
      const int bs = b & ~7, be = bs + 8;
      if (up != b) {
        lut[b] = static_cast<uint16_t>(i | 0x8000);
        for (int x = b + 1; x < min(up, be); ++x) lut[x] = static_cast<uint16_t>(i);
      }
      if (dn < bs)
        for (int x = bs; x < b; ++x) lut[x] = static_cast<uint16_t>(i + 1);
    }
    __syncthreads();

    uint32_t lut_s = static_cast<uint32_t>(__cvta_generic_to_shared(lut)) - (kV3Magic << 1);
    asm volatile("mov.u32 %0, %0;" : "+r"(lut_s));  // keep the constant folded into the base register
    // table entry of an element: #{thresholds <= element}, bit 15 = "exact compare needed"
    auto entry_of = [&](float d) -> uint32_t {
      uint32_t e;
      asm volatile("ld.shared.u16 %0, [%1];" : "=r"(e) : "r"(lut_s + (bin_bits(d) << 1)));
      return e;
    };
    // rare: the bin holds thresholds.  Thresholds and the element are finite here (NaN and the
    // infinities never share a bin with a finite threshold), so (distance, gallery id) compares as
    // floats; -0 == +0 as in dist_key.
    auto exact = [&](uint32_t e, float d, uint32_t g) -> uint32_t {
      uint32_t b = e & 0x7FFFu;
      while (true) {
        const float t = Tf[b];
        if (t < d || (t == d && Tg[b] <= g)) ++b; else break;
      }
      return b;
    };

    auto run = [&](auto nw_tag) {
      constexpr int NW = decltype(nw_tag)::value;
      SlicedCounters<NW, PL> scnt;
      scnt.clear();
      uint32_t mk[8][NW];
      auto to_mask = [&](int s, uint32_t b) {
        mk[s][0] = shl_clamp(0xFFFFFFFFu, b);
        if (NW == 2) mk[s][NW - 1] = shl_clamp(0xFFFFFFFFu, max(b, 32u) - 32u);
      };
      // eight elements: the vector at gallery id ga and the one THREADS vectors further on
      auto group = [&](const float4 &x0, const float4 &x1, uint32_t ga) {
        const float d[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
        uint32_t e[8];
#pragma unroll
        for (int s = 0; s < 8; ++s) e[s] = entry_of(d[s]);
        if ((e[0] | e[1] | e[2] | e[3] | e[4] | e[5] | e[6] | e[7]) & 0x8000u) {
#pragma unroll
          for (int s = 0; s < 8; ++s)
            if (e[s] & 0x8000u) e[s] = exact(e[s], d[s], ga + (s < 4 ? 0u : 4u * THREADS) + (s & 3));
        }
#pragma unroll
        for (int s = 0; s < 8; ++s) to_mask(s, e[s]);
        scnt.add8(mk);
      };
      uint32_t ga = gbase + static_cast<uint32_t>(cv0) + 4u * tid;
      int slot = 0;
#pragma unroll 1
      for (int it = 0; it < niter; ++it) {
        asm volatile("cp.async.wait_group %0;" ::"n"(DEPTH - 1) : "memory");
        float4 x0, x1;
        const uint32_t src = ring_s + static_cast<uint32_t>(slot * 2 * THREADS * 16);
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x0.x), "=f"(x0.y), "=f"(x0.z), "=f"(x0.w) : "r"(src) : "memory");
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x1.x), "=f"(x1.y), "=f"(x1.z), "=f"(x1.w) : "r"(src + THREADS * 16) : "memory");
        issue(slot, it + DEPTH);  // the slot is this thread's own: no barrier between its read and its refill
        group(x0, x1, ga);
        ga += 8u * THREADS;
        slot = slot + 1 == DEPTH ? 0 : slot + 1;
      }
      // the unaligned head and the tail of the range (<= 3 columns each, one per thread)
      if (head | tail) {  // uniform
#pragma unroll
        for (int s = 0; s < 8; ++s)
#pragma unroll
          for (int w = 0; w < NW; ++w) mk[s][w] = 0u;
        auto put = [&](int s, float d, uint32_t g) {
          uint32_t e = entry_of(d);
          if (e & 0x8000u) e = exact(e, d, g);
          to_mask(s, e);
        };
        if (tid < head) put(0, __ldg(row + c0 + tid), gbase + static_cast<uint32_t>(c0 + tid));
        const int64_t ct0 = cv0 + 4 * static_cast<int64_t>(nvec);
        if (tid < tail) put(1, __ldg(row + ct0 + tid), gbase + static_cast<uint32_t>(ct0 + tid));
        scnt.add8(mk);
      }

      // 4. totals.  Neighbouring lanes first add their counters in bit-sliced form (a ripple-carry
      // adder per stage: 2 LOP3 and a shuffle per plane) until the CTA is down to 32 partial sums;
      // those go to shared memory, and ONE warp per 32 thresholds transposes them (lane i then holds
      // the bits of threshold i of all 32 partials: popc).  Transposing every warp's seven planes
      // directly cost 300 instructions per warp -- a sixth of the kernel at the Market row length.
      constexpr int ST = THREADS == 256 ? 3 : 2;  // (THREADS / 32) * (32 >> ST) == 32 partials
      constexpr int PLR = PL + ST;
      const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
      for (int w = 0; w < NW; ++w) {
        uint32_t r[PLR];
#pragma unroll
        for (int k2 = 0; k2 < PLR; ++k2) r[k2] = k2 < PL ? scnt.p[k2][w] : 0u;
#pragma unroll
        for (int st = 0; st < ST; ++st) {
          uint32_t carry = 0u;
#pragma unroll
          for (int k2 = 0; k2 < PL + st; ++k2) {
            const uint32_t o2 = __shfl_xor_sync(0xffffffffu, r[k2], 1 << st);
            const uint32_t sum = lop3_xor3(r[k2], o2, carry);
            carry = lop3_maj(r[k2], o2, carry);
            r[k2] = sum;
          }
          r[PL + st] = carry;
        }
        if ((lane & ((1 << ST) - 1)) == 0) {
#pragma unroll
          for (int k2 = 0; k2 < PLR; ++k2) part[(w * kPartPlanes + k2) * 32 + warp * (32 >> ST) + (lane >> ST)] = r[k2];
        }
      }
      __syncthreads();
#if DALI_V3_SPREAD_TOTALS
      // the NW * PLR plane transposes are dealt out to all warps (one warp doing them in a row is a
      // chain of 5 PLR dependent shuffles while the others wait at the next barrier); hist[] is zero
#pragma unroll 1
      for (int job = warp; job < NW * PLR; job += THREADS / 32) {
        const int w = job / PLR, k2 = job - w * PLR;
        const uint32_t v = static_cast<uint32_t>(__popc(warp_bit_transpose(part[(w * kPartPlanes + k2) * 32 + lane], lane))) << k2;
        if (v) atomicAdd(&hist[32 * w + lane], v);
      }
#else
      if (warp < NW) {
        uint32_t total = 0;
#pragma unroll
        for (int k2 = 0; k2 < PLR; ++k2)
          total += static_cast<uint32_t>(__popc(warp_bit_transpose(part[(warp * kPartPlanes + k2) * 32 + lane], lane))) << k2;
        hist[32 * warp + lane] = total;
      }
#endif
    };
    // many thresholds: 8-bit private counters in shared memory, bucket b = #{thresholds <= element}
    auto run_bytes = [&]() {
      extern __shared__ __align__(16) uint32_t cnt_dyn[];
      const int groups = (n + 4) >> 2;  // buckets 0 .. n
      for (int i = tid; i < groups * THREADS; i += THREADS) cnt_dyn[i] = 0u;
      __syncthreads();
      const uint32_t cnt_s = static_cast<uint32_t>(__cvta_generic_to_shared(cnt_dyn)) + static_cast<uint32_t>(tid << 2);
      auto bump = [&](uint32_t b) {
        asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(cnt_s + (((b >> 2) * THREADS) << 2)),
                     "r"(1u << ((b & 3u) << 3)) : "memory");
      };
      uint32_t ga = gbase + static_cast<uint32_t>(cv0) + 4u * tid;
      int slot = 0;
#pragma unroll 1
      for (int it = 0; it < niter; ++it) {
        asm volatile("cp.async.wait_group %0;" ::"n"(DEPTH - 1) : "memory");
        float4 x0, x1;
        const uint32_t src = ring_s + static_cast<uint32_t>(slot * 2 * THREADS * 16);
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x0.x), "=f"(x0.y), "=f"(x0.z), "=f"(x0.w) : "r"(src) : "memory");
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x1.x), "=f"(x1.y), "=f"(x1.z), "=f"(x1.w) : "r"(src + THREADS * 16) : "memory");
        issue(slot, it + DEPTH);
        const float d[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
        uint32_t e[8];
#pragma unroll
        for (int sidx = 0; sidx < 8; ++sidx) e[sidx] = entry_of(d[sidx]);
        // (with ~120 thresholds in 4096 bins some lane needs the exact path in nearly every group; a
        // warp-wide "while any lane has a flagged slot" loop that re-reads the element from the ring
        // was measured slower than these eight per-slot branches: 2.01 vs 1.88 ms at DeepChange)
        if ((e[0] | e[1] | e[2] | e[3] | e[4] | e[5] | e[6] | e[7]) & 0x8000u) {
#pragma unroll
          for (int sidx = 0; sidx < 8; ++sidx)
            if (e[sidx] & 0x8000u) e[sidx] = exact(e[sidx], d[sidx], ga + (sidx < 4 ? 0u : 4u * THREADS) + (sidx & 3));
        }
#pragma unroll
        for (int sidx = 0; sidx < 8; ++sidx) bump(e[sidx]);
        ga += 8u * THREADS;
        slot = slot + 1 == DEPTH ? 0 : slot + 1;
      }
      auto put = [&](float dv, uint32_t g) {
        uint32_t e = entry_of(dv);
        if (e & 0x8000u) e = exact(e, dv, g);
        bump(e);
      };
      if (tid < head) put(__ldg(row + c0 + tid), gbase + static_cast<uint32_t>(c0 + tid));
      const int64_t ct0 = cv0 + 4 * static_cast<int64_t>(nvec);
      if (tid < tail) put(__ldg(row + ct0 + tid), gbase + static_cast<uint32_t>(ct0 + tid));
      __syncthreads();
      // per bucket group: the four byte fields of all threads' words, summed as two pairs of 16-bit
      // fields (256 x 255 < 65536), one warp per group
      const int lane = tid & 31, warp = tid >> 5;
      for (int g4 = warp; g4 < groups; g4 += THREADS / 32) {
        uint32_t s02 = 0, s13 = 0;
        for (int e2 = lane; e2 < THREADS; e2 += 32) {
          const uint32_t w = cnt_dyn[g4 * THREADS + e2];
          s02 += w & 0x00FF00FFu;
          s13 += (w >> 8) & 0x00FF00FFu;
        }
        s02 = __reduce_add_sync(0xffffffffu, s02);
        s13 = __reduce_add_sync(0xffffffffu, s13);
        if (lane == 0) {
          const int b = 4 * g4;
          bh2[b] = s02 & 0xFFFFu;
          if (b + 1 <= kCap) bh2[b + 1] = s13 & 0xFFFFu;
          if (b + 2 <= kCap) bh2[b + 2] = s02 >> 16;
          if (b + 3 <= kCap) bh2[b + 3] = s13 >> 16;
        }
      }
      __syncthreads();
      if (tid < n) {  // count_below(T[i]) = elements in buckets 0 .. i
        uint32_t sum = 0;
        for (int b = 0; b <= tid; ++b) sum += bh2[b];
        hist[tid] = sum;
      }
    };
    if constexpr (BYTES) {
      run_bytes();
    } else {
      if (n <= 32) run(std::integral_constant<int, 1>{}); else run(std::integral_constant<int, 2>{});
    }
  } else {
    // generic: bucket histogram with exact compares, then a prefix sum
    __shared__ uint32_t bh[kCap + 1];
    for (int t = tid; t <= kCap; t += THREADS) bh[t] = 0u;
    __syncthreads();
    for (int64_t c = c0 + tid; c < c1; c += THREADS) {
      const uint64_t cc = composite(dist_key(__ldg(row + c)), gbase + static_cast<uint32_t>(c));
      int b = 0;
      while (T[b] <= cc) ++b;
      atomicAdd(&bh[b], 1u);
    }
    __syncthreads();
    if (tid < n) {
      uint32_t sum = 0;
      for (int b = 0; b <= tid; ++b) sum += bh[b];
      hist[tid] = sum;
    }
  }
  __syncthreads();

  // 5. hist[i] = count_below(T[i]); scatter back to plan order
  if (!FUSED) {
    if (tid < n) {
      const uint32_t below = hist[tid];
      int32_t *dst = counts + o + orig[tid];
      if (nsplit > 1) {
        if (below) atomicAdd(dst, static_cast<int32_t>(below));
      } else {
        *dst = static_cast<int32_t>(below);
      }
    }
    return;
  }
  // 6. FUSED: finish the query (as rank_count_v2_kernel / rank_finalize_kernel do).  T[] is sorted
  // by rank; the junk matches ahead of a positive are subtracted; the AP terms k / rank_k are formed
  // in double by the positives' threads, thread 0 then only runs torchreid's sequential float sum.
  __shared__ int32_t s_first;
  if (tid < n) {
    const uint64_t ck = T[tid];
    int below_junk = 0;
    for (int u = 0; u < nj; ++u) below_junk += Tu[n + u] < ck ? 1 : 0;
    for (int u = n + nj; u < m; ++u) {
      const uint32_t g = static_cast<uint32_t>(__ldg(gid + o + u));
      below_junk += composite(dist_key(__ldg(row + g)), g) < ck ? 1 : 0;
    }
    const int r = static_cast<int>(hist[tid]) - below_junk + 1;  // 1-based rank among kept items
    if (tid == 0) s_first = r;
    fo.ranks_sorted[o + tid] = r;
    s_term[tid] = static_cast<double>(tid + 1) / static_cast<double>(r);
  }
  __syncthreads();
  if (tid == 0) {
    float sum = 0.f;  // torchreid Cython accumulation: float running sum, each term formed in double
    for (int k = 0; k < n; ++k) sum = static_cast<float>(static_cast<double>(sum) + s_term[k]);
    fo.ap[q] = sum / static_cast<float>(n);
    const int fr = s_first;
    fo.first_rank[q] = fr;
    if (fr <= fo.max_rank) atomicAdd(fo.cmc_cnt + (fr - 1), 1);
    atomicAdd(fo.cmc_cnt + fo.max_rank, 1);  // num_valid_q
  }
}

// ------------------- one warp per query (single GPU, <= 62 matches) -------------------
// The CTA-per-query kernel above passes every query through nine block-wide barriers and repeats
// the per-query prologue / epilogue code in all of its eight warps.  Here ONE WARP owns a query:
// lane i holds thresholds i and i + 32 (sorted by shuffles), the table is the warp's own 4 KB of
// shared memory, the row is streamed through a per-lane cp.async ring DEPTH iterations deep (1 KB per
// warp and iteration, contiguous), the counters are the same bit-sliced registers (one word of
// thresholds for queries with <= 32 positives, two otherwise), the totals are PL warp transposes per
// word, and the epilogue (junk, kept ranks, torchreid's sequential float AP, first-match histogram)
// runs on shuffles.  No __syncthreads anywhere: the warps of a CTA only share its shared-memory
// allocation.  Same arithmetic and the same results as rank_count_v3_kernel<.., FUSED>.
#ifndef DALI_WPQ_DEPTH
#define DALI_WPQ_DEPTH 4
#endif
#ifndef DALI_WPQ_WARPS
#define DALI_WPQ_WARPS 4
#endif
#ifndef DALI_WPQ_MINB
#define DALI_WPQ_MINB 6  // resident CTAs per SM the kernel is compiled for (one wave at the Market shapes)
#endif
constexpr int kWpqWarps = DALI_WPQ_WARPS;
constexpr int kWpqMaxM = 62;  // matches (valid + junk bound the valid ones) of one query

template <int LOG2NB, int PL>
__global__ void __launch_bounds__(kWpqWarps * 32, DALI_WPQ_MINB)
rank_count_wpq_kernel(const float *__restrict__ dist, int64_t ld, int64_t G,
                      const int64_t *__restrict__ off, const int32_t *__restrict__ nvalid,
                      const int32_t *__restrict__ gid, FusedOut fo, int64_t nq) {
  constexpr int NB = 1 << LOG2NB;
  constexpr int NE = NB + 8;
  constexpr int DEPTH = DALI_WPQ_DEPTH;
  constexpr uint32_t FULL = 0xffffffffu;
  struct __align__(16) WarpShared {
    float4 ring[DEPTH][2][32];
    uint16_t lut[NE];  // before the table is built (and on the generic path): Tc[64] | bh[65]
    float Tf[65];      // sorted thresholds as distances, +inf from index n on
    uint32_t Tg[65];   // their gallery ids
    uint16_t tb[66];   // tb[1 + i] = bin of threshold i
  };
  static_assert(NE * 2 >= 64 * 8 + 65 * 4, "the sort's scatter buffer and the generic histogram alias the table");
  __shared__ WarpShared shared_[kWpqWarps];
  const int lane = threadIdx.x & 31;
  const int64_t q = static_cast<int64_t>(blockIdx.x) * kWpqWarps + (threadIdx.x >> 5);
  if (q >= nq) return;  // (whole warps; nothing below synchronises across warps)
  WarpShared &S = shared_[threadIdx.x >> 5];
  uint64_t *Tc = reinterpret_cast<uint64_t *>(S.lut);             // sorted composites
  uint32_t *bh = reinterpret_cast<uint32_t *>(S.lut) + 64 * 2;    // generic path: elements per bucket
  const float *row = dist + q * ld;

  // [cv0, cv0 + 4 nvec) is the 16-byte aligned part of the row, dealt out as float4 vectors:
  // iteration `it` gives a lane the vectors lane + 64 it and that + 32; vectors beyond the row read as
  // NaN (above every threshold)
  const int mis = static_cast<int>((reinterpret_cast<uintptr_t>(row) >> 2) & 3);
  int head = (4 - mis) & 3;
  if (head > G) head = static_cast<int>(G);
  const int64_t cv0 = head;
  const int nvec = static_cast<int>((G - cv0) >> 2);
  const int tail = static_cast<int>(G - cv0) & 3;
  const int niter = (nvec + 63) / 64;
  const int nfull = nvec / 64;
  const float qnan = __int_as_float(0x7FFFFFFF);
  const float *pv = row + cv0 + 4 * lane;
  const uint32_t ring_s = static_cast<uint32_t>(__cvta_generic_to_shared(&S.ring[0][0][lane]));
  auto issue = [&](int slot, int itx) {  // always one commit group
    if (itx < nfull) {
      const uint32_t dst = ring_s + static_cast<uint32_t>(slot * 64 * 16);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(pv) : "memory");
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 32 * 16), "l"(pv + 4 * 32) : "memory");
    } else if (itx < niter) {
      const int rem = nvec - lane - itx * 64;  // vectors left from pv on, in steps of 32
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const uint32_t dst = ring_s + static_cast<uint32_t>((slot * 2 + j) * 32 * 16);
        if (rem > j * 32)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(pv + 4 * j * 32) : "memory");
        else
          asm volatile("st.shared.v4.f32 [%0], {%1, %1, %1, %1};" ::"r"(dst), "f"(qnan) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    pv += 8 * 32;
  };
#pragma unroll
  for (int sl = 0; sl < DEPTH; ++sl) issue(sl, sl);

  const int n = __ldg(nvalid + q);  // <= kWpqMaxM: the launcher checks the bound of every query
  const int64_t o = __ldg(off + q);
  const int m = static_cast<int>(__ldg(off + q + 1) - o);
  if (n == 0) {
    if (lane == 0) {
      fo.ap[q] = 0.f;
      fo.first_rank[q] = -1;
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    return;
  }

  // 1. thresholds: lane i fetches matches i and i + 32; valid ones are the thresholds, the junk ones
  // (same identity, same camera, listed behind the valid ones; m <= 62) are used in the epilogue
  uint64_t c[2] = {~0ull, ~0ull}, jc[2] = {~0ull, ~0ull};
  {
    uint32_t g[2] = {0u, 0u};
#pragma unroll
    for (int w = 0; w < 2; ++w)
      if (lane + 32 * w < m) g[w] = static_cast<uint32_t>(__ldg(gid + o + lane + 32 * w));
#pragma unroll
    for (int w = 0; w < 2; ++w)
      if (lane + 32 * w < m) {
        const uint64_t x = composite(dist_key(__ldg(row + g[w])), g[w]);
        if (lane + 32 * w < n) c[w] = x; else jc[w] = x;
      }
  }
  int pos[2] = {0, 0};  // composites are distinct (gallery ids differ): a permutation of 0 .. n-1
  for (int u = 0; u < min(n, 32); ++u) {
    const uint64_t x = __shfl_sync(FULL, c[0], u);
    pos[0] += x < c[0] ? 1 : 0;
    pos[1] += x < c[1] ? 1 : 0;
  }
  for (int u = 32; u < n; ++u) {
    const uint64_t x = __shfl_sync(FULL, c[1], u - 32);
    pos[0] += x < c[0] ? 1 : 0;
    pos[1] += x < c[1] ? 1 : 0;
  }
#pragma unroll
  for (int w = 0; w < 2; ++w)
    if (lane + 32 * w < n) Tc[pos[w]] = c[w];
  __syncwarp();
  uint64_t cs[2];     // lane i: the i-th and (i + 32)-th smallest threshold
  uint32_t tkey[2];
  float tf[2];
#pragma unroll
  for (int w = 0; w < 2; ++w) {
    const int i = lane + 32 * w;
    cs[w] = i < n ? Tc[i] : ~0ull;
    tkey[w] = static_cast<uint32_t>(cs[w] >> 32);
    tf[w] = i < n ? key_to_dist(tkey[w]) : __int_as_float(0x7F800000);
    S.Tf[i] = tf[w];
    S.Tg[i] = i < n ? static_cast<uint32_t>(cs[w]) : 0xFFFFFFFFu;
  }
  if (lane == 0) {
    S.Tf[64] = __int_as_float(0x7F800000);
    S.Tg[64] = 0xFFFFFFFFu;
  }
  __syncwarp();
  Tc[lane] = cs[0];  // sorted, ~0 from index n on (the generic path walks it)
  Tc[lane + 32] = cs[1];

  // 2. the bin map of rank_count_v3_kernel
  const int il = (n - 1) & 31;
  const uint32_t klo = __shfl_sync(FULL, tkey[0], 0);
  const uint32_t khi = n > 32 ? __shfl_sync(FULL, tkey[1], il) : __shfl_sync(FULL, tkey[0], il);
  const float lo = __shfl_sync(FULL, tf[0], 0);
  const float hi = n > 32 ? __shfl_sync(FULL, tf[1], il) : __shfl_sync(FULL, tf[0], il);
  float sc = fminf(__fdividef(static_cast<float>(NB - 4) / static_cast<float>(NB), hi - lo),
                   __fdividef(524288.0f / static_cast<float>(NB), fabsf(hi)));
  sc = fminf(fmaxf(sc, 1.0e-30f), 1.0e30f);
  const float nsc = -sc;
  const float C = fmaf(hi, sc, 2.0f / static_cast<float>(NB));
  const bool fast = klo > 0x007FFFFFu && khi < 0xFF800000u;
  auto bin_bits = [&](float d) -> uint32_t {
    float u;
    asm("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(u) : "f"(d), "f"(nsc), "f"(C));
    return __float_as_uint(fmaf(u, static_cast<float>(NB), 8388608.0f));
  };
  uint32_t total[2] = {0u, 0u};  // lane i: count_below(threshold i), count_below(threshold i + 32)

  if (fast) {
    // 3. table: entry(b) = #{thresholds in higher bins}, | 0x8000 if bin b holds thresholds itself
    uint16_t *tbs = S.tb + 1;
    __syncwarp();  // (the generic-path arrays alias the table: every lane is past its reads of Tc)
#pragma unroll
    for (int w = 0; w < 2; ++w)
      if (lane + 32 * w < n) tbs[lane + 32 * w] = static_cast<uint16_t>(bin_bits(tf[w]) - kV3Magic);
    __syncwarp();
    for (int blk = lane; blk < NE / 8; blk += 32) {
      const int b0 = blk * 8;
      int l = 0, h = n;  // first index with tb < b0
      while (l < h) {
        const int mid = (l + h) >> 1;
        if (static_cast<int>(tbs[mid]) >= b0) l = mid + 1; else h = mid;
      }
      const uint32_t wv = static_cast<uint32_t>(l) * 0x00010001u;
      *reinterpret_cast<uint4 *>(S.lut + b0) = make_uint4(wv, wv, wv, wv);
    }
    __syncwarp();
    for (int i = lane; i < n; i += 32) {
      const int b = tbs[i];
      const int up = i > 0 ? static_cast<int>(tbs[i - 1]) : NE;
      const int dn = i + 1 < n ? static_cast<int>(tbs[i + 1]) : -1;
      const int bs = b & ~7, be = bs + 8;
      if (up != b) {
        S.lut[b] = static_cast<uint16_t>(i | 0x8000);
        for (int x = b + 1; x < min(up, be); ++x) S.lut[x] = static_cast<uint16_t>(i);
      }
      if (dn < bs)
        for (int x = bs; x < b; ++x) S.lut[x] = static_cast<uint16_t>(i + 1);
    }
    __syncwarp();

    uint32_t lut_s = static_cast<uint32_t>(__cvta_generic_to_shared(S.lut)) - (kV3Magic << 1);
    asm volatile("mov.u32 %0, %0;" : "+r"(lut_s));
    auto entry_of = [&](float d) -> uint32_t {
      uint32_t e;
      asm volatile("ld.shared.u16 %0, [%1];" : "=r"(e) : "r"(lut_s + (bin_bits(d) << 1)));
      return e;
    };
    auto exact = [&](uint32_t e, float d, uint32_t g) -> uint32_t {
      uint32_t b = e & 0x7FFFu;
      while (true) {
        const float t = S.Tf[b];
        if (t < d || (t == d && S.Tg[b] <= g)) ++b; else break;
      }
      return b;
    };
    auto run = [&](auto nw_tag) {
      constexpr int NW = decltype(nw_tag)::value;
      SlicedCounters<NW, PL> scnt;
      scnt.clear();
      uint32_t mk[8][NW];
      auto to_mask = [&](int s, uint32_t b) {
        mk[s][0] = shl_clamp(0xFFFFFFFFu, b);
        if (NW == 2) mk[s][NW - 1] = shl_clamp(0xFFFFFFFFu, max(b, 32u) - 32u);
      };
      auto group = [&](const float4 &x0, const float4 &x1, uint32_t ga) {
        const float d[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
        uint32_t e[8];
#pragma unroll
        for (int s = 0; s < 8; ++s) e[s] = entry_of(d[s]);
        if ((e[0] | e[1] | e[2] | e[3] | e[4] | e[5] | e[6] | e[7]) & 0x8000u) {
#pragma unroll
          for (int s = 0; s < 8; ++s)
            if (e[s] & 0x8000u) e[s] = exact(e[s], d[s], ga + (s < 4 ? 0u : 4u * 32u) + (s & 3));
        }
#pragma unroll
        for (int s = 0; s < 8; ++s) to_mask(s, e[s]);
        scnt.add8(mk);
      };
      uint32_t ga = static_cast<uint32_t>(cv0) + 4u * lane;
      int slot = 0;
#pragma unroll 1
      for (int it = 0; it < niter; ++it) {
        asm volatile("cp.async.wait_group %0;" ::"n"(DEPTH - 1) : "memory");
        float4 x0, x1;
        const uint32_t src = ring_s + static_cast<uint32_t>(slot * 64 * 16);
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x0.x), "=f"(x0.y), "=f"(x0.z), "=f"(x0.w) : "r"(src) : "memory");
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x1.x), "=f"(x1.y), "=f"(x1.z), "=f"(x1.w) : "r"(src + 32 * 16) : "memory");
        issue(slot, it + DEPTH);  // the slot is this lane's own
        group(x0, x1, ga);
        ga += 8u * 32u;
        slot = slot + 1 == DEPTH ? 0 : slot + 1;
      }
      if (head | tail) {  // the unaligned head and the tail of the row (<= 3 columns each)
#pragma unroll
        for (int s = 0; s < 8; ++s)
#pragma unroll
          for (int w = 0; w < NW; ++w) mk[s][w] = 0u;
        auto put = [&](int s, float d, uint32_t g) {
          uint32_t e = entry_of(d);
          if (e & 0x8000u) e = exact(e, d, g);
          to_mask(s, e);
        };
        if (lane < head) put(0, __ldg(row + lane), static_cast<uint32_t>(lane));
        const int64_t ct0 = cv0 + 4 * static_cast<int64_t>(nvec);
        if (lane < tail) put(1, __ldg(row + ct0 + lane), static_cast<uint32_t>(ct0 + lane));
        scnt.add8(mk);
      }
      // 4. totals: plane k of word w transposed across the warp gives lane i bit k of every lane's
      // count for threshold i + 32 w
#pragma unroll
      for (int w = 0; w < NW; ++w)
#pragma unroll
        for (int k = 0; k < PL; ++k)
          total[w] += static_cast<uint32_t>(__popc(warp_bit_transpose(scnt.p[k][w], lane))) << k;
    };
    if (n <= 32) run(std::integral_constant<int, 1>{}); else run(std::integral_constant<int, 2>{});
  } else {
    // generic (thresholds not all finite): bucket histogram with exact compares, then a prefix sum
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncwarp();
    bh[lane] = 0u;
    bh[lane + 32] = 0u;
    if (lane == 0) bh[64] = 0u;
    __syncwarp();
    for (int64_t col = lane; col < G; col += 32) {
      const uint64_t cc = composite(dist_key(__ldg(row + col)), static_cast<uint32_t>(col));
      int b = 0;
      while (b < n && Tc[b] <= cc) ++b;
      atomicAdd(&bh[b], 1u);
    }
    __syncwarp();
#pragma unroll
    for (int w = 0; w < 2; ++w)
      if (lane + 32 * w < n)
        for (int b = 0; b <= lane + 32 * w; ++b) total[w] += bh[b];
  }

  // 5. the query's results.  Junk matches ahead of a positive do not count; the AP terms k / rank_k
  // are formed in double, torchreid's float running sum adds them in rank order.
  int below_junk[2] = {0, 0};
  for (int u = n; u < m; ++u) {  // match u sits in lane u & 31, word u >> 5
    const uint64_t x = u < 32 ? __shfl_sync(FULL, jc[0], u) : __shfl_sync(FULL, jc[1], u - 32);
    below_junk[0] += x < cs[0] ? 1 : 0;
    below_junk[1] += x < cs[1] ? 1 : 0;
  }
  int r[2];
  double term[2];
#pragma unroll
  for (int w = 0; w < 2; ++w) {
    const int i = lane + 32 * w;
    r[w] = i < n ? static_cast<int>(total[w]) - below_junk[w] + 1 : 1;  // 1-based rank among kept items
    if (i < n) fo.ranks_sorted[o + i] = r[w];
    term[w] = static_cast<double>(i + 1) / static_cast<double>(r[w]);
  }
  float sum = 0.f;
  for (int k = 0; k < min(n, 32); ++k) sum = static_cast<float>(static_cast<double>(sum) + __shfl_sync(FULL, term[0], k));
  for (int k = 32; k < n; ++k) sum = static_cast<float>(static_cast<double>(sum) + __shfl_sync(FULL, term[1], k - 32));
  if (lane == 0) {
    fo.ap[q] = sum / static_cast<float>(n);
    fo.first_rank[q] = r[0];
    if (r[0] <= fo.max_rank) atomicAdd(fo.cmc_cnt + (r[0] - 1), 1);
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
  // (for queries with many matches -- DeepChange: ~120 valid + junk -- a bitonic sort of the m
  //  composites + a ballot scan over the junk flags instead of this all-pairs loop was measured
  //  slower: 0.218 vs 0.178 ms, 36 block barriers per query)
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
  // the AP terms k / rank_k, formed in double by all threads (the composites are dead now: their
  // shared memory holds the terms); thread 0 then only runs the sequential sum -- the ~120 dependent
  // double divisions of a DeepChange query used to sit in its loop
  double *s_term = reinterpret_cast<double *>(s_c);
  if (staged) {
    for (int k = tid; k < nv; k += kFinThreads)
      s_term[k] = static_cast<double>(k + 1) / static_cast<double>(s_rank[k]);
    __syncthreads();
  }
  if (tid == 0) {
    // torchreid Cython accumulation: float running sum, each term formed in double
    float s = 0.f;
    for (int k = 1; k <= nv; ++k) {
      const double term = staged ? s_term[k - 1]
                                 : static_cast<double>(k) / static_cast<double>(ranks_sorted[o + k - 1]);
      s = static_cast<float>(static_cast<double>(s) + term);
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

int launch_plan_expand(dali_ctx *ctx, const dali_rank_plan *plan, cudaStream_t stream) {
  if (plan->Q == 0) return DALI_OK;
  plan_expand_kernel<<<static_cast<unsigned>((plan->Q + 3) / 4), 128, 0, stream ? stream : ctx->stream>>>(
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

// v3 (register-resident bit-sliced counters): queries with <= 64 valid positives.  PL = number of
// counter planes: a thread must see fewer than 2^PL elements.
static bool v3_enabled() {
  static const char *env = getenv("DALI_RANK_V3");
  return !(env && atoi(env) == 0);
}
static int v3_threads() {
  static const char *env = getenv("DALI_RANK_V3_THREADS");
  return env && atoi(env) == 128 ? 128 : 256;
}
template <int THREADS, bool FUSED>
static int launch_v3(dali_ctx *ctx, dim3 grid, const dali_rank_plan *plan, const float *dist, int64_t ld,
                     int64_t g0, int64_t Gs, const uint32_t *keys, int32_t *counts, int nsplit, FusedOut fo) {
  const int64_t per = (((Gs + nsplit - 1) / nsplit) + 3) & ~int64_t(3);
  const int64_t per_thread = ((per / 4 + THREADS - 1) / THREADS) * 4 + 2;
  static const char *env_nb = getenv("DALI_RANK_V3_LOG2NB");
  const int lnb = env_nb ? atoi(env_nb) : 11;
#define DALI_V3_LAUNCH(LNB, PL)                                                                      \
  rank_count_v3_kernel<LNB, THREADS, PL, FUSED><<<grid, THREADS, 0, ctx->stream>>>(                  \
      dist, ld, g0, Gs, plan->d_off, plan->d_nv, plan->d_gid, keys, counts, nsplit, fo)
  // rows of the Market shapes (a thread sees < 128 elements): the variant compiled for five resident
  // 256-thread CTAs per SM (48 registers) -- occupancy is worth 25 % here (0.094 -> 0.072 ms at C2)
  static const char *env_tight = getenv("DALI_RANK_V3_TIGHT");
  if (per_thread < (1 << 7)) {
    if (!(env_tight && atoi(env_tight) == 0) && lnb == 11)
      rank_count_v3_kernel<11, THREADS, 7, FUSED, true><<<grid, THREADS, 0, ctx->stream>>>(
          dist, ld, g0, Gs, plan->d_off, plan->d_nv, plan->d_gid, keys, counts, nsplit, fo);
    else if (!(env_tight && atoi(env_tight) == 0) && lnb == 12)
      rank_count_v3_kernel<12, THREADS, 7, FUSED, true><<<grid, THREADS, 0, ctx->stream>>>(
          dist, ld, g0, Gs, plan->d_off, plan->d_nv, plan->d_gid, keys, counts, nsplit, fo);
    else if (lnb == 12) DALI_V3_LAUNCH(12, 7); else DALI_V3_LAUNCH(11, 7);
  } else if (per_thread < (1 << 10)) {
    if (lnb == 12) DALI_V3_LAUNCH(12, 10); else DALI_V3_LAUNCH(11, 10);
  } else if (per_thread < (1 << 14)) {
    if (lnb == 12) DALI_V3_LAUNCH(12, 14); else DALI_V3_LAUNCH(11, 14);
  } else {
    return set_err(ctx, DALI_ERR_UNSUPPORTED, "row segment too long for the counting kernel");
  }
#undef DALI_V3_LAUNCH
  return DALI_OK;
}

// v3 front end + byte counters: queries with 65 .. 254 valid positives (one chunk), rows of which a
// thread sees at most 255 elements
static bool v3b_enabled() {
  static const char *env = getenv("DALI_RANK_V3B");
  return v3_enabled() && !(env && atoi(env) == 0);
}
static bool v3b_fits(int64_t Gs, int nsplit) {
  const int64_t per = (((Gs + nsplit - 1) / nsplit) + 3) & ~int64_t(3);
  return ((per / 4 + 2 * 256 - 1) / (2 * 256)) * 8 + 2 <= 255;
}
template <bool FUSED>
static int launch_v3b(dali_ctx *ctx, dim3 grid, const dali_rank_plan *plan, const float *dist, int64_t ld,
                      int64_t g0, int64_t Gs, const uint32_t *keys, int32_t *counts, int nsplit, FusedOut fo) {
  const size_t smem = static_cast<size_t>((std::min(plan->max_nv, 254) + 4) >> 2) * 256 * sizeof(uint32_t);
  // static (38 KB) + dynamic shared memory exceed 48 KB even for few buckets: always opt in, for the most
  if (int rc = ensure_dyn_smem(ctx, reinterpret_cast<const void *>(&rank_count_v3_kernel<DALI_V3B_LOG2NB, 256, 7, FUSED, false, 256>),
                               size_t(64) * 256 * sizeof(uint32_t)))
    return rc;
  rank_count_v3_kernel<DALI_V3B_LOG2NB, 256, 7, FUSED, false, 256><<<grid, 256, smem, ctx->stream>>>(
      dist, ld, g0, Gs, plan->d_off, plan->d_nv, plan->d_gid, keys, counts, nsplit, fo);
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
  // v3 (register-resident counters, 64 thresholds per CTA): queries with more positives are counted
  // by several CTAs that each stream the row (DALI_RANK_V3_CHUNKED, default on up to kV3MaxChunks)
  static const char *env_ck = getenv("DALI_RANK_V3_CHUNKED");
  const int v3_max_chunks = env_ck ? atoi(env_ck) : kV3MaxChunks;
  const bool use_v3 = !use_v1 && !env_c && !env_t && v3_enabled() &&
                      plan->max_nv <= static_cast<int64_t>(kV3Chunk) * std::max(1, v3_max_chunks);
  const int per_cta = use_v1 ? kChunk : use_v3 ? kV3Chunk : tchunk;
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
  const bool use_v3b = !use_v1 && !use_v3 && !env_c && !env_t && v3b_enabled() && plan->max_nv > kV3Chunk && plan->max_nv <= 254;
  if (use_v3b) ns = std::max<int64_t>(ns, (Gs + 250ll * 256 - 1) / (250ll * 256));
  const bool bytec = !use_v1 && !use_v3 && !use_v3b && std::min(plan->max_nv, tchunk) > 64 && !(env_b && atoi(env_b) == 0);
  if (bytec) ns = std::max<int64_t>(ns, (Gs + 250ll * 128 - 1) / (250ll * 128));
  if (ns > 65535) return set_err(ctx, DALI_ERR_UNSUPPORTED, "slab too wide for one launch");
  const int nsplit = static_cast<int>(ns);
  dim3 grid(static_cast<unsigned>(plan->Q), nchunk, nsplit);
  KTimer t(ctx, DALI_K_RANK_COUNT);
  if (use_v3b && v3b_fits(Gs, nsplit)) {
    if (int rc = launch_v3b<false>(ctx, grid, plan, dist, ld, g0, Gs, keys, counts, nsplit, FusedOut{})) return rc;
  } else if (use_v3) {
    const int rc = v3_threads() == 128
        ? launch_v3<128, false>(ctx, grid, plan, dist, ld, g0, Gs, keys, counts, nsplit, FusedOut{})
        : launch_v3<256, false>(ctx, grid, plan, dist, ld, g0, Gs, keys, counts, nsplit, FusedOut{});
    if (rc) return rc;
  } else if (use_v1) {
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
  static const char *env_ck = getenv("DALI_RANK_V3_CHUNKED");
  if (bytec && v3_enabled() && plan->max_nv <= static_cast<int64_t>(kV3Chunk) * (env_ck ? atoi(env_ck) : kV3MaxChunks))
    return DALI_OK;  // gather + chunked v3 count + finalize
  static const char *env_b = getenv("DALI_RANK_FUSED_BYTE");
  if (bytec && (G > 250ll * 256 || (env_b && atoi(env_b) == 0))) return DALI_OK;
  DALI_CUDA_OK(ctx, cudaMemsetAsync(cmc_cnt, 0, sizeof(int32_t) * (max_rank + 1), ctx->stream));
  FusedOut fo{ranks_sorted, ap, first_rank, cmc_cnt, max_rank};
  KTimer t(ctx, DALI_K_RANK_COUNT);
  const dim3 grid(static_cast<unsigned>(plan->Q), 1, 1);
  // one warp per query (opt-in, DALI_RANK_WPQ=1; =2 also for few queries): measured 4 % faster than
  // the CTA-per-query kernel at C1 and equal at C2 -- one wave of warps ends with its
  // slowest query
  static const char *env_w = getenv("DALI_RANK_WPQ");
  const int wpq = env_w ? atoi(env_w) : 0;
  const int64_t per_lane = ((G / 4 + 63) / 64) * 8 + 2;  // elements a lane may see
  if (wpq > 0 && plan->max_m <= kWpqMaxM && v3_enabled() &&
      (plan->Q >= 8ll * ctx->num_sms || wpq == 2) && per_lane < (1 << 12)) {
    const unsigned ctas = static_cast<unsigned>((plan->Q + kWpqWarps - 1) / kWpqWarps);
    if (per_lane < (1 << 9))
      rank_count_wpq_kernel<11, 9><<<ctas, kWpqWarps * 32, 0, ctx->stream>>>(
          dist, ld, G, plan->d_off, plan->d_nv, plan->d_gid, fo, plan->Q);
    else
      rank_count_wpq_kernel<11, 12><<<ctas, kWpqWarps * 32, 0, ctx->stream>>>(
          dist, ld, G, plan->d_off, plan->d_nv, plan->d_gid, fo, plan->Q);
  } else if (plan->max_nv <= kV3Chunk && v3_enabled()) {
    const int rc = v3_threads() == 128
        ? launch_v3<128, true>(ctx, grid, plan, dist, ld, 0, G, nullptr, nullptr, 1, fo)
        : launch_v3<256, true>(ctx, grid, plan, dist, ld, 0, G, nullptr, nullptr, 1, fo);
    if (rc) return rc;
  } else if (bytec && v3b_enabled() && v3b_fits(G, 1)) {
    if (int rc = launch_v3b<true>(ctx, grid, plan, dist, ld, 0, G, nullptr, nullptr, 1, fo)) return rc;
  } else if (bytec) {
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
