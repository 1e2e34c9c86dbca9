// k-reciprocal re-ranking: the hook the reference keeps commented out at every
// distance-matrix site,
//     distmat = torchreid.utils.re_ranking(distmat, distmat_qq, distmat_gg)
// (validateModels.py:49-53, evaluate.py:294-298, evaluate_ensembled_models.py:284-288,303-307;
// `rerank` flag: validateModels.py:28-31).  Algorithm: Zhong et al., CVPR 2017, in the numpy form
// torchreid ships (restated in oracle/rerank_oracle.py), N = Q + G samples:
//   1. od[i,j] = C[j,i]^2 / max_r C[r,i]^2 for the concatenated matrix C = [[qq, qg], [qg^T, gg]]
//   2. initial_rank = the k1+1 nearest of every row (stable order) -> topk_kernel
//   3. per sample: k-reciprocal set, its 2/3-overlap expansion with the k1/2 sets of its members,
//      V0[i, e] = exp(-od[i,e]) / sum over the expansion set
//   4. query expansion: V[i] = mean of V0 over the k2 nearest (sequential fp32 sum in rank order)
//   5. Jaccard distance of every query to every sample through the inverted index of V, in
//      ascending column order like the reference's loop; final = (1-lambda) jaccard + lambda od
// V is kept sparse (<= 512 entries per row before, k2 * 512 after the expansion).  Index sets are
// exact; values agree with the numpy form to fp32 rounding (exp differs in the last ulp).
#include <algorithm>
#include <cmath>

#include "common.cuh"

namespace dali {

namespace {

constexpr int kCap1 = 512;  // entries of one V0 row: (k1+1) * (round(k1/2)+2) <= 512  <=>  k1 <= 28

// ---- 1. column maxima of the squared concatenated matrix, then od ---------------------------
// values are squares (>= 0), so their bit patterns order like unsigned integers
__global__ void colmax_sq_kernel(const float *__restrict__ m, int64_t rows, int64_t cols, int64_t ld,
                                 uint32_t *__restrict__ out) {
  const int64_t c = static_cast<int64_t>(blockIdx.x) * 32 + threadIdx.x;
  if (c >= cols) return;
  float best = 0.f;
  for (int64_t r = blockIdx.y * blockDim.y + threadIdx.y; r < rows; r += static_cast<int64_t>(gridDim.y) * blockDim.y) {
    const float x = __ldg(m + r * ld + c);
    best = fmaxf(best, __fmul_rn(x, x));
  }
  atomicMax(out + c, __float_as_uint(best));
}

__global__ void rowmax_sq_kernel(const float *__restrict__ m, int64_t rows, int64_t cols, int64_t ld,
                                 uint32_t *__restrict__ out) {
  const int64_t r = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  float best = 0.f;
  for (int64_t c = threadIdx.x & 31; c < cols; c += 32) {
    const float x = __ldg(m + r * ld + c);
    best = fmaxf(best, __fmul_rn(x, x));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out + r, __float_as_uint(best));
}

// dst[r, c] = (s * s) / cmax[r],  s = transposed ? src[c, r] : src[r, c].  One CTA (32 x 8 threads)
// produces 32 rows x 128 columns of dst: sixteen elements per thread, all sixteen loads issued before
// the first division (32 x 32 tiles with four elements per thread ran at 2.3 TB/s; od's rows are N
// floats apart with N odd, so its stores stay 4 bytes wide).
constexpr int kOdCols = 128;
__global__ void __launch_bounds__(256)
od_block_kernel(const float *__restrict__ src, int64_t src_ld, int transposed,
                int64_t rows, int64_t cols, float *__restrict__ dst, int64_t dst_ld,
                const float *__restrict__ cmax) {
  __shared__ float tile[kOdCols][33];
  const int64_t r0 = static_cast<int64_t>(blockIdx.y) * 32, c0 = static_cast<int64_t>(blockIdx.x) * kOdCols;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  if (transposed) {
    // src rows c0 .. c0+127, columns r0 .. r0+31 (coalesced along src's columns)
    float v[kOdCols / 8];
#pragma unroll
    for (int i = 0; i < kOdCols / 8; ++i) {
      const int64_t sr = c0 + ty + 8 * i, sc = r0 + tx;
      v[i] = (sr < cols && sc < rows) ? __ldg(src + sr * src_ld + sc) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < kOdCols / 8; ++i) tile[ty + 8 * i][tx] = v[i];
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int k = ty + 8 * i;
      const int64_t r = r0 + k;
      if (r < rows) {
        const float cm = cmax[r];
#pragma unroll
        for (int j = 0; j < kOdCols / 32; ++j) {
          const int64_t c = c0 + tx + 32 * j;
          if (c < cols) {
            const float s = tile[tx + 32 * j][k];
            dst[r * dst_ld + c] = __fdiv_rn(__fmul_rn(s, s), cm);
          }
        }
      }
    }
  } else {
    float v[4][kOdCols / 32];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < kOdCols / 32; ++j) {
        const int64_t r = r0 + ty + 8 * i, c = c0 + tx + 32 * j;
        v[i][j] = (r < rows && c < cols) ? __ldg(src + r * src_ld + c) : 0.f;
      }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t r = r0 + ty + 8 * i;
      if (r < rows) {
        const float cm = cmax[r];
#pragma unroll
        for (int j = 0; j < kOdCols / 32; ++j) {
          const int64_t c = c0 + tx + 32 * j;
          if (c < cols) dst[r * dst_ld + c] = __fdiv_rn(__fmul_rn(v[i][j], v[i][j]), cm);
        }
      }
    }
  }
}

// ---- 3. k-reciprocal sets, expansion, V0 ----------------------------------------------------
// one warp per sample
__global__ void __launch_bounds__(128)
krecip_kernel(const float *__restrict__ od, int64_t N, const int32_t *__restrict__ rank, int K, int H,
              int32_t *__restrict__ v_idx, float *__restrict__ v_val, int32_t *__restrict__ v_cnt) {
  __shared__ int32_t s_exp[4][kCap1];
  __shared__ int32_t s_rec[4][32];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t i = static_cast<int64_t>(blockIdx.x) * 4 + w;
  if (i >= N) return;
  int32_t *ex = s_exp[w];
  int32_t *rec = s_rec[w];
  const int32_t *ri = rank + i * K;
  // forward neighbours whose own K nearest contain i
  int32_t f = -1;
  bool is_rec = false;
  if (lane < K) {
    f = ri[lane];
    if (f >= 0) {
      const int32_t *rf = rank + static_cast<int64_t>(f) * K;
      for (int t = 0; t < K; ++t) is_rec |= (rf[t] == static_cast<int32_t>(i));
    }
  }
  const unsigned m_rec = __ballot_sync(0xffffffffu, is_rec);
  const int n_rec = __popc(m_rec);
  if (is_rec) {
    const int p = __popc(m_rec & ((1u << lane) - 1u));
    rec[p] = f;
    ex[p] = f;
  }
  __syncwarp();
  int n_ex = n_rec;
  for (int p = 0; p < n_rec; ++p) {
    const int32_t c = rec[p];
    const int32_t *rc = rank + static_cast<int64_t>(c) * K;  // first H entries = its k1/2 neighbours
    int32_t cf = -1;
    bool is_c = false;
    if (lane < H) {
      cf = rc[lane];
      if (cf >= 0) {
        const int32_t *rcf = rank + static_cast<int64_t>(cf) * K;
        for (int t = 0; t < H; ++t) is_c |= (rcf[t] == c);
      }
    }
    const unsigned m_c = __ballot_sync(0xffffffffu, is_c);
    const int n_c = __popc(m_c);
    bool in_rec = false;
    if (is_c)
      for (int t = 0; t < n_rec; ++t) in_rec |= (rec[t] == cf);
    const int n_int = __popc(__ballot_sync(0xffffffffu, in_rec));
    if (static_cast<double>(n_int) > 2.0 / 3.0 * static_cast<double>(n_c)) {
      if (is_c) ex[n_ex + __popc(m_c & ((1u << lane) - 1u))] = cf;
      n_ex += n_c;
    }
    __syncwarp();
  }
  // sort + unique (np.unique): bitonic over the next power of two, padding with INT32_MAX
  int np2 = 32;
  while (np2 < n_ex) np2 <<= 1;
  for (int t = n_ex + lane; t < np2; t += 32) ex[t] = INT32_MAX;
  __syncwarp();
  for (int size = 2; size <= np2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = lane; t < (np2 >> 1); t += 32) {
        const int pos = 2 * t - (t & (stride - 1));
        const bool asc = (pos & size) == 0;
        const int32_t a = ex[pos], b = ex[pos + stride];
        if ((a > b) == asc) { ex[pos] = b; ex[pos + stride] = a; }
      }
      __syncwarp();
    }
  }
  // weights of the unique members; the sum runs over them in ascending index order per lane
  const float *odi = od + i * N;
  float part = 0.f;
  for (int t = lane; t < n_ex; t += 32) {
    const int32_t e = ex[t];
    if (t == 0 || ex[t - 1] != e) part += expf(-__ldg(odi + e));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  int base = 0;
  int32_t *oi = v_idx + i * kCap1;
  float *ov = v_val + i * kCap1;
  for (int t0 = 0; t0 < n_ex; t0 += 32) {
    const int t = t0 + lane;
    bool uniq = false;
    int32_t e = 0;
    if (t < n_ex) {
      e = ex[t];
      uniq = (t == 0) || (ex[t - 1] != e);
    }
    const unsigned mu = __ballot_sync(0xffffffffu, uniq);
    if (uniq) {
      const int p = base + __popc(mu & ((1u << lane) - 1u));
      oi[p] = e;
      ov[p] = __fdiv_rn(expf(-__ldg(odi + e)), part);
    }
    base += __popc(mu);
  }
  if (lane == 0) v_cnt[i] = base;
}

// ---- 4. query expansion: V[i] = mean_{t<k2} V0[rank[i,t]] -----------------------------------
// one CTA per sample; entries of the k2 rows are sorted by (column, t) and summed in t order
__global__ void __launch_bounds__(128)
qexpand_kernel(int64_t N, const int32_t *__restrict__ rank, int K, int k2,
               const int32_t *__restrict__ v0_idx, const float *__restrict__ v0_val,
               const int32_t *__restrict__ v0_cnt, int cap2, int32_t *__restrict__ v_idx,
               float *__restrict__ v_val, int32_t *__restrict__ v_cnt) {
  extern __shared__ uint8_t sm_qe[];
  const int np2max = 1 << (32 - __clz(cap2 - 1));
  uint64_t *key = reinterpret_cast<uint64_t *>(sm_qe);           // [np2max]  (column << 32 | t << 16 | slot)
  float *val = reinterpret_cast<float *>(key + np2max);           // [cap2] values by slot
  __shared__ int s_off[9];
  __shared__ int s_total;
  const int64_t i = blockIdx.x;
  const int tid = threadIdx.x;
  if (tid == 0) {
    int acc = 0;
    for (int t = 0; t < k2; ++t) {
      s_off[t] = acc;
      const int32_t r = rank[i * K + t];
      acc += r >= 0 ? v0_cnt[r] : 0;
    }
    s_off[k2] = acc;
    s_total = acc;
  }
  __syncthreads();
  const int total = s_total;
  for (int t = 0; t < k2; ++t) {
    const int32_t r = rank[i * K + t];
    if (r < 0) continue;
    const int n = s_off[t + 1] - s_off[t];
    for (int e = tid; e < n; e += 128) {
      const int slot = s_off[t] + e;
      key[slot] = (static_cast<uint64_t>(static_cast<uint32_t>(v0_idx[static_cast<int64_t>(r) * kCap1 + e])) << 32) |
                  (static_cast<uint64_t>(t) << 16) | static_cast<uint64_t>(slot);
      val[slot] = v0_val[static_cast<int64_t>(r) * kCap1 + e];
    }
  }
  int np2 = 2;
  while (np2 < total) np2 <<= 1;
  for (int e = total + tid; e < np2; e += 128) key[e] = ~0ull;
  for (int size = 2; size <= np2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int t = tid; t < (np2 >> 1); t += 128) {
        const int pos = 2 * t - (t & (stride - 1));
        const bool asc = (pos & size) == 0;
        const uint64_t a = key[pos], b = key[pos + stride];
        if ((a > b) == asc) { key[pos] = b; key[pos + stride] = a; }
      }
    }
  }
  __syncthreads();
  // segment heads: one thread per distinct column sums its (<= k2) values in t order
  __shared__ int s_cnt;
  if (tid == 0) s_cnt = 0;
  __syncthreads();
  // rank of each head among heads = number of heads before it: block-wide prefix in chunks
  int32_t *oi = v_idx + i * cap2;
  float *ov = v_val + i * cap2;
  for (int e0 = 0; e0 < total; e0 += 128) {
    const int e = e0 + tid;
    bool head = false;
    if (e < total) head = (e == 0) || ((key[e - 1] >> 32) != (key[e] >> 32));
    // block prefix of `head`
    const unsigned mw = __ballot_sync(0xffffffffu, head);
    __shared__ int s_w[4];
    if ((tid & 31) == 0) s_w[tid >> 5] = __popc(mw);
    __syncthreads();
    int before = s_cnt;
    for (int w = 0; w < (tid >> 5); ++w) before += s_w[w];
    before += __popc(mw & ((1u << (tid & 31)) - 1u));
    if (head) {
      const uint32_t col = static_cast<uint32_t>(key[e] >> 32);
      float s = 0.f;
      bool first = true;
      for (int u = e; u < total && static_cast<uint32_t>(key[u] >> 32) == col; ++u) {
        const float x = val[key[u] & 0xFFFFu];
        s = first ? x : __fadd_rn(s, x);  // rows without this column contribute exact zeros
        first = false;
      }
      oi[before] = static_cast<int32_t>(col);
      ov[before] = __fdiv_rn(s, static_cast<float>(k2));
    }
    __syncthreads();
    if (tid == 0) s_cnt += s_w[0] + s_w[1] + s_w[2] + s_w[3];
    __syncthreads();
  }
  if (tid == 0) v_cnt[i] = s_cnt;
}

// ---- 5. inverted index of V, Jaccard distance, final mix -------------------------------------
__global__ void csc_count_kernel(int64_t N, const int32_t *__restrict__ v_idx, const int32_t *__restrict__ v_cnt,
                                 int cap, int32_t *__restrict__ col_cnt) {
  const int64_t i = blockIdx.x;
  const int n = v_cnt[i];
  for (int e = threadIdx.x; e < n; e += blockDim.x) atomicAdd(col_cnt + v_idx[i * cap + e], 1);
}

// exclusive scan of col_cnt[N] into col_off[N+1]: one CTA, every thread sums a contiguous run of
// ceil(n / 1024) counts, the 1024 run sums are scanned with shuffles (two barriers in all; the
// 1024-wide Hillis-Steele scan per chunk of the first version took 38 us at N = 19281)
__global__ void __launch_bounds__(1024) scan_kernel(const int32_t *__restrict__ in, int64_t n, int64_t *__restrict__ out) {
  __shared__ int64_t s_warp[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t per = (n + 1023) / 1024;
  const int64_t b = per * tid < n ? per * tid : n, e = b + per < n ? b + per : n;
  int64_t mine = 0;
  for (int64_t i = b; i < e; ++i) mine += in[i];
  int64_t incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int64_t y = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += y;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    const int64_t w = s_warp[lane];
    int64_t wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int64_t y = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += y;
    }
    s_warp[lane] = wi - w;  // exclusive prefix of the warp totals
    if (lane == 31) out[n] = wi;
  }
  __syncthreads();
  int64_t run = s_warp[warp] + incl - mine;
  for (int64_t i = b; i < e; ++i) {
    out[i] = run;
    run += in[i];
  }
}

__global__ void csc_fill_kernel(int64_t N, const int32_t *__restrict__ v_idx, const float *__restrict__ v_val,
                                const int32_t *__restrict__ v_cnt, int cap, const int64_t *__restrict__ col_off,
                                int32_t *__restrict__ col_fill, int32_t *__restrict__ csc_row,
                                float *__restrict__ csc_val) {
  const int64_t i = blockIdx.x;
  const int n = v_cnt[i];
  for (int e = threadIdx.x; e < n; e += blockDim.x) {
    const int32_t c = v_idx[i * cap + e];
    const int64_t p = col_off[c] + atomicAdd(col_fill + c, 1);
    csc_row[p] = static_cast<int32_t>(i);
    csc_val[p] = v_val[i * cap + e];
  }
}

// persistent CTAs, each with a temp[N] accumulator; columns of V[i] are visited in ascending
// order with a barrier between them, so every temp[r] receives its terms in the reference's order
template <int THREADS>
__global__ void __launch_bounds__(THREADS)
jaccard_kernel(int64_t Q, int64_t G, const float *__restrict__ od, const int32_t *__restrict__ v_idx,
               const float *__restrict__ v_val, const int32_t *__restrict__ v_cnt, int cap,
               const int64_t *__restrict__ col_off, const int32_t *__restrict__ csc_row,
               const float *__restrict__ csc_val, float *__restrict__ temp_all, float w_jac, float w_od,
               float *__restrict__ out, int64_t ld_out) {
  const int64_t N = Q + G;
  float *temp = temp_all + static_cast<int64_t>(blockIdx.x) * N;
  for (int64_t i = blockIdx.x; i < Q; i += gridDim.x) {
    for (int64_t r = threadIdx.x; r < N; r += THREADS) temp[r] = 0.f;
    __syncthreads();
    const int n = v_cnt[i];
    for (int e = 0; e < n; ++e) {
      const int32_t c = v_idx[i * cap + e];
      const float vi = v_val[i * cap + e];
      const int64_t b = col_off[c], en = col_off[c + 1];
      for (int64_t p = b + threadIdx.x; p < en; p += THREADS) {
        const int32_t r = csc_row[p];
        temp[r] = __fadd_rn(temp[r], fminf(vi, csc_val[p]));
      }
      __syncthreads();
    }
    const float *odi = od + i * N + Q;
    for (int64_t j = threadIdx.x; j < G; j += THREADS) {
      const float t = temp[Q + j];
      const float jac = __fsub_rn(1.0f, __fdiv_rn(t, __fsub_rn(2.0f, t)));
      out[i * ld_out + j] = __fadd_rn(__fmul_rn(jac, w_jac), __fmul_rn(odi[j], w_od));
    }
    __syncthreads();
  }
}

}  // namespace

// All pointers are DEVICE pointers (capi.cu stages host operands).  Workspaces are taken from the
// context.  k1 <= 28, 1 <= k2 <= 8.
int launch_rerank(dali_ctx *ctx, const float *qg, int64_t ld_qg, const float *qq, int64_t ld_qq,
                  const float *gg, int64_t ld_gg, int64_t Q, int64_t G, int k1, int k2, double lambda,
                  float *out, int64_t ld_out) {
  if (Q == 0 || G == 0) return DALI_OK;
  const int64_t N = Q + G;
  const int K = k1 + 1;
  const int H = static_cast<int>(std::nearbyint(k1 / 2.0)) + 1;  // np.around: half to even
  if (k1 < 1 || k1 > 28 || k2 < 1 || k2 > 8 || k2 > K || K > N)
    return set_err(ctx, DALI_ERR_INVALID, "re_ranking: 1 <= k1 <= 28, 1 <= k2 <= min(8, k1 + 1), k1 < Q + G");
  if (N > INT32_MAX) return set_err(ctx, DALI_ERR_UNSUPPORTED, "re_ranking: too many samples");
  cudaStream_t st = ctx->stream;
  void *p;
  int rc;
  // workspaces
  if ((rc = ws_ensure(ctx, WS_RR_OD, sizeof(float) * N * N, &p))) return rc;
  float *od = static_cast<float *>(p);
  if ((rc = ws_ensure(ctx, WS_RR_MISC, sizeof(float) * N + sizeof(int32_t) * 3 * N + sizeof(int64_t) * (N + 1) + 64, &p))) return rc;
  uint32_t *cmax = static_cast<uint32_t *>(p);
  int32_t *v0_cnt = reinterpret_cast<int32_t *>(cmax + N);
  int32_t *v_cnt = v0_cnt + N;
  int32_t *col_cnt = v_cnt + N;
  int64_t *col_off = reinterpret_cast<int64_t *>((reinterpret_cast<uintptr_t>(col_cnt + N) + 7) & ~uintptr_t(7));
  if ((rc = ws_ensure(ctx, WS_RR_RANK, (sizeof(int32_t) + sizeof(float)) * N * K, &p))) return rc;
  int32_t *rank = static_cast<int32_t *>(p);
  float *rank_d = reinterpret_cast<float *>(rank + N * K);
  if ((rc = ws_ensure(ctx, WS_RR_V0, (sizeof(int32_t) + sizeof(float)) * N * kCap1, &p))) return rc;
  int32_t *v0_idx = static_cast<int32_t *>(p);
  float *v0_val = reinterpret_cast<float *>(v0_idx + N * kCap1);
  const int cap2 = k2 == 1 ? kCap1 : k2 * kCap1;
  int32_t *v_idx = v0_idx;
  float *v_val = v0_val;
  if (k2 > 1) {
    if ((rc = ws_ensure(ctx, WS_RR_V, (sizeof(int32_t) + sizeof(float)) * N * cap2, &p))) return rc;
    v_idx = static_cast<int32_t *>(p);
    v_val = reinterpret_cast<float *>(v_idx + N * cap2);
  }

  // 1. od
  DALI_CUDA_OK(ctx, cudaMemsetAsync(cmax, 0, sizeof(uint32_t) * N, st));
  {
    KTimer t(ctx, DALI_K_RERANK);
    const dim3 blk(32, 8);
    auto colmax = [&](const float *m, int64_t rows, int64_t cols, int64_t ld, uint32_t *dst) {
      const dim3 grid(static_cast<unsigned>((cols + 31) / 32), static_cast<unsigned>(std::min<int64_t>((rows + 255) / 256, 64)));
      colmax_sq_kernel<<<grid, blk, 0, st>>>(m, rows, cols, ld, dst);
    };
    colmax(qq, Q, Q, ld_qq, cmax);
    rowmax_sq_kernel<<<static_cast<unsigned>((Q + 3) / 4), 128, 0, st>>>(qg, Q, G, ld_qg, cmax);
    colmax(qg, Q, G, ld_qg, cmax + Q);
    colmax(gg, G, G, ld_gg, cmax + Q);
    const float *cm = reinterpret_cast<const float *>(cmax);
    auto block = [&](const float *src, int64_t sld, int tr, int64_t rows, int64_t cols, float *dst, const float *cmr) {
      const dim3 grid(static_cast<unsigned>((cols + kOdCols - 1) / kOdCols), static_cast<unsigned>((rows + 31) / 32));
      od_block_kernel<<<grid, blk, 0, st>>>(src, sld, tr, rows, cols, dst, N, cmr);
    };
    block(qq, ld_qq, 1, Q, Q, od, cm);                     // rows < Q, cols < Q : qq^T
    block(qg, ld_qg, 0, Q, G, od + Q, cm);                 // rows < Q, cols >= Q: qg
    block(qg, ld_qg, 1, G, Q, od + Q * N, cm + Q);         // rows >= Q, cols < Q: qg^T
    block(gg, ld_gg, 1, G, G, od + Q * N + Q, cm + Q);     // rows >= Q, cols >= Q: gg^T
    DALI_CUDA_OK(ctx, cudaGetLastError());
  }
  // 2. initial_rank
  rc = launch_topk(ctx, od, N, N, N, K, 0, nullptr, 0, rank_d, rank);
  if (rc) return rc;
  {
    KTimer t(ctx, DALI_K_RERANK);
    // 3. V0
    krecip_kernel<<<static_cast<unsigned>((N + 3) / 4), 128, 0, st>>>(od, N, rank, K, H, v0_idx, v0_val, v0_cnt);
    // 4. V
    const int32_t *cnt_final = v0_cnt;
    if (k2 > 1) {
      int np2 = 2;
      while (np2 < cap2) np2 <<= 1;
      const size_t smem = sizeof(uint64_t) * np2 + sizeof(float) * cap2;
      if (int rc2 = ensure_dyn_smem(ctx, reinterpret_cast<const void *>(&qexpand_kernel), smem)) return rc2;
      qexpand_kernel<<<static_cast<unsigned>(N), 128, smem, st>>>(N, rank, K, k2, v0_idx, v0_val, v0_cnt, cap2,
                                                                 v_idx, v_val, v_cnt);
      cnt_final = v_cnt;
    }
    // 5. inverted index
    DALI_CUDA_OK(ctx, cudaMemsetAsync(col_cnt, 0, sizeof(int32_t) * N, st));
    csc_count_kernel<<<static_cast<unsigned>(N), 128, 0, st>>>(N, v_idx, cnt_final, cap2, col_cnt);
    scan_kernel<<<1, 1024, 0, st>>>(col_cnt, N, col_off);
    DALI_CUDA_OK(ctx, cudaGetLastError());
    // the index holds sum(v_cnt) <= N * cap2 entries (its exact size, col_off[N], stays on the device)
    const int64_t nnz_cap = N * static_cast<int64_t>(cap2);
    if ((rc = ws_ensure(ctx, WS_RR_CSC, (sizeof(int32_t) + sizeof(float)) * nnz_cap, &p))) return rc;
    int32_t *csc_row = static_cast<int32_t *>(p);
    float *csc_val = reinterpret_cast<float *>(csc_row + nnz_cap);
    DALI_CUDA_OK(ctx, cudaMemsetAsync(col_cnt, 0, sizeof(int32_t) * N, st));
    csc_fill_kernel<<<static_cast<unsigned>(N), 128, 0, st>>>(N, v_idx, v_val, cnt_final, cap2, col_off, col_cnt,
                                                              csc_row, csc_val);
    // 6. Jaccard + final
    // (every column of a query costs a chain of dependent global loads and a barrier: what hides
    // them is the number of resident CTAs -- 2 / 4 / 8 CTAs of 256 threads per SM and 16 of 128 measured at the Market shape)
    static const char *env_j = getenv("DALI_RR_CTAS_PER_SM");
    static const char *env_jt = getenv("DALI_RR_JACCARD_THREADS");
    const int jthreads = env_jt && atoi(env_jt) == 128 ? 128 : 256;
    const int per_sm = env_j ? std::max(1, std::min(2048 / jthreads, atoi(env_j))) : 2048 / jthreads;
    const int ctas = static_cast<int>(std::min<int64_t>(Q, static_cast<int64_t>(per_sm) * ctx->num_sms));
    if ((rc = ws_ensure(ctx, WS_RR_TEMP, sizeof(float) * N * ctas, &p))) return rc;
    const float w_jac = static_cast<float>(1.0 - lambda), w_od = static_cast<float>(lambda);
    if (jthreads == 256)
      jaccard_kernel<256><<<ctas, 256, 0, st>>>(Q, G, od, v_idx, v_val, cnt_final, cap2, col_off, csc_row, csc_val,
                                                static_cast<float *>(p), w_jac, w_od, out, ld_out);
    else
      jaccard_kernel<128><<<ctas, 128, 0, st>>>(Q, G, od, v_idx, v_val, cnt_final, cap2, col_off, csc_row, csc_val,
                                                static_cast<float *>(p), w_jac, w_od, out, ld_out);
    DALI_CUDA_OK(ctx, cudaGetLastError());
  }
  return DALI_OK;
}

}  // namespace dali
