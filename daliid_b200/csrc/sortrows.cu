// Full per-row ordering of a [Q,G] fp32 matrix (SURVEY 8f row N4, and the `indices` matrix of a5 for
// callers that want the whole ranked list, e.g. to display it).
//
// replaces  torch.argsort(total_similarities, dim=1, descending=True)[0]   getFeatures.py:303, 347
//           np.argsort(distmat, axis=1)                                    (torchreid eval)
//
// A segmented LSD radix sort written for this layout (round 2; round 1 called cub::DeviceRadixSort
// on 64-bit composite keys): every row is its own segment, keys are the 32-bit order-preserving
// images of the fp32 values, the payload is the column number.  Four passes of 8 bits, three
// launches each, over a grid of (tiles of a row) x (rows), so a single long row (get_subset: 1 x N)
// and many short rows load the GPU alike:
//   1. digit_hist_kernel   per tile: how many of its keys carry each of the 256 digit values;
//   2. digit_scan_kernel   per row: exclusive prefix over (digit, tile) -- digit-major, which is
//                          what makes the pass stable across tiles;
//   3. digit_scatter_kernel per tile: the stable rank of every key among the tile's keys of the same
//                          digit (warp-striped items; __match_any_sync gives the rank inside one
//                          32-key step, a per-warp counter the ranks of the earlier steps, a prefix
//                          over the warps the rest), key and column written to their place.
// Pass 0 reads the fp32 matrix itself (any leading dimension) and forms the keys on the fly; the
// last pass writes the columns straight into the caller's index matrix.  The sort is stable and
// the payload starts in ascending column order, hence ties come out by ascending column -- the
// order torch.argsort(stable=True) produces; NaN sorts after +inf (before everything when
// `descending`), -0 == +0.  Rows are processed in batches so that the (rows x 256 x tiles) counters
// stay below 64 MB.
#include <algorithm>

#include "common.cuh"

namespace dali {

namespace {

constexpr int kSortThreads = 256;
constexpr int kSortItems = 16;                           // keys per thread
constexpr int kSortTile = kSortThreads * kSortItems;     // keys per CTA
constexpr int kSortWarps = kSortThreads / 32;

__device__ __forceinline__ uint32_t sort_key(float v, int descending) {
  const float x = v + 0.0f;  // -0 -> +0
  uint32_t u = __float_as_uint(x);
  u = (x != x) ? 0xFFFFFFFFu : (u ^ ((u >> 31) ? 0xFFFFFFFFu : 0x80000000u));
  return descending ? ~u : u;
}

struct SortArgs {
  const float *dist;       // pass 0 source (row pitch ld)
  int64_t ld;
  const uint32_t *kin;     // later passes: keys / columns of the previous pass, [rows][G]
  const int32_t *vin;
  uint32_t *kout;          // [rows][G]; null on the last pass
  int32_t *vout;           // [rows][G]
  uint32_t *counts;        // [rows][256][tiles]
  int64_t G;
  int tiles, shift, first, descending;
};

// element e of the tile handled by (warp w, step j, lane l): w * 32 * ITEMS + j * 32 + l -- coalesced,
// and ascending in (w, j, l), the order the stable ranks are counted in
__device__ __forceinline__ int64_t item_col(int tile, int w, int j, int lane) {
  return static_cast<int64_t>(tile) * kSortTile + w * (32 * kSortItems) + j * 32 + lane;
}

__global__ void __launch_bounds__(kSortThreads) digit_hist_kernel(SortArgs a) {
  __shared__ uint32_t h[256];
  const int tile = blockIdx.x, tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  const int64_t row = blockIdx.y;
  h[tid] = 0u;
  __syncthreads();
#pragma unroll 4
  for (int j = 0; j < kSortItems; ++j) {
    const int64_t c = item_col(tile, w, j, lane);
    if (c < a.G) {
      const uint32_t k = a.first ? sort_key(a.dist[row * a.ld + c], a.descending) : a.kin[row * a.G + c];
      atomicAdd(&h[(k >> a.shift) & 255u], 1u);
    }
  }
  __syncthreads();
  a.counts[(row * 256 + tid) * a.tiles + tile] = h[tid];
}

// one CTA per row: exclusive prefix of counts[row][digit][tile] in (digit, tile) order, in place
__global__ void __launch_bounds__(256) digit_scan_kernel(uint32_t *counts, int tiles) {
  __shared__ uint32_t tot[256];
  const int d = threadIdx.x;
  uint32_t *c = counts + (static_cast<int64_t>(blockIdx.x) * 256 + d) * tiles;
  uint32_t s = 0;
  for (int t = 0; t < tiles; ++t) s += c[t];
  tot[d] = s;
  __syncthreads();
  // exclusive prefix over the 256 digit totals (Hillis-Steele in shared memory)
  uint32_t incl = s;
  for (int o = 1; o < 256; o <<= 1) {
    const uint32_t add = d >= o ? tot[d - o] : 0u;
    __syncthreads();
    incl += add;
    tot[d] = incl;
    __syncthreads();
  }
  uint32_t run = incl - s;
  for (int t = 0; t < tiles; ++t) {
    const uint32_t x = c[t];
    c[t] = run;
    run += x;
  }
}

__global__ void __launch_bounds__(kSortThreads) digit_scatter_kernel(SortArgs a) {
  __shared__ uint32_t wc[kSortWarps][256];  // per warp: keys of each digit seen so far, then the prefix
  const int tile = blockIdx.x, tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  const int64_t row = blockIdx.y;
  for (int i = tid; i < kSortWarps * 256; i += kSortThreads) (&wc[0][0])[i] = 0u;
  __syncthreads();
  uint32_t key[kSortItems], rank[kSortItems];
  int32_t val[kSortItems];
#pragma unroll
  for (int j = 0; j < kSortItems; ++j) {
    const int64_t c = item_col(tile, w, j, lane);
    const bool ok = c < a.G;
    key[j] = 0u;
    val[j] = 0;
    if (ok) {
      if (a.first) {
        key[j] = sort_key(a.dist[row * a.ld + c], a.descending);
        val[j] = static_cast<int32_t>(c);
      } else {
        key[j] = a.kin[row * a.G + c];
        val[j] = a.vin[row * a.G + c];
      }
    }
    // lanes beyond the row take part in the vote with a value no key digit can equal
    const uint32_t dg = ok ? ((key[j] >> a.shift) & 255u) : (256u + lane);
    const uint32_t peers = __match_any_sync(0xffffffffu, dg);
    const uint32_t below = __popc(peers & ((1u << lane) - 1u));
    uint32_t base = 0;
    if (ok) base = wc[w][dg];
    __syncwarp();
    if (ok && below == 0) wc[w][dg] = base + __popc(peers);  // the lowest lane of the group
    __syncwarp();
    rank[j] = base + below;
  }
  __syncthreads();
  {  // prefix over the warps, per digit (thread = digit), plus the tile's global offset
    const uint32_t goff = a.counts[(row * 256 + tid) * a.tiles + tile];
    uint32_t run = goff;
#pragma unroll
    for (int ww = 0; ww < kSortWarps; ++ww) {
      const uint32_t x = wc[ww][tid];
      wc[ww][tid] = run;
      run += x;
    }
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < kSortItems; ++j) {
    const int64_t c = item_col(tile, w, j, lane);
    if (c < a.G) {
      const uint32_t dg = (key[j] >> a.shift) & 255u;
      const int64_t dst = row * a.G + wc[w][dg] + rank[j];
      if (a.kout) a.kout[dst] = key[j];
      a.vout[dst] = val[j];
    }
  }
}

}  // namespace

// idx_out: device int32 [Q, G] contiguous
int launch_argsort_rows(dali_ctx *ctx, const float *dist, int64_t Q, int64_t G, int64_t ld,
                        int descending, int32_t *idx_out) {
  if (Q == 0 || G == 0) return DALI_OK;
  if (G > 0x7fffffffLL || Q > 0x7fffffffLL) return set_err(ctx, DALI_ERR_UNSUPPORTED, "argsort: more than 2^31 rows or columns");
  KTimer timer(ctx, DALI_K_TOPK);
  const int64_t tiles64 = (G + kSortTile - 1) / kSortTile;
  if (tiles64 > 65535 * 16) return set_err(ctx, DALI_ERR_UNSUPPORTED, "argsort: row too long");
  const int tiles = static_cast<int>(tiles64);
  // rows per batch: counters below 64 MB, grid.y below 65536
  const int64_t per_row = 256ll * tiles * sizeof(uint32_t);
  const int64_t batch = std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(Q, 65535), (64ll << 20) / per_row));
  const size_t kb = (sizeof(uint32_t) * batch * G + 255) / 256 * 256;
  void *p;
  int rc = ws_ensure(ctx, WS_SORT, 4 * kb + static_cast<size_t>(batch) * per_row, &p);
  if (rc) return rc;
  char *base = static_cast<char *>(p);
  uint32_t *kbuf[2] = {reinterpret_cast<uint32_t *>(base), reinterpret_cast<uint32_t *>(base + kb)};
  int32_t *vbuf[2] = {reinterpret_cast<int32_t *>(base + 2 * kb), reinterpret_cast<int32_t *>(base + 3 * kb)};
  uint32_t *counts = reinterpret_cast<uint32_t *>(base + 4 * kb);
  for (int64_t r0 = 0; r0 < Q; r0 += batch) {
    const int64_t rows = std::min(batch, Q - r0);
    const dim3 grid(static_cast<unsigned>(tiles), static_cast<unsigned>(rows));
    for (int pass = 0; pass < 4; ++pass) {
      SortArgs a{};
      a.dist = dist + r0 * ld; a.ld = ld;
      a.kin = kbuf[(pass + 1) & 1]; a.vin = vbuf[(pass + 1) & 1];
      a.kout = pass == 3 ? nullptr : kbuf[pass & 1];
      a.vout = pass == 3 ? idx_out + r0 * G : vbuf[pass & 1];
      a.counts = counts; a.G = G; a.tiles = tiles; a.shift = 8 * pass; a.first = pass == 0;
      a.descending = descending;
      digit_hist_kernel<<<grid, kSortThreads, 0, ctx->stream>>>(a);
      digit_scan_kernel<<<static_cast<unsigned>(rows), 256, 0, ctx->stream>>>(counts, tiles);
      digit_scatter_kernel<<<grid, kSortThreads, 0, ctx->stream>>>(a);
      ctx->launches += 3;
    }
    DALI_CUDA_OK(ctx, cudaGetLastError());
  }
  return DALI_OK;
}

}  // namespace dali
