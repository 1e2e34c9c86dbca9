// Full per-row ordering of a [Q,G] fp32 matrix (SURVEY 8f row N4, and the `indices` matrix of a5 for
// callers that want the whole ranked list, e.g. to display it).
//
// replaces  torch.argsort(total_similarities, dim=1, descending=True)[0]   getFeatures.py:303, 347
//           np.argsort(distmat, axis=1)                                    (torchreid eval)
//
// One device-wide LSD radix sort over composite 64-bit keys (row << 32 | order-preserving image of the
// fp32 value) with the column number as payload: the whole matrix is one flat problem, so a single
// long row (get_subset: 1 x N) and many short rows use the GPU equally well.  The sort is stable and
// the payload starts in ascending column order, hence ties come out by ascending column -- the
// order torch.argsort(stable=True) produces; NaN sorts after +inf (before everything when
// `descending`), -0 == +0.  The radix passes are CUB's (cub::DeviceRadixSort); only the bits that
// can differ are sorted (32 value bits + ceil(log2 Q) row bits).
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace dali {

namespace {

__global__ void __launch_bounds__(256)
sort_keys_kernel(const float *__restrict__ d, int64_t ld, int64_t Q, int64_t G, int descending,
                 uint64_t *__restrict__ keys, int32_t *__restrict__ vals) {
  const int64_t g = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (g >= G) return;
  for (int64_t q = blockIdx.y; q < Q; q += gridDim.y) {
    const float x = d[q * ld + g] + 0.0f;  // -0 -> +0
    uint32_t u = __float_as_uint(x);
    u = isnan(x) ? 0xFFFFFFFFu : (u ^ ((u >> 31) ? 0xFFFFFFFFu : 0x80000000u));
    if (descending) u = ~u;
    keys[q * G + g] = (static_cast<uint64_t>(q) << 32) | u;
    vals[q * G + g] = static_cast<int32_t>(g);
  }
}

}  // namespace

// idx_out: device int32 [Q, G] contiguous
int launch_argsort_rows(dali_ctx *ctx, const float *dist, int64_t Q, int64_t G, int64_t ld,
                        int descending, int32_t *idx_out) {
  if (Q == 0 || G == 0) return DALI_OK;
  if (G > 0x7fffffffLL || Q > 0x7fffffffLL) return set_err(ctx, DALI_ERR_UNSUPPORTED, "argsort: more than 2^31 rows or columns");
  KTimer timer(ctx, DALI_K_TOPK);
  const int64_t n = Q * G;
  int row_bits = 0;
  while ((1LL << row_bits) < Q) ++row_bits;
  size_t temp_bytes = 0;
  uint64_t *k0 = nullptr;
  int32_t *v0 = nullptr;
  DALI_CUDA_OK(ctx, cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, k0, k0, v0, v0, n, 0, 32 + row_bits,
                                                    ctx->stream));
  const size_t kb = (sizeof(uint64_t) * n + 255) / 256 * 256, vb = (sizeof(int32_t) * n + 255) / 256 * 256;
  void *p;
  int rc = ws_ensure(ctx, WS_SORT, 2 * kb + vb + temp_bytes, &p);
  if (rc) return rc;
  char *base = static_cast<char *>(p);
  uint64_t *keys_in = reinterpret_cast<uint64_t *>(base), *keys_out = reinterpret_cast<uint64_t *>(base + kb);
  int32_t *vals_in = reinterpret_cast<int32_t *>(base + 2 * kb);
  void *temp = base + 2 * kb + vb;
  const unsigned gx = static_cast<unsigned>((G + 255) / 256);
  dim3 grid(gx, static_cast<unsigned>(std::min<int64_t>(Q, std::max<int64_t>(1, 16 * ctx->num_sms / gx))));
  ctx->launches++;
  sort_keys_kernel<<<grid, 256, 0, ctx->stream>>>(dist, ld, Q, G, descending, keys_in, vals_in);
  DALI_CUDA_OK(ctx, cudaGetLastError());
  ctx->launches += 1 + (32 + row_bits + 7) / 8;
  DALI_CUDA_OK(ctx, cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys_in, keys_out, vals_in, idx_out, n, 0,
                                                    32 + row_bits, ctx->stream));
  return DALI_OK;
}

}  // namespace dali
