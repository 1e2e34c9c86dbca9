// Gallery-sharded exchange over NVLink peer memory: the two exchange steps of the
// sharded evaluation -- match keys after the gather, per-positive counts after the counting pass --
// are sums of ~70k 32-bit words per rank.  They are latency, not bandwidth: an NCCL all-reduce
// costs a host-side launch plus ~20-60 us on the device each.  Here every rank keeps its
// contribution in a cudaMalloc'ed block that all peers map through CUDA IPC, and ONE kernel per
// exchange does flag signalling + waiting + the reduction with plain peer loads:
//
//   block layout (per rank):  flags[8] u32 (one per peer, monotonic epochs) | buf0[cap] | buf1[cap]
//   peer_allreduce_kernel:    CTA 0 publishes "my contribution for epoch e is complete" into every
//                             peer's flag word (st.release.sys); every CTA waits until all peers'
//                             words reached e (ld.acquire.sys), then out[i] = sum_r peer_r.buf[i]
//
// Stream order makes the contribution complete before the kernel starts; the next overwrite of a
// buffer is two epochs later and therefore after every peer finished reading it.
// One process per GPU; kernels of different ranks run on different GPUs (never on one device).
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

struct dali_peer {
  dali_ctx *ctx = nullptr;
  int rank = 0, world = 1;
  int64_t cap = 0;          // elements per buffer
  size_t bytes = 0;
  void *local = nullptr;    // this rank's block
  void *peers[8] = {nullptr};  // mapped blocks of every rank (peers[rank] == local)
  bool connected = false;
  uint32_t epoch = 0;
};

namespace dali {

namespace {

constexpr int kFlagBytes = 256;
constexpr int kStatusWord = 32;  // flags[32]: first epoch whose wait ran out of time (0 = none)

struct PeerPtrs {
  const uint32_t *flags_local;   // this rank's flag words
  uint32_t *flags[8];            // every rank's flag block
  const int32_t *buf[8];         // every rank's contribution
  int world, rank;
  uint32_t *status;              // this rank's timeout word
  unsigned long long watchdog;   // clock64 ticks a wait may take
};

__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(256)
peer_allreduce_kernel(PeerPtrs pp, uint32_t epoch, int32_t *__restrict__ out, int64_t n) {
  if (blockIdx.x == 0 && threadIdx.x < pp.world) {
    __threadfence_system();  // the producer kernel's writes (stream-ordered before us) -> system scope
    st_release_sys(pp.flags[threadIdx.x] + pp.rank, epoch);
  }
  if (threadIdx.x < pp.world) {
    const unsigned long long t0 = clock64();
    // epochs are compared modulo 2^32 (signed difference): they only ever differ by a few
    // A peer that never arrives (a rank died, or is stuck in first-call allocations for longer
    // than the limit) must not hang the GPU, and a trap would poison the whole context: the wait
    // gives up, records the epoch in the status word and lets the kernel finish (its sums are then
    // meaningless); the host reads the word with the results and reports DALI_ERR_PEER_TIMEOUT.
    while (static_cast<int32_t>(ld_acquire_sys(pp.flags_local + threadIdx.x) - epoch) < 0) {
      if (clock64() - t0 > pp.watchdog) {
        atomicCAS(pp.status, 0u, epoch);
        break;
      }
    }
  }
  __syncthreads();
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    int32_t s = 0;
#pragma unroll 8
    for (int r = 0; r < pp.world; ++r) s += pp.buf[r][i];
    out[i] = s;
  }
}

}  // namespace

}  // namespace dali

using namespace dali;

extern "C" {

int dali_peer_create(dali_ctx *ctx, int rank, int world, int64_t capacity, dali_peer **out) {
  if (!ctx || !out || world < 1 || world > 8 || rank < 0 || rank >= world || capacity < 1)
    return set_err(ctx, DALI_ERR_INVALID, "peer_create: 1 <= world <= 8, 0 <= rank < world, capacity >= 1");
  DeviceGuard dg;
  if (int rc = dg.enter(ctx)) return rc;
  dali_peer *p = new dali_peer();
  p->ctx = ctx; p->rank = rank; p->world = world;
  p->cap = (capacity + 63) / 64 * 64;
  p->bytes = kFlagBytes + 2 * sizeof(int32_t) * p->cap;
  cudaError_t e = cudaMalloc(&p->local, p->bytes);  // IPC needs a cudaMalloc allocation
  if (e != cudaSuccess) {
    delete p;
    return set_err(ctx, DALI_ERR_NOMEM, std::string("peer block: ") + cudaGetErrorString(e));
  }
  cudaMemset(p->local, 0, p->bytes);
  cudaDeviceSynchronize();
  p->peers[rank] = p->local;
  p->connected = world == 1;
  *out = p;
  return DALI_OK;
}

int dali_peer_ipc_handle(dali_peer *p, void *handle64) {
  if (!p || !handle64) return DALI_ERR_INVALID;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t h;
  DeviceGuard dg;
  if (int rc = dg.enter(p->ctx)) return rc;
  DALI_CUDA_OK(p->ctx, cudaIpcGetMemHandle(&h, p->local));
  std::memcpy(handle64, &h, 64);
  return DALI_OK;
}

int dali_peer_connect(dali_peer *p, const void *handles) {
  if (!p || !handles) return DALI_ERR_INVALID;
  DeviceGuard dg;
  if (int rc = dg.enter(p->ctx)) return rc;
  for (int r = 0; r < p->world; ++r) {
    if (r == p->rank) continue;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, static_cast<const char *>(handles) + 64 * r, 64);
    void *ptr = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess)
      return set_err(p->ctx, DALI_ERR_CUDA, std::string("cudaIpcOpenMemHandle(rank ") + std::to_string(r) +
                                                "): " + cudaGetErrorString(e));
    p->peers[r] = ptr;
  }
  p->connected = true;
  return DALI_OK;
}

void dali_peer_destroy(dali_peer *p) {
  if (!p) return;
  DeviceGuard dg;
  dg.enter(p->ctx);
  cudaStreamSynchronize(p->ctx->stream);
  for (int r = 0; r < p->world; ++r)
    if (r != p->rank && p->peers[r]) cudaIpcCloseMemHandle(p->peers[r]);
  if (p->local) cudaFree(p->local);
  delete p;
}

int64_t dali_peer_capacity(const dali_peer *p) { return p ? p->cap : 0; }

void *dali_peer_buffer(dali_peer *p, int which) {
  if (!p || which < 0 || which > 1) return nullptr;
  return static_cast<char *>(p->local) + kFlagBytes + sizeof(int32_t) * p->cap * which;
}

int dali_peer_status(dali_peer *p) {
  if (!p) return DALI_ERR_INVALID;
  DeviceGuard dg;
  if (int rc = dg.enter(p->ctx)) return rc;
  uint32_t word = 0;
  DALI_CUDA_OK(p->ctx, cudaMemcpyAsync(&word, reinterpret_cast<uint32_t *>(p->local) + kStatusWord, sizeof(word),
                                       cudaMemcpyDeviceToHost, p->ctx->stream));
  DALI_CUDA_OK(p->ctx, cudaStreamSynchronize(p->ctx->stream));
  if (word)
    return set_err(p->ctx, DALI_ERR_PEER_TIMEOUT,
                   "peer exchange " + std::to_string(word) + " timed out waiting for another rank (DALI_PEER_TIMEOUT_MS)");
  return DALI_OK;
}

int dali_peer_allreduce_i32(dali_ctx *ctx, dali_peer *p, int which, int32_t *out, int64_t n) {
  if (!ctx || !p || !out || which < 0 || which > 1 || n < 0 || n > p->cap || p->ctx != ctx)
    return set_err(ctx, DALI_ERR_INVALID, "peer_allreduce: bad arguments");
  if (!p->connected) return set_err(ctx, DALI_ERR_INVALID, "peer_allreduce: peers not connected");
  DeviceGuard dg;
  if (int rc = dg.enter(ctx)) return rc;
  PeerPtrs pp;
  pp.world = p->world; pp.rank = p->rank;
  pp.flags_local = static_cast<const uint32_t *>(p->local);
  for (int r = 0; r < 8; ++r) {
    char *base = static_cast<char *>(r < p->world ? p->peers[r] : p->local);
    pp.flags[r] = reinterpret_cast<uint32_t *>(base);
    pp.buf[r] = reinterpret_cast<const int32_t *>(base + kFlagBytes + sizeof(int32_t) * p->cap * which);
  }
  pp.status = reinterpret_cast<uint32_t *>(p->local) + kStatusWord;
  {
    static const char *env = getenv("DALI_PEER_TIMEOUT_MS");  // default 20 s
    const double ms = env ? std::max(1.0, atof(env)) : 20000.0;
    pp.watchdog = static_cast<unsigned long long>(ms * 1e-3 * ctx->clock_khz * 1e3);
  }
  p->epoch += 1;
  const int blocks = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, 2 * ctx->num_sms)));
  KTimer t(ctx, DALI_K_PEER_EXCHANGE);
  peer_allreduce_kernel<<<blocks, 256, 0, ctx->stream>>>(pp, p->epoch, out, n);
  DALI_CUDA_OK(ctx, cudaGetLastError());
  return DALI_OK;
}

}  // extern "C"
