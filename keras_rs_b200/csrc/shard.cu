// shard.cu — row-wise MOD shard routing and peer-memory setup for row-sharded tables (config C5).
//
// The reference only shards tables on TPU SparseCore, with MOD layout: row r lives on shard r % S at
// local row r / S (jax/embedding_utils.py:187-197 sharding_strategy="MOD";
// tensorflow/distributed_embedding.py:316-328).  On one NVSwitch box every GPU can load from and
// atomically add to every peer's HBM, so the B200 design needs no index exchange at all: shards are
// cudaMalloc'd, exported with cudaIpc and the fused gather / scatter kernels (gather.cu) address
// `shard_tables[r % S] + (r / S) * E` directly — rows cross NVLink exactly once, tile by tile, inside
// the same kernel that writes the concatenated activation.  krs_mod_route is the explicit routing
// (used by the NCCL baseline exchange and by the bit-exact integer parity tests).
#include "common.cuh"

namespace krs {
namespace {
template <typename IdT>
__global__ void mod_route_kernel(const IdT* __restrict__ ids, int64_t n, int S, int32_t* __restrict__ owner,
                                 int64_t* __restrict__ local, int32_t* __restrict__ counts) {
  __shared__ int hist[64];
  for (int i = threadIdx.x; i < 64; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t id = (int64_t)ids[i];
    const int o = (int)(id % S);
    owner[i] = o;
    local[i] = id / S;
    if (counts) atomicAdd(&hist[o], 1);
  }
  __syncthreads();
  if (counts)
    for (int i = threadIdx.x; i < S; i += blockDim.x)
      if (hist[i]) atomicAdd(counts + i, hist[i]);
}
}  // namespace
}  // namespace krs

using namespace krs;

extern "C" int krs_mod_route(const void* ids, int ids_i64, int64_t n, int S, int32_t* owner, int64_t* local,
                             int32_t* counts, void* stream) {
  KRS_REQUIRE(ids && owner && local, "krs_mod_route: null argument");
  KRS_REQUIRE(S >= 1 && S <= 64, "krs_mod_route: num_shards must be in 1..64, got %d", S);
  if (n == 0) return KRS_OK;
  const unsigned grid = (unsigned)krs::imax<int64_t>(1, krs::imin<int64_t>(ceil_div<int64_t>(n, 256), (int64_t)sm_count() * 8));
  if (ids_i64) mod_route_kernel<int64_t><<<grid, 256, 0, as_stream(stream)>>>((const int64_t*)ids, n, S, owner, local, counts);
  else mod_route_kernel<int32_t><<<grid, 256, 0, as_stream(stream)>>>((const int32_t*)ids, n, S, owner, local, counts);
  KRS_LAUNCH_CHECK();
  return KRS_OK;
}

extern "C" int krs_ipc_alloc(void** dptr, size_t bytes, void* handle_out_64B) {
  KRS_REQUIRE(dptr && handle_out_64B && bytes > 0, "krs_ipc_alloc: bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  KRS_CUDA(cudaMalloc(dptr, bytes));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, *dptr);
  if (e != cudaSuccess) {
    cudaFree(*dptr);
    *dptr = nullptr;
    return fail_cuda(e, "cudaIpcGetMemHandle", __FILE__, __LINE__);
  }
  memcpy(handle_out_64B, &h, 64);
  return KRS_OK;
}
extern "C" int krs_ipc_open(const void* handle_64B, void** dptr) {
  KRS_REQUIRE(handle_64B && dptr, "krs_ipc_open: bad argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle_64B, 64);
  KRS_CUDA(cudaIpcOpenMemHandle(dptr, h, cudaIpcMemLazyEnablePeerAccess));
  return KRS_OK;
}
extern "C" int krs_ipc_close(void* dptr) {
  KRS_CUDA(cudaIpcCloseMemHandle(dptr));
  return KRS_OK;
}
extern "C" int krs_ipc_free(void* dptr) {
  KRS_CUDA(cudaFree(dptr));
  return KRS_OK;
}
extern "C" int krs_enable_peer_access(int peer_device) {
  int dev = 0;
  KRS_CUDA(cudaGetDevice(&dev));
  if (peer_device == dev) return KRS_OK;
  int can = 0;
  KRS_CUDA(cudaDeviceCanAccessPeer(&can, dev, peer_device));
  KRS_REQUIRE(can, "krs_enable_peer_access: device %d cannot access peer %d", dev, peer_device);
  cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) {
    cudaGetLastError();
    return KRS_OK;
  }
  KRS_CUDA(e);
  return KRS_OK;
}
