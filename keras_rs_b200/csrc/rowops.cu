// rowops.cu — row-wise selection / masking kernels of the retrieval path.
//
//  * krs_row_topk: exact top-k of every row of a (rows, n) matrix, values descending, ties -> lowest index first (the
//    jax.lax.top_k order keras.ops.top_k has on the JAX backend).  Used to merge the per-shard lists of the
//    candidate-sharded BruteForceRetrieval (examples/data_parallel_retrieval.py:145-165 is the multi-GPU caller) and, with a
//    label boost, as HardNegativeMining (hard_negative_mining.py:43-94: top-(k+1) of logits + labels * MAX_FLOAT, then
//    take_along_axis of logits and labels).
//  * krs_remove_accidental_hits (remove_accidental_hits.py:32-97) and krs_sampling_prob_correction
//    (sampling_probability_correction.py:39-58): one pass over the logits.
// A row is sorted inside ONE CTA: 64-bit keys (order-preserving image of the fp32 value << 32 | ~index) in shared memory,
// bitonic network, n <= 16384 (128 KB of keys).  HBM traffic = the row once in, k results out.
#include <float.h>

#include "common.cuh"

namespace krs {
namespace {

constexpr int ROW_MAX_N = 16384;

__device__ __forceinline__ uint32_t ord_key(float v) {          // larger float -> larger unsigned (NaN sorts below everything)
  if (v != v) return 0u;
  const uint32_t u = __float_as_uint(v);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// boost (nullable): key value = x + boost * boost_scale (HardNegativeMining's `logits + labels * MAX_FLOAT`)
__global__ void __launch_bounds__(512) row_topk_kernel(const float* __restrict__ x, const float* __restrict__ boost, float boost_scale,
                                                       int64_t ld, int64_t boost_ld, int n, int npad, int k,
                                                       float* __restrict__ out_vals, int32_t* __restrict__ out_idx,
                                                       const float* __restrict__ gather2, int64_t gather2_ld,
                                                       float* __restrict__ out_gather2, const int32_t* __restrict__ gather_i,
                                                       int64_t gather_i_ld, int32_t* __restrict__ out_gather_i) {
  extern __shared__ unsigned long long keys[];
  const int64_t r = blockIdx.x;
  const float* row = x + r * ld;
  for (int i = threadIdx.x; i < npad; i += blockDim.x) {
    unsigned long long key = 0ull;                         // padding sorts last (a real key's low word ~i is never 0)
    if (i < n) {
      float v = row[i];
      if (boost) v = v + boost[r * boost_ld + i] * boost_scale;
      key = ((unsigned long long)ord_key(v) << 32) | (uint32_t)(~(uint32_t)i);
    }
    keys[i] = key;
  }
  __syncthreads();
  // bitonic sort, descending
  for (int size = 2; size <= npad; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < (npad >> 1); t += blockDim.x) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool desc = ((lo & size) == 0);
        const unsigned long long a = keys[lo], b = keys[hi];
        if ((a < b) == desc) { keys[lo] = b; keys[hi] = a; }
      }
      __syncthreads();
    }
  }
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    const unsigned long long key = keys[j];
    const int idx = (int)(~(uint32_t)(key & 0xffffffffu));
    out_idx[r * k + j] = idx;
    out_vals[r * k + j] = row[idx];                        // the unboosted value
    if (gather2) out_gather2[r * k + j] = gather2[r * gather2_ld + idx];
    if (gather_i) out_gather_i[r * k + j] = gather_i[r * gather_i_ld + idx];
  }
}

// dst (rows, n) = 0 ; dst[r, idx[r, j]] = g[r, j]   (backward of the row selection)
__global__ void __launch_bounds__(256) row_scatter_kernel(const float* __restrict__ g, const int32_t* __restrict__ idx, int64_t rows,
                                                          int k, int n, float* __restrict__ dst) {
  const int64_t total = rows * (int64_t)k;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / k;
    dst[r * n + idx[i]] = g[i];
  }
}

// remove_accidental_hits.py:84-97, literally: positive index = argmax(labels[r, :]) (first maximum), positive id =
// take(candidate_ids FLATTENED, positive index), duplicate = (ids[r or 0, j] == positive id) - labels[r, j],
// out = logits + duplicate * smallest.
template <typename IdT>
__global__ void __launch_bounds__(256) remove_hits_kernel(const float* __restrict__ logits, const float* __restrict__ labels,
                                                          const IdT* __restrict__ ids, int64_t ids_row_stride, int n, float smallest,
                                                          float* __restrict__ out) {
  __shared__ float s_val[8];
  __shared__ int s_idx[8];
  const int64_t r = blockIdx.x;
  const float* lab = labels + r * n;
  float best = -FLT_MAX;
  int bi = 0x7fffffff;
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    const float v = lab[j];
    if (v > best || (v == best && j < bi)) { best = v; bi = j; }
  }
  if (bi == 0x7fffffff) { bi = 0; best = -FLT_MAX; }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, d);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, d);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { s_val[warp] = best; s_idx[warp] = bi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
      if (s_val[w] > s_val[0] || (s_val[w] == s_val[0] && s_idx[w] < s_idx[0])) { s_val[0] = s_val[w]; s_idx[0] = s_idx[w]; }
  }
  __syncthreads();
  const IdT pos = ids[s_idx[0]];                              // flattened take
  const IdT* idr = ids + r * ids_row_stride;
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    const float dup = (idr[j] == pos ? 1.f : 0.f) - lab[j];
    out[r * n + j] = logits[r * n + j] + dup * smallest;
  }
}

__global__ void __launch_bounds__(256) prob_correction_kernel(const float* __restrict__ logits, const float* __restrict__ probs,
                                                              int64_t total, int64_t period, float eps, float* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const float p = fminf(fmaxf(probs[i % period], eps), 1.f);
    out[i] = logits[i] - logf(p);
  }
}

}  // namespace
}  // namespace krs

using namespace krs;

extern "C" int krs_row_topk(const float* x, int64_t rows, int n, int64_t ld, const float* boost, int64_t boost_ld, float boost_scale,
                            int k, float* out_vals, int32_t* out_idx, const float* gather2, int64_t gather2_ld, float* out_gather2,
                            const int32_t* gather_i32, int64_t gather_i32_ld, int32_t* out_gather_i32, void* stream) {
  KRS_REQUIRE(x && out_vals && out_idx, "krs_row_topk: null argument");
  KRS_REQUIRE(rows >= 0 && n >= 1 && n <= ROW_MAX_N && ld >= n, "krs_row_topk: need 1 <= n <= %d and ld >= n", ROW_MAX_N);
  KRS_REQUIRE(k >= 1 && k <= n, "krs_row_topk: need 1 <= k <= n");
  KRS_REQUIRE((gather2 == nullptr) == (out_gather2 == nullptr), "krs_row_topk: gather2 and out_gather2 go together");
  KRS_REQUIRE((gather_i32 == nullptr) == (out_gather_i32 == nullptr), "krs_row_topk: gather_i32 and out_gather_i32 go together");
  if (rows == 0) return KRS_OK;
  int npad = 2;
  while (npad < n) npad <<= 1;
  const size_t smem = (size_t)npad * sizeof(unsigned long long);
  if (smem > 48 * 1024) KRS_CUDA(cudaFuncSetAttribute(row_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int threads = npad / 2 < 64 ? 64 : (npad / 2 > 512 ? 512 : npad / 2);
  row_topk_kernel<<<(unsigned)rows, threads, smem, as_stream(stream)>>>(x, boost, boost_scale, ld, boost_ld, n, npad, k, out_vals, out_idx,
                                                                       gather2, gather2_ld, out_gather2, gather_i32, gather_i32_ld,
                                                                       out_gather_i32);
  KRS_LAUNCH_CHECK();
  return KRS_OK;
}

extern "C" int krs_row_scatter(const float* g, const int32_t* idx, int64_t rows, int k, int n, float* dst, void* stream) {
  KRS_REQUIRE(g && idx && dst && rows >= 0 && k >= 1 && n >= k, "krs_row_scatter: bad argument");
  if (rows == 0) return KRS_OK;
  cudaStream_t s = as_stream(stream);
  KRS_CUDA(cudaMemsetAsync(dst, 0, sizeof(float) * (size_t)rows * (size_t)n, s));
  const int64_t total = rows * (int64_t)k;
  row_scatter_kernel<<<(unsigned)krs::imax<int64_t>(1, krs::imin<int64_t>(ceil_div<int64_t>(total, 256), (int64_t)sm_count() * 16)), 256, 0, s>>>(
      g, idx, rows, k, n, dst);
  KRS_LAUNCH_CHECK();
  return KRS_OK;
}

extern "C" int krs_remove_accidental_hits(const float* logits, const float* labels, const void* candidate_ids, int ids_i64,
                                          int ids_per_row, int64_t rows, int n, float smallest, float* out, void* stream) {
  KRS_REQUIRE(logits && labels && candidate_ids && out && rows >= 0 && n >= 1, "krs_remove_accidental_hits: bad argument");
  if (rows == 0) return KRS_OK;
  const int64_t stride = ids_per_row ? n : 0;
  if (ids_i64)
    remove_hits_kernel<int64_t><<<(unsigned)rows, 256, 0, as_stream(stream)>>>(logits, labels, (const int64_t*)candidate_ids, stride, n, smallest, out);
  else
    remove_hits_kernel<int32_t><<<(unsigned)rows, 256, 0, as_stream(stream)>>>(logits, labels, (const int32_t*)candidate_ids, stride, n, smallest, out);
  KRS_LAUNCH_CHECK();
  return KRS_OK;
}

extern "C" int krs_sampling_prob_correction(const float* logits, const float* probs, int64_t total, int64_t probs_period, float eps,
                                            float* out, void* stream) {
  KRS_REQUIRE(logits && probs && out && total >= 0 && probs_period >= 1, "krs_sampling_prob_correction: bad argument");
  if (total == 0) return KRS_OK;
  prob_correction_kernel<<<(unsigned)krs::imax<int64_t>(1, krs::imin<int64_t>(ceil_div<int64_t>(total, 256), (int64_t)sm_count() * 16)), 256, 0,
                           as_stream(stream)>>>(logits, probs, total, probs_period, eps, out);
  KRS_LAUNCH_CHECK();
  return KRS_OK;
}
