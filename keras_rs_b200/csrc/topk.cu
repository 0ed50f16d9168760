// topk.cu — BruteForceRetrieval: streaming score + exact top-k.
//
// Replaces Retrieval.compute_score (retrieval.py:101-117: matmul(q, transpose(c))) followed by
// keras.ops.top_k and the optional ops.take(candidate_ids, top_ids) (brute_force_retrieval.py:139-143).
// The (nq, nc) score matrix (164 GB at C4) is never materialised: each CTA owns a 64-query tile and a
// slice of the candidates, computes 64x128 score tiles in registers (exact fp32 FMA), filters them
// against the running k-th best score of each query and merges the few survivors into per-query
// sorted lists held in shared memory.  A second kernel merges the per-slice lists.
// Ordering is total and deterministic: score descending, ties -> lowest candidate index first
// (jax.lax.top_k rule adopted in SURVEY.md Appendix A.3), independent of thread scheduling.
#include <limits.h>
#include <math.h>

#include "common.cuh"

namespace krs {
namespace {

constexpr int QT = 64;     // queries per CTA
constexpr int CT = 128;    // candidates per tile
constexpr int KP = 128;    // list capacity (k <= KP)
constexpr int MAXD = 128;
constexpr int MERGE_MAX = 8192;

struct TopkArgs {
  const float* Q;
  const float* C;
  float* part_s;   // (nq, S, k)
  int32_t* part_i;
  int64_t nq, nc;
  int d, dpad, k, S;
  int64_t tiles_per_split, ntiles;
};

__device__ __forceinline__ bool beats(float as, int ai, float bs, int bi) { return (as > bs) || (as == bs && ai < bi); }

__global__ void __launch_bounds__(256, 1) topk_partial_kernel(const TopkArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int ldq = a.dpad + 4;
  float* Qs = reinterpret_cast<float*>(smem);                 // [QT][ldq]
  float* Cs = Qs + QT * ldq;                                  // [CT][ldq]
  float* list_s = Cs + CT * ldq;                              // [QT][KP]
  int* list_i = reinterpret_cast<int*>(list_s + QT * KP);     // [QT][KP]
  float* queue_s = reinterpret_cast<float*>(list_i + QT * KP);  // [QT][CT]
  float* thr = queue_s + QT * CT;                             // [QT]
  int* qcount = reinterpret_cast<int*>(thr + QT);             // [QT]
  unsigned char* queue_n = reinterpret_cast<unsigned char*>(qcount + QT);  // [QT][CT]

  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int tx = t & 15, ty = t >> 4;
  const int64_t q0 = (int64_t)blockIdx.x * QT;
  const int split = blockIdx.y;
  const int k = a.k;

  // stage the query tile (zero padded), init lists
  for (int idx = t; idx < QT * a.dpad; idx += 256) {
    const int m = idx / a.dpad, c = idx - m * a.dpad;
    const int64_t q = q0 + m;
    Qs[m * ldq + c] = (q < a.nq && c < a.d) ? a.Q[q * a.d + c] : 0.f;
  }
  for (int idx = t; idx < QT * KP; idx += 256) {
    list_s[idx] = -INFINITY;
    list_i[idx] = INT_MAX;
  }
  if (t < QT) {
    thr[t] = -INFINITY;
    qcount[t] = 0;
  }
  __syncthreads();

  const int64_t tile_beg = (int64_t)split * a.tiles_per_split;
  const int64_t tile_end = min(a.ntiles, tile_beg + a.tiles_per_split);
  for (int64_t tile = tile_beg; tile < tile_end; ++tile) {
    const int64_t c0 = tile * CT;
    // stage candidate tile
    if ((a.d & 3) == 0 && ((reinterpret_cast<uintptr_t>(a.C) & 15u) == 0)) {
      const int d4 = a.d >> 2;
      for (int idx = t; idx < CT * d4; idx += 256) {
        const int n = idx / d4, c4 = idx - n * d4;
        const int64_t c = c0 + n;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < a.nc) v = ldg_nc_f4(a.C + c * a.d + c4 * 4);
        *reinterpret_cast<float4*>(&Cs[n * ldq + c4 * 4]) = v;
      }
    } else {
      for (int idx = t; idx < CT * a.dpad; idx += 256) {
        const int n = idx / a.dpad, c = idx - n * a.dpad;
        const int64_t cc = c0 + n;
        Cs[n * ldq + c] = (cc < a.nc && c < a.d) ? a.C[cc * a.d + c] : 0.f;
      }
    }
    __syncthreads();

    float s[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) s[i][j] = 0.f;
    for (int kk = 0; kk < a.dpad; kk += 4) {
      float4 qv[4], cv[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) qv[i] = *reinterpret_cast<const float4*>(&Qs[(ty * 4 + i) * ldq + kk]);
#pragma unroll
      for (int j = 0; j < 8; ++j) cv[j] = *reinterpret_cast<const float4*>(&Cs[(tx + 16 * j) * ldq + kk]);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          s[i][j] = fmaf(qv[i].x, cv[j].x, s[i][j]);
          s[i][j] = fmaf(qv[i].y, cv[j].y, s[i][j]);
          s[i][j] = fmaf(qv[i].z, cv[j].z, s[i][j]);
          s[i][j] = fmaf(qv[i].w, cv[j].w, s[i][j]);
        }
    }
    // filter against the running k-th best
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = ty * 4 + i;
      const float th = thr[m];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int n = tx + 16 * j;
        if (c0 + n < a.nc && s[i][j] >= th) {
          const int pos = atomicAdd(&qcount[m], 1);
          queue_s[m * CT + pos] = s[i][j];
          queue_n[m * CT + pos] = (unsigned char)n;
        }
      }
    }
    __syncthreads();
    // merge survivors: warp w owns rows w, w+8, ...
    for (int m = warp; m < QT; m += 8) {
      const int cnt = qcount[m];
      if (cnt == 0) continue;
      float* ls = list_s + m * KP;
      int* li = list_i + m * KP;
      for (int e = 0; e < cnt; ++e) {
        const float sc = queue_s[m * CT + e];
        const int ci = (int)(c0 + queue_n[m * CT + e]);
        float cs[4];
        int cix[4];
        int pos = 0;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          cs[c] = ls[lane + 32 * c];
          cix[c] = li[lane + 32 * c];
          pos += __popc(__ballot_sync(0xffffffffu, beats(cs[c], cix[c], sc, ci)));
        }
        if (pos >= k) continue;    // warp-uniform
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int x = lane + 32 * c;
          if (x >= pos && x + 1 < KP) {
            ls[x + 1] = cs[c];
            li[x + 1] = cix[c];
          }
        }
        if (lane == 0) {
          ls[pos] = sc;
          li[pos] = ci;
        }
        __syncwarp();
      }
      if (lane == 0) {
        thr[m] = ls[k - 1];
        qcount[m] = 0;
      }
    }
    __syncthreads();
  }
  // write partial lists
  for (int idx = t; idx < QT * k; idx += 256) {
    const int m = idx / k, r = idx - m * k;
    const int64_t q = q0 + m;
    if (q < a.nq) {
      const int64_t o = (q * a.S + split) * k + r;
      a.part_s[o] = list_s[m * KP + r];
      a.part_i[o] = list_i[m * KP + r];
    }
  }
}

// One block per query: bitonic sort of the S*k partial entries (padded to P, a power of two).
__global__ void __launch_bounds__(256) topk_merge_kernel(const float* __restrict__ part_s, const int32_t* __restrict__ part_i,
                                                         const int32_t* __restrict__ cand_ids, float* __restrict__ top_s,
                                                         int32_t* __restrict__ top_i, int S, int k, int P) {
  extern __shared__ __align__(16) unsigned char smem[];
  float* ss = reinterpret_cast<float*>(smem);
  int* si = reinterpret_cast<int*>(ss + P);
  const int64_t q = blockIdx.x;
  const int n = S * k;
  for (int i = threadIdx.x; i < P; i += blockDim.x) {
    ss[i] = i < n ? part_s[q * n + i] : -INFINITY;
    si[i] = i < n ? part_i[q * n + i] : INT_MAX;
  }
  __syncthreads();
  for (int size = 2; size <= P; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = threadIdx.x; i < (P >> 1); i += blockDim.x) {
        const int lo = 2 * i - (i & (stride - 1));
        const int hi = lo + stride;
        const bool up = ((lo & size) == 0);   // "up" blocks put the better element first
        const float a_s = ss[lo], b_s = ss[hi];
        const int a_i = si[lo], b_i = si[hi];
        const bool b_first = beats(b_s, b_i, a_s, a_i);
        if (b_first == up) {
          ss[lo] = b_s; ss[hi] = a_s;
          si[lo] = b_i; si[hi] = a_i;
        }
      }
      __syncthreads();
    }
  }
  for (int r = threadIdx.x; r < k; r += blockDim.x) {
    const int idx = si[r];
    if (top_s) top_s[q * k + r] = ss[r];
    top_i[q * k + r] = (cand_ids && idx != INT_MAX) ? cand_ids[idx] : idx;
  }
}

int choose_splits(int64_t nq, int64_t nc, int k) {
  const int64_t qtiles = ceil_div<int64_t>(nq, QT);
  const int64_t ntiles = ceil_div<int64_t>(nc, CT);
  int64_t S = ceil_div<int64_t>((int64_t)sm_count() * 4, qtiles);
  S = krs::imin<int64_t>(S, MERGE_MAX / k);
  S = krs::imin<int64_t>(S, ntiles);
  S = krs::imin<int64_t>(S, 65535);
  return (int)krs::imax<int64_t>(1, S);
}

size_t partial_smem(int dpad) {
  const int ldq = dpad + 4;
  return sizeof(float) * ((size_t)QT * ldq + (size_t)CT * ldq + (size_t)QT * KP * 2 + (size_t)QT * CT + QT * 2) +
         (size_t)QT * CT;
}

}  // namespace
}  // namespace krs

using namespace krs;

extern "C" size_t krs_topk_workspace_bytes(int64_t nq, int64_t nc, int d, int k) {
  (void)d;
  if (nq <= 0 || nc <= 0 || k <= 0 || k > KP) return 0;
  const int S = choose_splits(nq, nc, k);
  return (size_t)nq * S * k * (sizeof(float) + sizeof(int32_t));
}

extern "C" int krs_topk(const float* Q, const float* C, const int32_t* cand_ids, float* top_scores, int32_t* top_ids,
                        int64_t nq, int64_t nc, int d, int k, void* workspace, size_t workspace_bytes, void* stream) {
  KRS_REQUIRE(Q && C && top_ids, "krs_topk: null argument");
  KRS_REQUIRE(nq >= 0 && nc > 0 && d > 0, "krs_topk: bad shape");
  KRS_REQUIRE(k >= 1 && k <= KP, "krs_topk: k must be in 1..%d, got %d", KP, k);
  KRS_REQUIRE(nc >= k, "krs_topk: The number of candidates provided (%lld) is less than the number of candidates to retrieve (k=%d).",
              (long long)nc, k);
  KRS_REQUIRE(d <= MAXD, "krs_topk: embedding dim %d > %d not supported", d, MAXD);
  KRS_REQUIRE(nc < (int64_t)INT_MAX, "krs_topk: too many candidates for int32 indices");
  if (nq == 0) return KRS_OK;
  cudaStream_t s = as_stream(stream);
  const int S = choose_splits(nq, nc, k);
  const size_t need = (size_t)nq * S * k * (sizeof(float) + sizeof(int32_t));
  KRS_REQUIRE(workspace && workspace_bytes >= need, "krs_topk: workspace too small (%zu < %zu)", workspace_bytes, need);
  TopkArgs a;
  a.Q = Q; a.C = C;
  a.part_s = reinterpret_cast<float*>(workspace);
  a.part_i = reinterpret_cast<int32_t*>(a.part_s + (size_t)nq * S * k);
  a.nq = nq; a.nc = nc; a.d = d; a.dpad = (d + 3) & ~3; a.k = k; a.S = S;
  a.ntiles = ceil_div<int64_t>(nc, CT);
  a.tiles_per_split = ceil_div<int64_t>(a.ntiles, S);
  const size_t smem1 = partial_smem(a.dpad);
  KRS_CUDA(cudaFuncSetAttribute(topk_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
  dim3 grid((unsigned)ceil_div<int64_t>(nq, QT), (unsigned)S);
  topk_partial_kernel<<<grid, 256, smem1, s>>>(a);
  KRS_LAUNCH_CHECK();
  int P = 1;
  while (P < S * k) P <<= 1;
  const size_t smem2 = (size_t)P * 8;
  KRS_CUDA(cudaFuncSetAttribute(topk_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
  topk_merge_kernel<<<(unsigned)nq, 256, smem2, s>>>(a.part_s, a.part_i, cand_ids, top_scores, top_ids, S, k, P);
  KRS_LAUNCH_CHECK();
  return KRS_OK;
}
