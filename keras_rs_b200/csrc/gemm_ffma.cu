// gemm_ffma.cu — exact-fp32 SGEMM (FFMA pipe) with the fused epilogues of the hot path.
//
// This is the "engine 0" contraction used by FeatureCross / Dense / their backward passes: every
// product is a true fp32 FMA, so results sit within accumulation-order noise of the reference's
// fp32 matmul (keras.ops.matmul under the JAX-CPU backend; SURVEY.md §7 "hard parts").  The tensor
// pipe engine (gemm_tc.cu, tcgen05 3xTF32) shares the Epilogue contract and is validated against
// this kernel.
//
// Tile: 128x128x16 per CTA, 256 threads, 8x8 accumulators per thread (two 4x4 quadrant pairs so
// shared-memory reads are conflict-free LDS.128 and global stores are 16-byte vectors), register
// prefetch + double-buffered shared memory.  All four operand layouts (NN/NT/TN/TT) are handled by
// the loaders; ragged M/N/K and unaligned pointers fall back to guarded scalar loads.
#include "common.cuh"

namespace krs {
namespace {

constexpr int BM = 128, BN = 128, BK = 16, PAD = 4, NT = 256;

struct GemmArgs {
  const float* A;
  const float* B;
  float* C;
  int64_t lda, ldb, ldc;
  int64_t M, N, K;
  int64_t k_chunk;  // K range per blockIdx.z (multiple of BK)
  int atomic_out;   // split-K: atomicAdd into C
  int accumulate;   // C += result (non-atomic path)
  Epilogue epi;
};

// Loads a (128 mn) x (16 k) operand tile into registers.  KCONTIG: memory is [mn][k] (k fastest);
// otherwise [k][mn] (mn fastest).  Each thread carries 2 float4.
template <bool KCONTIG, bool VEC>
__device__ __forceinline__ void load_tile(const float* __restrict__ P, int64_t ld, int64_t mn0, int64_t MN,
                                          int64_t k0, int64_t Kend, int t, float4 (&r)[2]) {
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    int64_t mn, k;
    if (KCONTIG) {
      mn = mn0 + (t >> 2) + 64 * i;
      k = k0 + (t & 3) * 4;
    } else {
      k = k0 + (t >> 5) + 8 * i;
      mn = mn0 + (t & 31) * 4;
    }
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (KCONTIG) {
      if (mn < MN) {
        const float* p = P + mn * ld + k;
        if (VEC && k + 3 < Kend) {
          v = *reinterpret_cast<const float4*>(p);
        } else {
          if (k + 0 < Kend) v.x = p[0];
          if (k + 1 < Kend) v.y = p[1];
          if (k + 2 < Kend) v.z = p[2];
          if (k + 3 < Kend) v.w = p[3];
        }
      }
    } else {
      if (k < Kend) {
        const float* p = P + k * ld + mn;
        if (VEC && mn + 3 < MN) {
          v = *reinterpret_cast<const float4*>(p);
        } else {
          if (mn + 0 < MN) v.x = p[0];
          if (mn + 1 < MN) v.y = p[1];
          if (mn + 2 < MN) v.z = p[2];
          if (mn + 3 < MN) v.w = p[3];
        }
      }
    }
    r[i] = v;
  }
}

template <bool KCONTIG>
__device__ __forceinline__ void store_tile(float (*S)[BM + PAD], int t, const float4 (&r)[2]) {
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    if (KCONTIG) {
      int mn = (t >> 2) + 64 * i;
      int k = (t & 3) * 4;
      S[k + 0][mn] = r[i].x;
      S[k + 1][mn] = r[i].y;
      S[k + 2][mn] = r[i].z;
      S[k + 3][mn] = r[i].w;
    } else {
      int k = (t >> 5) + 8 * i;
      int mn = (t & 31) * 4;
      *reinterpret_cast<float4*>(&S[k][mn]) = r[i];
    }
  }
}

__device__ __forceinline__ void epilogue_store(const GemmArgs& g, int64_t m, int64_t n, float4 acc, bool vec_ok) {
  const Epilogue& e = g.epi;
  float v[4] = {acc.x, acc.y, acc.z, acc.w};
  const int64_t off = m * g.ldc + n;
  const bool full = vec_ok && (n + 3 < g.N);
  const int cnt = full ? 4 : (int)min((int64_t)4, g.N - n);
  if (g.atomic_out) {
    for (int j = 0; j < cnt; ++j) atomicAdd(g.C + off + j, v[j]);
    return;
  }
  float out[4], h2v[4], zv[4];
  if (e.kind == EPI_NONE) {
    for (int j = 0; j < cnt; ++j) out[j] = v[j] + (g.accumulate ? g.C[off + j] : 0.f);
  } else if (e.kind == EPI_BIAS_ACT) {
    for (int j = 0; j < cnt; ++j) {
      float z = v[j] + (e.bias ? e.bias[n + j] : 0.f);
      out[j] = act_apply(e.act, z);
    }
  } else if (e.kind == EPI_CROSS) {
    float x0v[4], xv[4];
    if (full) {
      float4 a = *reinterpret_cast<const float4*>(e.x0 + off);
      float4 b = *reinterpret_cast<const float4*>(e.x + off);
      x0v[0] = a.x; x0v[1] = a.y; x0v[2] = a.z; x0v[3] = a.w;
      xv[0] = b.x; xv[1] = b.y; xv[2] = b.z; xv[3] = b.w;
    } else {
      for (int j = 0; j < cnt; ++j) { x0v[j] = e.x0[off + j]; xv[j] = e.x[off + j]; }
    }
    for (int j = 0; j < cnt; ++j) {
      float z = v[j] + (e.bias ? e.bias[n + j] : 0.f);
      float a = act_apply(e.act, z);
      float h2 = (e.diag != 0.f) ? a + e.diag * xv[j] : a;   // feature_cross.py:191-192
      zv[j] = z;
      h2v[j] = h2;
      out[j] = x0v[j] * h2 + xv[j];                           // feature_cross.py:194
    }
    if (e.h2_out) {
      if (full) *reinterpret_cast<float4*>(e.h2_out + off) = make_float4(h2v[0], h2v[1], h2v[2], h2v[3]);
      else for (int j = 0; j < cnt; ++j) e.h2_out[off + j] = h2v[j];
    }
    if (e.z_out) {
      if (full) *reinterpret_cast<float4*>(e.z_out + off) = make_float4(zv[0], zv[1], zv[2], zv[3]);
      else for (int j = 0; j < cnt; ++j) e.z_out[off + j] = zv[j];
    }
  } else {  // EPI_ADD2
    for (int j = 0; j < cnt; ++j) {
      float r = v[j];
      if (e.add1) r += e.alpha1 * e.add1[off + j];
      if (e.add2) r += e.alpha2 * e.add2[off + j];
      out[j] = r;
    }
  }
  if (full) *reinterpret_cast<float4*>(g.C + off) = make_float4(out[0], out[1], out[2], out[3]);
  else for (int j = 0; j < cnt; ++j) g.C[off + j] = out[j];
}

template <bool TA, bool TB, bool VEC_A, bool VEC_B>
__global__ void __launch_bounds__(NT) sgemm_kernel(const GemmArgs g) {
  __shared__ __align__(16) float As[2][BK][BM + PAD];
  __shared__ __align__(16) float Bs[2][BK][BN + PAD];
  const int t = threadIdx.x;
  const int tx = t & 15, ty = t >> 4;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int64_t n0 = (int64_t)blockIdx.y * BN;
  const int64_t kbeg = (int64_t)blockIdx.z * g.k_chunk;
  const int64_t kend = min(g.K, kbeg + g.k_chunk);

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float4 ra[2], rb[2];
  // A is logically (M,K): not transposed => memory [m][k] => KCONTIG.  B logically (K,N): not
  // transposed => memory [k][n] => MN-contig; transposed => memory [n][k] => KCONTIG.
  load_tile<!TA, VEC_A>(g.A, g.lda, m0, g.M, kbeg, kend, t, ra);
  load_tile<TB, VEC_B>(g.B, g.ldb, n0, g.N, kbeg, kend, t, rb);
  store_tile<!TA>(As[0], t, ra);
  store_tile<TB>(Bs[0], t, rb);
  __syncthreads();

  int buf = 0;
  for (int64_t k0 = kbeg; k0 < kend; k0 += BK) {
    const bool has_next = (k0 + BK) < kend;
    if (has_next) {
      load_tile<!TA, VEC_A>(g.A, g.lda, m0, g.M, k0 + BK, kend, t, ra);
      load_tile<TB, VEC_B>(g.B, g.ldb, n0, g.N, k0 + BK, kend, t, rb);
    }
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
      float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (has_next) {
      store_tile<!TA>(As[buf ^ 1], t, ra);
      store_tile<TB>(Bs[buf ^ 1], t, rb);
      __syncthreads();
      buf ^= 1;
    }
  }

  const bool vec_ok = ((g.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(g.C) & 15u) == 0);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= g.M) continue;
#pragma unroll
    for (int jg = 0; jg < 2; ++jg) {
      const int64_t n = n0 + jg * 64 + tx * 4;
      if (n >= g.N) continue;
      epilogue_store(g, m, n, make_float4(acc[i][jg * 4 + 0], acc[i][jg * 4 + 1], acc[i][jg * 4 + 2], acc[i][jg * 4 + 3]),
                     vec_ok);
    }
  }
}

template <bool TA, bool TB>
int launch(const GemmArgs& g, dim3 grid, bool va, bool vb, cudaStream_t s) {
  if (va && vb) sgemm_kernel<TA, TB, true, true><<<grid, NT, 0, s>>>(g);
  else if (va) sgemm_kernel<TA, TB, true, false><<<grid, NT, 0, s>>>(g);
  else if (vb) sgemm_kernel<TA, TB, false, true><<<grid, NT, 0, s>>>(g);
  else sgemm_kernel<TA, TB, false, false><<<grid, NT, 0, s>>>(g);
  KRS_LAUNCH_CHECK();
  return KRS_OK;
}

}  // namespace

int pick_split_k(int64_t M, int64_t N, int64_t K) {
  const int64_t tiles = ceil_div<int64_t>(M, BM) * ceil_div<int64_t>(N, BN);
  const int sms = sm_count();
  if (tiles >= sms || K < 1024) return 1;
  int64_t want = ceil_div<int64_t>(2 * (int64_t)sms, tiles);
  int64_t max_by_k = K / 256;
  if (max_by_k < 1) max_by_k = 1;
  return (int)krs::imax<int64_t>(1, min(want, max_by_k));
}

int gemm_ffma(const float* A, int64_t lda, bool transA, const float* B, int64_t ldb, bool transB, float* C,
              int64_t ldc, int64_t M, int64_t N, int64_t K, const Epilogue& epi, int split_k, bool accumulate,
              cudaStream_t stream) {
  KRS_REQUIRE(M >= 0 && N >= 0 && K >= 0, "gemm: negative dimension");
  if (M == 0 || N == 0) return KRS_OK;
  KRS_REQUIRE(A && B && C, "gemm: null operand");
  if (split_k < 1) split_k = 1;
  KRS_REQUIRE(split_k == 1 || epi.kind == EPI_NONE, "gemm: split-K only with the plain epilogue");
  GemmArgs g;
  g.A = A; g.B = B; g.C = C;
  g.lda = lda; g.ldb = ldb; g.ldc = ldc;
  g.M = M; g.N = N; g.K = K;
  int64_t chunk = ceil_div<int64_t>(ceil_div<int64_t>(K, split_k), BK) * BK;
  if (chunk < BK) chunk = BK;
  split_k = (int)krs::imax<int64_t>(1, ceil_div<int64_t>(K, chunk));
  g.k_chunk = chunk;
  g.atomic_out = split_k > 1 ? 1 : 0;
  g.accumulate = accumulate ? 1 : 0;
  g.epi = epi;
  if (g.atomic_out && !accumulate)
    KRS_CUDA(cudaMemset2DAsync(C, (size_t)ldc * sizeof(float), 0, (size_t)N * sizeof(float), (size_t)M, stream));
  dim3 grid((unsigned)ceil_div<int64_t>(M, BM), (unsigned)ceil_div<int64_t>(N, BN), (unsigned)split_k);
  KRS_REQUIRE(grid.y <= 65535u, "gemm: N too large for grid.y (%lld cols)", (long long)N);
  const bool va = aligned16(A) && (lda % 4 == 0);
  const bool vb = aligned16(B) && (ldb % 4 == 0);
  if (!transA && !transB) return launch<false, false>(g, grid, va, vb, stream);
  if (!transA && transB) return launch<false, true>(g, grid, va, vb, stream);
  if (transA && !transB) return launch<true, false>(g, grid, va, vb, stream);
  return launch<true, true>(g, grid, va, vb, stream);
}

}  // namespace krs

extern "C" int krs_sgemm(const float* A, const float* Bm, float* C, int64_t M, int64_t N, int64_t K, int transA,
                         int transB, int accumulate, void* stream) {
  using namespace krs;
  Epilogue e;
  const int64_t lda = transA ? M : K, ldb = transB ? K : N;
  return gemm(A, lda, transA != 0, Bm, ldb, transB != 0, C, N, M, N, K, e, accumulate ? 1 : pick_split_k(M, N, K),
              accumulate != 0, as_stream(stream));
}
