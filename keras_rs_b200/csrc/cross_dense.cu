// cross_dense.cu — FeatureCross (DCN-v2) and Dense forward/backward on top of krs::gemm, plus the
// elementwise tails and the loss kernel.
//
// Reference behaviour restated (not ported): FeatureCross.call, feature_cross.py:155-194;
// keras.layers.Dense = act(x @ kernel + bias) (examples/dcn.py:444-447).
//
// Forward cross = ONE GEMM whose epilogue does +bias, pre_activation, +diag*x, x0*(.)+x in
// registers (the elementwise cross never round-trips HBM).  Backward = one elementwise "prep" pass
// (dz, dx0, column sums for db) + the three contractions dV = h^T dz, dh = dz V^T (+ residual adds in
// its epilogue), and for low rank dU = x^T dh, dx = dh U^T.
#include "common.cuh"

namespace krs {
namespace {

// dz = gy * mult * act'()   (mult = x0 for the cross, absent for Dense)
// aux = gy * h2             (cross: dx0)
// db[c] += sum_rows dz
// MODE 0: cross (derivative from z, recomputing a = act(z));  MODE 1: dense (derivative from y).
template <int MODE, bool VEC>
__global__ void __launch_bounds__(256) bwd_prep_kernel(const float* __restrict__ gy, const float* __restrict__ mult,
                                                       const float* __restrict__ zy, const float* __restrict__ h2,
                                                       float* __restrict__ dz, float* __restrict__ aux,
                                                       float* __restrict__ db, int64_t M, int N, int act,
                                                       int rows_per_block, int aux_mode) {
  __shared__ float red[8][128];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c0 = blockIdx.x * 128;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
  const int64_t r1 = min(M, r0 + rows_per_block);
  float s[4] = {0.f, 0.f, 0.f, 0.f};
  int col[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) col[j] = VEC ? c0 + lane * 4 + j : c0 + lane + 32 * j;
  for (int64_t r = r0 + warp; r < r1; r += 8) {
    const int64_t base = r * N;
    float g[4], mu[4], zz[4], hh[4], o[4], ax[4];
    if (VEC) {
      if (col[0] < N) {  // N % 4 == 0 in VEC mode, so the whole quad is in range
        float4 t = *reinterpret_cast<const float4*>(gy + base + col[0]);
        g[0] = t.x; g[1] = t.y; g[2] = t.z; g[3] = t.w;
        if (mult) { t = *reinterpret_cast<const float4*>(mult + base + col[0]); mu[0] = t.x; mu[1] = t.y; mu[2] = t.z; mu[3] = t.w; }
        if (zy) { t = *reinterpret_cast<const float4*>(zy + base + col[0]); zz[0] = t.x; zz[1] = t.y; zz[2] = t.z; zz[3] = t.w; }
        if (h2) { t = *reinterpret_cast<const float4*>(h2 + base + col[0]); hh[0] = t.x; hh[1] = t.y; hh[2] = t.z; hh[3] = t.w; }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (col[j] < N) {
          g[j] = gy[base + col[j]];
          if (mult) mu[j] = mult[base + col[j]];
          if (zy) zz[j] = zy[base + col[j]];
          if (h2) hh[j] = h2[base + col[j]];
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (col[j] < N) {
        float d = g[j];
        if (mult) d *= mu[j];
        if (act != KRS_ACT_LINEAR && zy) {
          if (MODE == 0) d *= act_grad(act, zz[j], act_apply(act, zz[j]));
          else d *= act_grad_from_out(act, zz[j]);
        }
        o[j] = d;
        s[j] += d;
        if (h2) ax[j] = g[j] * hh[j] + ((aux_mode & 2) ? g[j] : 0.f);
      }
    }
    if (VEC) {
      if (col[0] < N) {
        *reinterpret_cast<float4*>(dz + base + col[0]) = make_float4(o[0], o[1], o[2], o[3]);
        if (h2 && aux) {
          float4 prev = make_float4(0.f, 0.f, 0.f, 0.f);
          if (aux_mode & 1) prev = *reinterpret_cast<const float4*>(aux + base + col[0]);
          *reinterpret_cast<float4*>(aux + base + col[0]) =
              make_float4(prev.x + ax[0], prev.y + ax[1], prev.z + ax[2], prev.w + ax[3]);
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (col[j] < N) {
          dz[base + col[j]] = o[j];
          if (h2 && aux) aux[base + col[j]] = ((aux_mode & 1) ? aux[base + col[j]] : 0.f) + ax[j];
        }
    }
  }
  if (db == nullptr) return;
#pragma unroll
  for (int j = 0; j < 4; ++j) red[warp][VEC ? lane * 4 + j : lane + 32 * j] = s[j];
  __syncthreads();
  if (threadIdx.x < 128) {
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) tot += red[w][threadIdx.x];
    const int c = c0 + threadIdx.x;
    if (c < N) atomicAdd(db + c, tot);
  }
}

int launch_prep(int mode, const float* gy, const float* mult, const float* zy, const float* h2, float* dz, float* aux,
                float* db, int64_t M, int N, int act, int aux_mode, cudaStream_t s) {
  if (M == 0 || N == 0) return KRS_OK;
  if (db) KRS_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * (size_t)N, s));
  const int rows_per_block = 64;
  dim3 grid((unsigned)ceil_div(N, 128), (unsigned)ceil_div<int64_t>(M, rows_per_block));
  KRS_REQUIRE(grid.y <= 65535u, "bwd_prep: too many rows (%lld)", (long long)M);
  const bool vec = (N % 4 == 0) && aligned16(gy) && aligned16(dz) && (!mult || aligned16(mult)) &&
                   (!zy || aligned16(zy)) && (!h2 || aligned16(h2)) && (!aux || aligned16(aux));
  if (mode == 0) {
    if (vec) bwd_prep_kernel<0, true><<<grid, 256, 0, s>>>(gy, mult, zy, h2, dz, aux, db, M, N, act, rows_per_block, aux_mode);
    else bwd_prep_kernel<0, false><<<grid, 256, 0, s>>>(gy, mult, zy, h2, dz, aux, db, M, N, act, rows_per_block, aux_mode);
  } else {
    if (vec) bwd_prep_kernel<1, true><<<grid, 256, 0, s>>>(gy, mult, zy, h2, dz, aux, db, M, N, act, rows_per_block, aux_mode);
    else bwd_prep_kernel<1, false><<<grid, 256, 0, s>>>(gy, mult, zy, h2, dz, aux, db, M, N, act, rows_per_block, aux_mode);
  }
  KRS_LAUNCH_CHECK();
  return KRS_OK;
}

__global__ void cross_combine_fwd_kernel(const float* __restrict__ x0, const float* __restrict__ x,
                                         const float* __restrict__ a, float diag, float* __restrict__ y, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float h2 = diag != 0.f ? a[i] + diag * x[i] : a[i];
    y[i] = x0[i] * h2 + x[i];
  }
}
__global__ void cross_combine_bwd_kernel(const float* __restrict__ gy, const float* __restrict__ x0,
                                         const float* __restrict__ x, const float* __restrict__ a, float diag,
                                         float* __restrict__ dx0, float* __restrict__ dx, float* __restrict__ da,
                                         int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float g = gy[i];
    float h2 = diag != 0.f ? a[i] + diag * x[i] : a[i];
    float dh2 = g * x0[i];
    dx0[i] = g * h2;
    da[i] = dh2;
    dx[i] = g + diag * dh2;
  }
}

__global__ void axpy_mul_kernel(float* __restrict__ y, const float* __restrict__ a, const float* __restrict__ b,
                                    float alpha, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] += alpha * a[i] * b[i];
}


// ------------------------------------------------------------------ skinny Dense (N <= 8, e.g. the final Dense(1))
// A 128-wide GEMM tile wastes >99% of its work on a 1-column output; these are single streaming passes.
template <int NMAX>
__global__ void __launch_bounds__(256) dense_skinny_fwd_kernel(const float* __restrict__ x, const float* __restrict__ W,
                                                               const float* __restrict__ b, float* __restrict__ y, int64_t M,
                                                               int K, int N, int act) {
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t m = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; m < M; m += nwarps) {
    float acc[NMAX];
#pragma unroll
    for (int n = 0; n < NMAX; ++n) acc[n] = 0.f;
    const float* xr = x + m * K;
    for (int k = lane; k < K; k += 32) {
      const float xv = xr[k];
#pragma unroll
      for (int n = 0; n < NMAX; ++n)
        if (n < N) acc[n] = fmaf(xv, W[(int64_t)k * N + n], acc[n]);
    }
#pragma unroll
    for (int n = 0; n < NMAX; ++n)
      for (int o = 16; o > 0; o >>= 1) acc[n] += __shfl_xor_sync(0xffffffffu, acc[n], o);
    if (lane == 0) {
#pragma unroll
      for (int n = 0; n < NMAX; ++n)
        if (n < N) y[m * N + n] = act_apply(act, acc[n] + (b ? b[n] : 0.f));
    }
  }
}

// dW[k][n] += sum_m x[m][k] * dz[m][n]   (dW pre-zeroed); block = row chunk, thread = column k
template <int NMAX>
__global__ void __launch_bounds__(256) dense_skinny_dw_kernel(const float* __restrict__ x, const float* __restrict__ dz,
                                                              float* __restrict__ dW, int64_t M, int K, int N,
                                                              int rows_per_block) {
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  const int64_t r1 = min(M, r0 + rows_per_block);
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    float acc[NMAX];
#pragma unroll
    for (int n = 0; n < NMAX; ++n) acc[n] = 0.f;
    for (int64_t m = r0; m < r1; ++m) {
      const float xv = x[m * K + k];
#pragma unroll
      for (int n = 0; n < NMAX; ++n)
        if (n < N) acc[n] = fmaf(xv, dz[m * N + n], acc[n]);
    }
#pragma unroll
    for (int n = 0; n < NMAX; ++n)
      if (n < N) atomicAdd(dW + (int64_t)k * N + n, acc[n]);
  }
}

// dx[m][k] = sum_n dz[m][n] * W[k][n]
template <int NMAX>
__global__ void __launch_bounds__(256) dense_skinny_dx_kernel(const float* __restrict__ dz, const float* __restrict__ W,
                                                              float* __restrict__ dx, int64_t M, int K, int N) {
  const int64_t total = M * (int64_t)K;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = i / K;
    const int k = (int)(i - m * K);
    float acc = 0.f;
#pragma unroll
    for (int n = 0; n < NMAX; ++n)
      if (n < N) acc = fmaf(dz[m * N + n], W[(int64_t)k * N + n], acc);
    dx[i] = acc;
  }
}

// loss: kind 0 MSE, 1 BCE(prob), 2 BCE(logits).  loss pre-zeroed; mean over B.
__global__ void loss_kernel(const float* __restrict__ pred, const float* __restrict__ label, float* __restrict__ loss,
                            float* __restrict__ dpred, int64_t B, int kind, float invB) {
  __shared__ float red[32];
  float acc = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < B; i += (int64_t)gridDim.x * blockDim.x) {
    float p = pred[i], y = label[i], l, d;
    if (kind == 0) {
      float e = p - y;
      l = e * e;
      d = 2.f * e * invB;
    } else if (kind == 1) {
      const float eps = 1e-7f;
      float pc = fminf(fmaxf(p, eps), 1.f - eps);
      l = -(y * logf(pc) + (1.f - y) * logf(1.f - pc));
      bool inside = (p > eps) && (p < 1.f - eps);
      d = inside ? (-(y / pc) + (1.f - y) / (1.f - pc)) * invB : 0.f;
    } else {
      // max(z,0) - z*y + log(1+exp(-|z|))
      l = fmaxf(p, 0.f) - p * y + log1pf(expf(-fabsf(p)));
      d = (1.f / (1.f + expf(-p)) - y) * invB;
    }
    acc += l;
    if (dpred) dpred[i] = d;
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) atomicAdd(loss, v * invB);
  }
}

}  // namespace
}  // namespace krs

using namespace krs;

extern "C" {

int krs_cross_fwd(const float* x0, const float* x, const float* U, const float* V, const float* b, float diag_scale,
                  int act, float* y, float* h2_out, float* z_out, float* hproj, int64_t B, int D, int P,
                  void* stream) {
  KRS_REQUIRE(x0 && x && V && y, "krs_cross_fwd: null x0/x/V/y");
  KRS_REQUIRE(B >= 0 && D > 0, "krs_cross_fwd: bad shape B=%lld D=%d", (long long)B, D);
  KRS_REQUIRE(act >= KRS_ACT_LINEAR && act <= KRS_ACT_SWISH, "krs_cross_fwd: unknown activation %d", act);
  KRS_REQUIRE(!(diag_scale < 0.f), "krs_cross_fwd: `diag_scale` should be non-negative");
  cudaStream_t s = as_stream(stream);
  if (B == 0) return KRS_OK;
  Epilogue e;
  e.kind = EPI_CROSS;
  e.bias = b;
  e.act = act;
  e.x0 = x0;
  e.x = x;
  e.diag = diag_scale;
  e.h2_out = h2_out;
  e.z_out = z_out;
  if (U == nullptr) return gemm(x, D, false, V, D, false, y, D, B, D, D, e, 1, false, s);
  KRS_REQUIRE(P > 0 && hproj, "krs_cross_fwd: low-rank needs P > 0 and the hproj buffer");
  Epilogue none;
  int rc = gemm(x, D, false, U, P, false, hproj, P, B, P, D, none, 1, false, s);   // h = x @ U
  if (rc) return rc;
  return gemm(hproj, P, false, V, D, false, y, D, B, D, P, e, 1, false, s);         // y = cross(h @ V)
}

int krs_cross_bwd(const float* gy, const float* x0, const float* x, const float* U, const float* V, const float* h2,
                  const float* z, const float* hproj, float diag_scale, int act, float* dx0, float* dx, float* dU,
                  float* dV, float* db, float* dz, float* dh, int64_t B, int D, int P, int flags, void* stream) {
  KRS_REQUIRE(gy && x0 && x && V && h2 && dx0 && dx && dV && dz, "krs_cross_bwd: null argument");
  KRS_REQUIRE(act == KRS_ACT_LINEAR || z != nullptr, "krs_cross_bwd: non-linear pre_activation needs z");
  cudaStream_t s = as_stream(stream);
  if (B == 0) {
    const int Kh = U ? P : D;
    KRS_CUDA(cudaMemsetAsync(dV, 0, sizeof(float) * (size_t)Kh * D, s));
    if (db) KRS_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * D, s));
    if (dU) KRS_CUDA(cudaMemsetAsync(dU, 0, sizeof(float) * (size_t)D * P, s));
    return KRS_OK;
  }
  // 1. dz = gy * x0 * act'(z) ; dx0 = gy * h2 ; db = colsum(dz)
  const int aux_mode = ((flags & KRS_CROSS_SAME_INPUT) ? 2 : 0) | ((flags & KRS_CROSS_ACC_DX0) ? 1 : 0);
  int rc = launch_prep(0, gy, x0, act == KRS_ACT_LINEAR ? nullptr : z, h2, dz, dx0, db, B, D, act, aux_mode, s);
  if (rc) return rc;
  Epilogue none;
  Epilogue add2;
  add2.kind = EPI_ADD2;
  add2.add1 = (flags & KRS_CROSS_SAME_INPUT) ? dx0 : gy;   // same input: dx0 holds gy*h2 + gy
  add2.alpha1 = 1.f;
  // note: the diag term of dx uses dh2 = gy*x0 which equals dz only for the linear activation;
  // for non-linear activations dh2 is recomputed by a small combine below.
  if (U == nullptr) {
    // dV (D,D) = x^T @ dz
    rc = gemm(x, D, true, dz, D, false, dV, D, D, D, B, none, pick_split_k(D, D, B), false, s);
    if (rc) return rc;
    if (act == KRS_ACT_LINEAR) { add2.add2 = dz; add2.alpha2 = diag_scale; }
    // dx = dz @ V^T + gy (+ diag * dh2)
    rc = gemm(dz, D, false, V, D, true, dx, D, B, D, D, add2, 1, false, s);
  } else {
    KRS_REQUIRE(P > 0 && hproj && dh && dU, "krs_cross_bwd: low-rank needs P, hproj, dh, dU");
    rc = gemm(hproj, P, true, dz, D, false, dV, D, P, D, B, none, pick_split_k(P, D, B), false, s);  // dV = h^T dz
    if (rc) return rc;
    rc = gemm(dz, D, false, V, D, true, dh, P, B, P, D, none, 1, false, s);                          // dh = dz V^T
    if (rc) return rc;
    rc = gemm(x, D, true, dh, P, false, dU, P, D, P, B, none, pick_split_k(D, P, B), false, s);     // dU = x^T dh
    if (rc) return rc;
    if (act == KRS_ACT_LINEAR) { add2.add2 = dz; add2.alpha2 = diag_scale; }
    rc = gemm(dh, P, false, U, P, true, dx, D, B, D, P, add2, 1, false, s);                          // dx = dh U^T + ...
  }
  if (rc) return rc;
  if (act != KRS_ACT_LINEAR && diag_scale != 0.f) {
    // dx += diag * gy * x0  (dh2), elementwise
    const int64_t n = B * (int64_t)D;
    axpy_mul_kernel<<<(unsigned)krs::imin<int64_t>(ceil_div<int64_t>(n, 256), 148 * 16), 256, 0, s>>>(dx, gy, x0, diag_scale, n);
    KRS_LAUNCH_CHECK();
  }
  return KRS_OK;
}

int krs_cross_combine_fwd(const float* x0, const float* x, const float* a, float diag_scale, float* y, int64_t n,
                          void* stream) {
  KRS_REQUIRE(x0 && x && a && y, "krs_cross_combine_fwd: null argument");
  if (n == 0) return KRS_OK;
  cross_combine_fwd_kernel<<<(unsigned)krs::imin<int64_t>(ceil_div<int64_t>(n, 256), 148 * 16), 256, 0, as_stream(stream)>>>(
      x0, x, a, diag_scale, y, n);
  KRS_LAUNCH_CHECK();
  return KRS_OK;
}

int krs_cross_combine_bwd(const float* gy, const float* x0, const float* x, const float* a, float diag_scale,
                          float* dx0, float* dx, float* da, int64_t n, void* stream) {
  KRS_REQUIRE(gy && x0 && x && a && dx0 && dx && da, "krs_cross_combine_bwd: null argument");
  if (n == 0) return KRS_OK;
  cross_combine_bwd_kernel<<<(unsigned)krs::imin<int64_t>(ceil_div<int64_t>(n, 256), 148 * 16), 256, 0, as_stream(stream)>>>(
      gy, x0, x, a, diag_scale, dx0, dx, da, n);
  KRS_LAUNCH_CHECK();
  return KRS_OK;
}

int krs_dense_fwd(const float* x, const float* W, const float* b, int act, float* y, int64_t B, int K, int N,
                  void* stream) {
  KRS_REQUIRE(x && W && y, "krs_dense_fwd: null argument");
  KRS_REQUIRE(act >= KRS_ACT_LINEAR && act <= KRS_ACT_SWISH, "krs_dense_fwd: unknown activation %d", act);
  if (N <= 8 && B > 0) {
    const unsigned grid = (unsigned)krs::imax<int64_t>(1, krs::imin<int64_t>(ceil_div<int64_t>(B, 8), (int64_t)sm_count() * 16));
    dense_skinny_fwd_kernel<8><<<grid, 256, 0, as_stream(stream)>>>(x, W, b, y, B, K, N, act);
    KRS_LAUNCH_CHECK();
    return KRS_OK;
  }
  Epilogue e;
  e.kind = EPI_BIAS_ACT;
  e.bias = b;
  e.act = act;
  return gemm(x, K, false, W, N, false, y, N, B, N, K, e, 1, false, as_stream(stream));
}

int krs_dense_bwd(const float* gy, const float* x, const float* W, const float* y, int act, float* dx, float* dW,
                  float* db, float* dz, int64_t B, int K, int N, void* stream) {
  KRS_REQUIRE(gy && x && W && dW, "krs_dense_bwd: null argument");
  KRS_REQUIRE(act != KRS_ACT_SWISH, "krs_dense_bwd: swish needs the pre-activation; use linear/relu/sigmoid/tanh");
  KRS_REQUIRE(act == KRS_ACT_LINEAR || (y && dz), "krs_dense_bwd: activation needs y and the dz workspace");
  cudaStream_t s = as_stream(stream);
  if (B == 0) {
    KRS_CUDA(cudaMemsetAsync(dW, 0, sizeof(float) * (size_t)K * N, s));
    if (db) KRS_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * N, s));
    return KRS_OK;
  }
  const float* dzp = gy;
  if (act != KRS_ACT_LINEAR || db) {
    if (act == KRS_ACT_LINEAR && dz == nullptr) {
      // only the bias column sums are needed: run prep with dz aliasing nothing is not possible, so
      // require the workspace whenever db is requested.
      KRS_REQUIRE(dz, "krs_dense_bwd: db needs the dz workspace");
    }
    int rc = launch_prep(1, gy, nullptr, act == KRS_ACT_LINEAR ? nullptr : y, nullptr, dz, nullptr, db, B, N, act, 0, s);
    if (rc) return rc;
    dzp = dz;
  }
  if (N <= 8) {
    KRS_CUDA(cudaMemsetAsync(dW, 0, sizeof(float) * (size_t)K * N, s));
    const int rows_per_block = 128;
    dense_skinny_dw_kernel<8><<<(unsigned)ceil_div<int64_t>(B, rows_per_block), 256, 0, s>>>(x, dzp, dW, B, K, N, rows_per_block);
    KRS_LAUNCH_CHECK();
    if (dx) {
      const int64_t total = B * (int64_t)K;
      dense_skinny_dx_kernel<8><<<(unsigned)krs::imax<int64_t>(1, krs::imin<int64_t>(ceil_div<int64_t>(total, 256), (int64_t)sm_count() * 32)),
                                  256, 0, s>>>(dzp, W, dx, B, K, N);
      KRS_LAUNCH_CHECK();
    }
    return KRS_OK;
  }
  Epilogue none;
  int rc = gemm(x, K, true, dzp, N, false, dW, N, K, N, B, none, pick_split_k(K, N, B), false, s);   // dW = x^T dz
  if (rc) return rc;
  if (dx) rc = gemm(dzp, N, false, W, N, true, dx, K, B, K, N, none, 1, false, s);                   // dx = dz W^T
  return rc;
}

int krs_loss_fwd_bwd(const float* pred, const float* label, float* loss, float* dpred, int64_t B, int kind,
                     int64_t denom, void* stream) {
  KRS_REQUIRE(pred && label && loss, "krs_loss_fwd_bwd: null argument");
  KRS_REQUIRE(kind >= 0 && kind <= 2, "krs_loss_fwd_bwd: unknown loss kind %d", kind);
  KRS_REQUIRE(B > 0, "krs_loss_fwd_bwd: empty batch");
  cudaStream_t s = as_stream(stream);
  KRS_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), s));
  loss_kernel<<<(unsigned)krs::imin<int64_t>(ceil_div<int64_t>(B, 256), 148 * 4), 256, 0, s>>>(pred, label, loss, dpred, B, kind,
                                                                                      1.f / (float)(denom > 0 ? denom : B));
  KRS_LAUNCH_CHECK();
  return KRS_OK;
}

}  // extern "C"

