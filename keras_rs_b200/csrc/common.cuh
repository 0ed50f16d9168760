// common.cuh — shared helpers for libkrs_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/krs_b200.h"

namespace krs {

void set_error(const char* fmt, ...);
int fail_cuda(cudaError_t e, const char* what, const char* file, int line);
int sm_count();

#define KRS_CUDA(expr)                                                          \
  do {                                                                          \
    cudaError_t _e = (expr);                                                    \
    if (_e != cudaSuccess) return krs::fail_cuda(_e, #expr, __FILE__, __LINE__); \
  } while (0)

#define KRS_REQUIRE(cond, ...)          \
  do {                                  \
    if (!(cond)) {                      \
      krs::set_error(__VA_ARGS__);      \
      return KRS_EINVAL;                \
    }                                   \
  } while (0)

#define KRS_LAUNCH_CHECK() KRS_CUDA(cudaGetLastError())

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
template <typename T>
static inline T ceil_div(T a, T b) { return (a + b - 1) / b; }
template <typename T>
__host__ __device__ static inline T imax(T a, T b) { return a > b ? a : b; }
template <typename T>
__host__ __device__ static inline T imin(T a, T b) { return a < b ? a : b; }

// ---- device helpers -------------------------------------------------------
__device__ __forceinline__ float4 ldg_nc_f4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_cs_f4(float* p, float4 v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float act_apply(int act, float z) {
  switch (act) {
    case KRS_ACT_RELU: return fmaxf(z, 0.f);
    case KRS_ACT_SIGMOID: return 1.f / (1.f + expf(-z));
    case KRS_ACT_TANH: return tanhf(z);
    case KRS_ACT_SWISH: return z / (1.f + expf(-z));
    default: return z;
  }
}
// derivative given pre-activation z and a = act(z)
__device__ __forceinline__ float act_grad(int act, float z, float a) {
  switch (act) {
    case KRS_ACT_RELU: return z > 0.f ? 1.f : 0.f;
    case KRS_ACT_SIGMOID: return a * (1.f - a);
    case KRS_ACT_TANH: return 1.f - a * a;
    case KRS_ACT_SWISH: { float s = 1.f / (1.f + expf(-z)); return s * (1.f + z * (1.f - s)); }
    default: return 1.f;
  }
}
// derivative expressed through the OUTPUT y = act(z) (Dense backward keeps only y)
__device__ __forceinline__ float act_grad_from_out(int act, float y) {
  switch (act) {
    case KRS_ACT_RELU: return y > 0.f ? 1.f : 0.f;
    case KRS_ACT_SIGMOID: return y * (1.f - y);
    case KRS_ACT_TANH: return 1.f - y * y;
    default: return 1.f;
  }
}

// ---- internal GEMM interface (gemm_ffma.cu / gemm_tc.cu) -------------------
enum EpiKind { EPI_NONE = 0, EPI_BIAS_ACT = 1, EPI_CROSS = 2, EPI_ADD2 = 3 };
struct Epilogue {
  int kind = EPI_NONE;
  const float* bias = nullptr;   // [N]
  int act = KRS_ACT_LINEAR;
  // EPI_CROSS: z = acc + bias ; a = act(z) ; h2 = a + diag*x ; C = x0*h2 + x  (all ld = ldc)
  const float* x0 = nullptr;
  const float* x = nullptr;
  float diag = 0.f;
  float* h2_out = nullptr;
  float* z_out = nullptr;
  // EPI_ADD2: C = acc + alpha1*add1 + alpha2*add2
  const float* add1 = nullptr;
  const float* add2 = nullptr;
  float alpha1 = 0.f, alpha2 = 0.f;
};
// C(M,N) = op(A) @ op(B) with epilogue.  transA: A stored (K,M), lda = M-stride of K rows, etc.
// split_k > 1 requires EPI_NONE; C is zeroed by the callee and accumulated with atomics.
int gemm(const float* A, int64_t lda, bool transA, const float* B, int64_t ldb, bool transB, float* C,
         int64_t ldc, int64_t M, int64_t N, int64_t K, const Epilogue& epi, int split_k, bool accumulate,
         cudaStream_t stream);
int gemm_ffma(const float* A, int64_t lda, bool transA, const float* B, int64_t ldb, bool transB, float* C,
              int64_t ldc, int64_t M, int64_t N, int64_t K, const Epilogue& epi, int split_k, bool accumulate,
              cudaStream_t stream);
// returns KRS_EUNSUPPORTED when the shape/alignment is outside what the tcgen05 kernel handles
int gemm_tc(const float* A, int64_t lda, bool transA, const float* B, int64_t ldb, bool transB, float* C,
            int64_t ldc, int64_t M, int64_t N, int64_t K, const Epilogue& epi, int split_k, bool accumulate,
            cudaStream_t stream, int ver = 1);
int pick_split_k(int64_t M, int64_t N, int64_t K);

}  // namespace krs
