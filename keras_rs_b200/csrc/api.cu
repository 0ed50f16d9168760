// api.cu — error plumbing, version, device queries, engine switch.
#include <atomic>
#include <mutex>

#include "common.cuh"

namespace krs {
static thread_local char g_err[512] = "";
static std::atomic<int> g_engine{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int fail_cuda(cudaError_t e, const char* what, const char* file, int line) {
  set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
  return KRS_ECUDA;
}
int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}
int gemm(const float* A, int64_t lda, bool transA, const float* B, int64_t ldb, bool transB, float* C,
         int64_t ldc, int64_t M, int64_t N, int64_t K, const Epilogue& epi, int split_k, bool accumulate,
         cudaStream_t stream) {
  const int eng = g_engine.load();
  if (eng == 1 || eng == 2) {
    int rc = gemm_tc(A, lda, transA, B, ldb, transB, C, ldc, M, N, K, epi, split_k, accumulate, stream, eng);
    if (rc != KRS_EUNSUPPORTED) return rc;
  }
  return gemm_ffma(A, lda, transA, B, ldb, transB, C, ldc, M, N, K, epi, split_k, accumulate, stream);
}
}  // namespace krs

extern "C" {
int krs_version(void) { return 100; }
const char* krs_last_error(void) { return krs::g_err; }
int krs_device_sm_count(void) { return krs::sm_count(); }
int krs_set_gemm_engine(int engine) {
  if (engine < 0 || engine > 2) {
    krs::set_error("krs_set_gemm_engine: engine must be 0 (ffma), 1 (tcgen05, SS operands) or 2 (tcgen05, A in TMEM), got %d", engine);
    return KRS_EINVAL;
  }
  krs::g_engine.store(engine);
  return KRS_OK;
}
int krs_get_gemm_engine(void) { return krs::g_engine.load(); }
}
