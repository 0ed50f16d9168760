// gemm_tc.cu — tcgen05 3xTF32 GEMM (placeholder until the UMMA kernel lands: reports unsupported so
// krs::gemm falls through to the exact FFMA engine).
#include "common.cuh"
namespace krs {
int gemm_tc(const float*, int64_t, bool, const float*, int64_t, bool, float*, int64_t, int64_t, int64_t, int64_t,
            const Epilogue&, int, bool, cudaStream_t) {
  return KRS_EUNSUPPORTED;
}
}  // namespace krs
