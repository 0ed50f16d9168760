// umma_probe.cu — issue-rate microbenchmark of tcgen05.mma (kind::tf32) on sm_100a.  One warp per CTA streams MMAs
// back to back (each under elect.sync); all SMs run the same stream.  Reports clocks per MMA by operand source
// (SS: A and B by shared-memory descriptor; TS: A from tensor memory), M, N, B layout / swizzle, commit frequency
// and unroll.  Data are finite floats; results are not checked (layout is irrelevant to timing).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o benchmarks/bin/umma_probe benchmarks/umma_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__host__ __device__ inline uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(layout & 7) << 61;
  return d;
}
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}\n" : "=r"(pred));
  return pred;
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t ta, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(ta), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

struct Cfg {
  int ts, M, N, reps;
  uint32_t a_lbo, a_sbo, a_layout, b_lbo, b_sbo, b_layout;
};

// PC: a tcgen05.commit after every PC MMAs (0 = only at the end); UNR: MMAs per loop iteration (straight line)
template <int PC, int UNR>
__global__ void __launch_bounds__(128, 1) probe(const Cfg c, unsigned long long* out) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* A = reinterpret_cast<float*>(smem);               // 32 KB
  float* B = reinterpret_cast<float*>(smem + 32768);       // 64 KB
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  for (int i = threadIdx.x; i < (32768 + 65536) / 4; i += blockDim.x) A[i] = 1.0f + 0.001f * (float)(i % 97);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = __shfl_sync(0xffffffffu, tmem_ptr, 0);
  {
    const uint32_t ta = tbase + ((uint32_t)((threadIdx.x >> 5) * 32) << 16) + 496;
    uint32_t z = 0x3f800000u;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(ta), "r"(z) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int warp_u = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // provably warp-uniform (CUTLASS canonical_warp_idx_sync)
  if (warp_u == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(c.b_layout == 1 ? 1u : 0u) << 16) |
                           ((uint32_t)(c.N >> 3) << 17) | ((uint32_t)(c.M >> 4) << 24);
    const uint64_t da = make_desc(smem_u32(A), c.a_lbo, c.a_sbo, c.a_layout);
    const uint64_t db = make_desc(smem_u32(B), c.b_lbo, c.b_sbo, c.b_layout);
    const uint32_t ta = tbase + 496;
    uint32_t phase = 0;
    int since = 0;
    const unsigned long long t0 = clock64();
    for (int r = 0; r < c.reps; r += UNR) {
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const uint32_t acc = (r + u) >= 1 ? 1u : 0u;
        if (elect_one()) {
          if (c.ts) mma_ts(tbase, ta, db, idesc, acc);
          else mma_ss(tbase, da, db, idesc, acc);
        }
        if (PC > 0 && ++since == PC) {
          since = 0;
          if (elect_one()) commit(&bar);
          phase ^= 1;
        }
      }
    }
    const unsigned long long t_issue = clock64();
    if (elect_one()) commit(&bar);
    uint32_t ok = 0;
    long long spins = 0;
    while (!ok && ++spins < (1ll << 26)) {
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(phase) : "memory");
    }
    const unsigned long long t1 = clock64();
    if (threadIdx.x == 0) {
      out[2 * blockIdx.x] = t1 - t0;
      out[2 * blockIdx.x + 1] = t_issue - t0;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "n"(512));
  }
}

template <int PC, int UNR>
static void run(const char* what, Cfg c, int grid, unsigned long long* out) {
  static bool attr = false;
  cudaFuncSetAttribute(probe<PC, UNR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  (void)attr;
  for (int w = 0; w < 2; ++w) {
    probe<PC, UNR><<<grid, 128, 98 * 1024>>>(c, out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s (%s)\n", cudaGetErrorString(e), what); exit(1); }
  }
  double tot = 0, iss = 0;
  for (int b = 0; b < grid; ++b) { tot += (double)out[2 * b]; iss += (double)out[2 * b + 1]; }
  printf("%-44s %s M=%3d N=%3d commit/%d unroll %d : %7.1f clk/MMA (issue %7.1f)  N/2 = %d\n", what, c.ts ? "TS" : "SS", c.M, c.N, PC, UNR,
         tot / grid / c.reps, iss / grid / c.reps, c.N / 2);
}

int main(int argc, char** argv) {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int grid = argc > 1 ? atoi(argv[1]) : sms;
  unsigned long long* out;
  cudaMallocManaged(&out, sizeof(unsigned long long) * 2 * 1024);
  const int reps = 4096;
  // descriptor presets: K-major SWIZZLE_64B rows of 64 B (layout 4, SBO 512); K-major SWIZZLE_128B rows of 128 B
  // (layout 2, SBO 1024); MN-major SWIZZLE_128B_BASE32B (layout 1, LBO 2048, SBO 512)
  auto cfg = [&](int ts, int M, int N, int b_kind) {
    Cfg c{ts, M, N, reps, 16, 512, 4, 16, 512, 4};
    if (b_kind == 1) { c.b_lbo = 16; c.b_sbo = 1024; c.b_layout = 2; c.a_sbo = 1024; c.a_layout = 2; }
    if (b_kind == 2) { c.b_lbo = 2048; c.b_sbo = 512; c.b_layout = 1; }
    return c;
  };
  printf("grid %d CTAs, %d tf32 MMAs each (K = 8 per MMA)\n", grid, reps);
  for (int ts = 0; ts < 2; ++ts)
    for (int N : {64, 128, 192, 256}) run<0, 1>("K-major SW64", cfg(ts, 128, N, 0), grid, out);
  for (int ts = 0; ts < 2; ++ts)
    for (int N : {128, 256}) run<0, 4>("K-major SW64, 4 MMAs straight-line", cfg(ts, 128, N, 0), grid, out);
  for (int N : {128, 256}) run<0, 1>("K-major SW128 (A and B)", cfg(0, 128, N, 1), grid, out);
  for (int ts = 0; ts < 2; ++ts)
    for (int N : {128, 256}) run<0, 1>("B MN-major SW128_BASE32B", cfg(ts, 128, N, 2), grid, out);
  for (int ts = 0; ts < 2; ++ts)
    for (int N : {64, 128, 256}) run<0, 1>("M = 64", cfg(ts, 64, N, 0), grid, out);
  run<1, 1>("commit cost", cfg(0, 128, 128, 0), grid, out);
  run<2, 1>("commit cost", cfg(0, 128, 128, 0), grid, out);
  run<4, 1>("commit cost", cfg(0, 128, 128, 0), grid, out);
  run<8, 1>("commit cost", cfg(0, 128, 128, 0), grid, out);
  run<4, 1>("commit cost", cfg(0, 128, 256, 0), grid, out);
  run<4, 1>("commit cost", cfg(1, 128, 256, 0), grid, out);
  run<8, 1>("commit cost", cfg(1, 128, 256, 0), grid, out);
  return 0;
}
