// gemm_tc.cu — tcgen05 (5th-gen tensor core) GEMM with fp32-level accuracy via a 3-term TF32 split.
//
// Engines 1 and 2 of krs::gemm: the dense contractions of FeatureCross / Dense and their backward passes
// (W·x_i, dz·V^T, x^T·dz) on the tensor pipe, with the same fused epilogues as gemm_ffma.cu
// (+bias, pre_activation, +diag·x, x0 ⊙ (·) + x kept in registers).
//
// Accuracy: inputs are fp32.  Each operand is split in-kernel into hi = tf32(x) and
// lo = tf32(x - hi); the product is accumulated as lo·hi + hi·lo + hi·hi in fp32 TMEM accumulators
// (the dropped lo·lo term is ~2^-22 relative), which keeps results within ~1e-6 of an fp32 matmul —
// inside the 1e-5 bar of the north star, which a single-pass TF32 product (1e-3) would miss.
//
// Structure (one CTA per SM, persistent over output tiles, 512 threads = 4 warpgroups re-budgeted with setmaxnreg):
//   warps 0-3    epilogue: tcgen05.ld accumulator rows -> registers -> shared-memory transpose -> fused epilogue -> global,
//                instantiated per (kind, activation class) so the per-float4 code is straight-line
//   warps 4-11   converters.  VER 1 (engine "tcgen05"): shared -> registers -> shared, write the A_lo / B_lo tiles.
//                VER 2 (engine "tcgen05_ts", the default): two groups of 4 warps take alternate k-blocks; a thread owns one
//                tile row, splits its 16 raw floats in registers and parks hi / lo in a 4-deep TENSOR-MEMORY ring with
//                tcgen05.st, so the A operand never crosses the shared-memory port again (TS-form MMAs)
//   warp 12      TMA producer: cp.async.bulk.tensor fp32 tiles global -> shared (mbarrier complete_tx); also streams the
//                precomputed B_lo plane of weight operands (split_lo_kernel + krs_gemm_set_workspace)
//   warp 13      MMA issuer: tcgen05.mma.kind::tf32 from ONE elect.sync lane of a provably uniform warp (the warp index
//                comes from a shuffle broadcast), descriptors in uniform registers; per k-step  A_hi x [B_hi | B_lo] as one
//                N = 2*bn instruction ([main | cross] accumulators are adjacent) and  A_lo x B_hi  into the cross accumulator;
//                tcgen05.commit releases ring stages / A stages / publishes the accumulators
//   warps 14-15  idle (they complete the warpgroup)
// Shared memory ring (depth chosen per launch, 10 at bn = 96): VER 2  A_raw | B_raw | B_lo ; VER 1  A_raw | B_raw | B_lo | A_lo,
// BK = 16 floats per stage.
// TMEM: 2 accumulator stages x (main | cross) x bn columns (bn <= 128 in VER 1, <= 96 in VER 2 whose columns 384..511 hold the
// A ring), 128 lanes = the 128 rows of the tile, so the epilogue of tile i overlaps the mainloop of tile i+1.  The hi*hi
// products and the small cross terms (lo*hi + hi*lo) accumulate separately and are added in fp32 registers by the
// epilogue.  Reason (measured, tests/test_gpu_tc.py): the tensor core's fp32 accumulate truncates, so the error grows with
// the number of accumulations into one TMEM tile (~0.5 * 6e-8 per step, systematic).  Keeping the small terms out of the
// main accumulator cuts the step count 3x; long reductions (the batch-dimension weight gradients) are additionally split
// so that no accumulator sees more than KC_MAX/8 steps, partial tiles being combined with fp32 RED (round-to-nearest) in
// global memory.  What the measurements changed, step by step: DESIGN.md section 4.1.
//
// Operand layouts: K-contiguous operands use 64-byte rows with SWIZZLE_64B (K-major UMMA
// descriptors), MN-contiguous operands (x^T, dz in the weight-gradient GEMM, V in the forward) use
// 32-element x 16-row boxes with SWIZZLE_128B_BASE32B (MN-major descriptors; unswizzled for the A operand of VER 2, which only
// converter threads read).  Ragged edges are handled by TMA out-of-bounds zero fill on loads and guards on stores.
#include <cuda.h>
#include <cudaTypedefs.h>

#include <stdlib.h>

#include <atomic>
#include <mutex>

#include "common.cuh"
#include "tc_common.cuh"

namespace krs {
namespace {
using namespace tcx;

constexpr int BM = 128;          // rows per tile == TMEM lanes
constexpr int BK = 16;           // fp32 elements per k-block (64 B)
constexpr int MAX_STAGES = 12;    // ring depth is chosen per launch from the shared-memory budget
constexpr int MAX_BN = 128;       // 2 accumulator stages x (main + cross-term) x 128 columns = 512 TMEM columns
constexpr int NUM_THREADS = 512;          // WG0: 4 epilogue warps | WG1-2: 8 converter warps | WG3: TMA, MMA, 2 idle
constexpr int W_TMA = 12, W_MMA = 13;
constexpr int NUM_CONV_THREADS = 256;
constexpr int EPI_LD = 36;                // staging row pitch in floats (32 + 4: conflict-free 16-byte accesses)
constexpr int EPI_SMEM_BYTES = 4 * 32 * EPI_LD * 4;   // one 32x32 staging tile per epilogue warp
constexpr int A_BYTES = BM * BK * 4;             // 8192
constexpr int KC_MAX = 1024;                     // max K elements accumulated inside one TMEM tile
// Kernel versions.  VER 1: both operands read by the tensor core from shared memory (SS), converters write
// A_lo / B_lo tiles to shared memory.  VER 2: the A operand (hi and lo) lives in TENSOR MEMORY (TS): converter
// threads own one tile row each, read it from the TMA-landed raw tile, split it in registers and tcgen05.st the
// two halves into a 4-deep TMEM ring, so A never crosses the shared-memory port again (VER 1 spends 8 KB of
// writes + 3 x 8 KB of UMMA operand reads per k-block on it; the shared-memory pipe was the measured limiter,
// profiles/r1_gemm_tc_clock_trace.txt).  TMEM: 2 x (main + cross) x 96 accumulator columns + 4 x 32 A columns.
template <int VER> struct Cfg {
  static constexpr int ACC_BN = VER == 2 ? 96 : 128;   // max N tile == accumulator width
  static constexpr int A_COL0 = 4 * ACC_BN;            // VER 2: first TMEM column of the A ring
  static constexpr int KSUB = VER == 2 ? 2 : 1;        // 16-wide k sub-blocks per ring stage: VER 2 moves 32 k per barrier
                                                       // round trip (the converter chain is latency-bound)
  static constexpr int SA = 4 / KSUB;                  // VER 2: A ring depth (KSUB x 32 columns each: hi | lo per sub-block)
};

std::atomic<long long> g_tc_launches{0};
std::atomic<long long> g_split_launches{0};
std::atomic<unsigned long long*> g_trace{nullptr};

struct TcArgs {
  float* C;
  int64_t ldc;
  int64_t M, N, K;
  int bn;                 // N tile (multiple of 16, <= 256)
  int stages;             // shared-memory ring depth (3..MAX_STAGES)
  int tiles_m, tiles_n, splits;
  int64_t kblocks_total;  // ceil(K / BK)
  int64_t kblocks_per_split;
  int a_mn_major, b_mn_major;
  int a_3d, b_3d;         // MN-major operand loaded by ONE 3-D box {32, BK, blocks} per k-block
  int atomic_out, accumulate;
  unsigned long long* trace;   // debug: (tag, clock) pairs from CTA 0 (nullptr normally)
  int epi_vec;            // rows of C / h2 / z / aux streams are 16-byte aligned (vector epilogue)
  int no_mask;            // 1: leave hi = raw fp32 bits (hardware ignores the low 13 mantissa bits)
  int mn_lbo, mn_sbo, mn_kstep, mn_layout;   // MN-major descriptor strides (bytes) and UMMA layout type
  int fuse_n;             // one N = 2*bn MMA computes A_hi x [B_hi | B_lo] (main and hi*lo terms together)
  uint32_t wait_ns;       // suspend-time hint of the critical-path mbarrier waits
  int b_lo_tma;           // B_lo comes precomputed from global memory by TMA (tmap_blo); converters skip the B tile
  Epilogue epi;
};

// debug timeline (CTA 0 only): four role-private regions of 2000 (tag, clock64) pairs written with plain
// stores — a returning atomic would stall the traced thread for ~700 clocks per event and distort the timeline
// Compiled out unless the library is built with -DKRS_TC_TRACE=1 (KRS_EXTRA_FLAGS=-DKRS_TC_TRACE=1 bash build.sh): even
// with a null trace pointer the guards (LDC of the pointer, predicate chains, predicated-off address math) were ~17 % of
// the converter warps' stall samples (ncu source view of the first tcgen05_ts capture).
#ifndef KRS_TC_TRACE
#define KRS_TC_TRACE 0
#endif
__device__ __forceinline__ void trace_ev(const unsigned long long* tr_c, int role, int& count, unsigned tag, unsigned idx) {
  if (!KRS_TC_TRACE) return;
  unsigned long long* tr = const_cast<unsigned long long*>(tr_c);
  if (tr == nullptr || blockIdx.x != 0 || count >= 2000) return;
  tr[1 + role * 4000 + 2 * count] = ((unsigned long long)tag << 32) | idx;
  tr[2 + role * 4000 + 2 * count] = (unsigned long long)clock64();
  ++count;
}

// MN-major tile (32-bit elements): blocks of 32 mn x 16 k (2048 B).  For tf32 the ONLY MN-major layout
// the tensor core accepts is SWIZZLE_128B_BASE32B (layout type 1; cutlass sm100_common.inl:92): 32-byte
// chunks swizzled within a 128-byte row over groups of 4 k-rows — TMA mode SWIZZLE_128B_ATOM_32B.
// 4-k groups are 512 B apart (SBO), mn blocks 2048 B apart (LBO); a k-step (8 rows) = +1024 B.
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t tile_addr, int kstep, const TcArgs& g) {
  return make_desc(tile_addr + kstep * g.mn_kstep, g.mn_lbo, g.mn_sbo, (uint32_t)g.mn_layout);
}


// ---------------------------------------------------------------- fused epilogue on one float4 (4 columns of one row)
// Called AFTER the accumulator chunk has been transposed through shared memory, so that a warp instruction
// touches 4 rows x 128 contiguous bytes (4 cache lines) instead of 32 rows x 16 bytes (32 lines): the first
// version's row-per-thread epilogue spent 5.6K clocks per 16-column chunk in the LSU (tests/tc_trace.py).
__device__ __forceinline__ const float* epi_stream1(const Epilogue& e) {
  return e.kind == EPI_CROSS ? e.x0 : (e.kind == EPI_ADD2 ? e.add1 : nullptr);
}
__device__ __forceinline__ const float* epi_stream2(const Epilogue& e) {
  return e.kind == EPI_CROSS ? e.x : (e.kind == EPI_ADD2 ? e.add2 : nullptr);
}
// Compile-time epilogue kinds: the kernel is instantiated per (kind, activation class) so that the per-float4 code
// is straight-line.  The first version dispatched on g.epi.kind / act / atomic_out / accumulate at run time inside the
// innermost loop: ~10 LDCU -> UISETP -> BRA.U chains plus a jump table per ELEMENT for the activation cost ~1250 clocks
// per 4-row group and made the epilogue as long as the mainloop (tests/tc_trace.py tags 13/14).
enum TcKind { TK_STORE = 0, TK_ACCUM = 1, TK_ATOMIC = 2, TK_BIAS_ACT = 3, TK_CROSS = 4, TK_ADD2 = 5, TK_COUNT = 6 };
enum TcAct { TA_LINEAR = 0, TA_RELU = 1, TA_GENERIC = 2, TA_COUNT = 3 };
template <int ACT>
__device__ __forceinline__ float tc_act(int act, float z) {
  if (ACT == TA_LINEAR) return z;
  if (ACT == TA_RELU) return fmaxf(z, 0.f);
  return act_apply(act, z);
}
// Run-time epilogue operands copied to registers once per kernel (warp-uniform values).
struct EpiRegs {
  float* C;
  float* h2_out;
  float* z_out;
  int64_t ldc;
  float diag, alpha1, alpha2;
  int act;
  bool has1, has2;
};
// vec path: n..n+3 in range, all pointers 16-byte aligned; o = m * ldc + n
template <int KIND, int ACT>
__device__ __forceinline__ void epilogue_f4(const EpiRegs& e, int64_t o, float4 acc, float4 s1, float4 s2, float4 bv) {
  if (KIND == TK_ATOMIC) {
    atomicAdd(reinterpret_cast<float4*>(e.C + o), acc);          // RED.E.ADD.F32x4, round-to-nearest in L2
    return;
  }
  const float v[4] = {acc.x, acc.y, acc.z, acc.w};
  const float a1[4] = {s1.x, s1.y, s1.z, s1.w};
  const float a2[4] = {s2.x, s2.y, s2.z, s2.w};
  const float b[4] = {bv.x, bv.y, bv.z, bv.w};
  float out[4];
  if (KIND == TK_STORE) {
#pragma unroll
    for (int j = 0; j < 4; ++j) out[j] = v[j];
  } else if (KIND == TK_ACCUM) {
    const float4 c = *reinterpret_cast<const float4*>(e.C + o);
    out[0] = v[0] + c.x; out[1] = v[1] + c.y; out[2] = v[2] + c.z; out[3] = v[3] + c.w;
  } else if (KIND == TK_BIAS_ACT) {
#pragma unroll
    for (int j = 0; j < 4; ++j) out[j] = tc_act<ACT>(e.act, v[j] + b[j]);
  } else if (KIND == TK_CROSS) {
    float h2v[4], zv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float z = v[j] + b[j];
      const float a = tc_act<ACT>(e.act, z);
      const float h2 = (e.diag != 0.f) ? a + e.diag * a2[j] : a;     // feature_cross.py:191-192
      zv[j] = z;
      h2v[j] = h2;
      out[j] = a1[j] * h2 + a2[j];                                   // x0 * h2 + x   (feature_cross.py:194)
    }
    if (e.h2_out) *reinterpret_cast<float4*>(e.h2_out + o) = make_float4(h2v[0], h2v[1], h2v[2], h2v[3]);
    if (e.z_out) *reinterpret_cast<float4*>(e.z_out + o) = make_float4(zv[0], zv[1], zv[2], zv[3]);
  } else {  // TK_ADD2
#pragma unroll
    for (int j = 0; j < 4; ++j) out[j] = v[j] + (e.has1 ? e.alpha1 * a1[j] : 0.f) + (e.has2 ? e.alpha2 * a2[j] : 0.f);
  }
  *reinterpret_cast<float4*>(e.C + o) = make_float4(out[0], out[1], out[2], out[3]);
}
// scalar path for ragged / unaligned outputs (p1 / p2: the two operand streams, bias nullable)
template <int KIND, int ACT>
__device__ __forceinline__ void epilogue_scalar(const EpiRegs& e, const float* p1, const float* p2, const float* bias,
                                                int64_t o, int64_t n, float acc) {
  if (KIND == TK_ATOMIC) { atomicAdd(e.C + o, acc); return; }
  const float s1 = p1 ? p1[o] : 0.f, s2 = p2 ? p2[o] : 0.f, b = bias ? bias[n] : 0.f;
  float out;
  if (KIND == TK_STORE) out = acc;
  else if (KIND == TK_ACCUM) out = acc + e.C[o];
  else if (KIND == TK_BIAS_ACT) out = tc_act<ACT>(e.act, acc + b);
  else if (KIND == TK_CROSS) {
    const float z = acc + b;
    const float a = tc_act<ACT>(e.act, z);
    const float h2 = (e.diag != 0.f) ? a + e.diag * s2 : a;
    if (e.h2_out) e.h2_out[o] = h2;
    if (e.z_out) e.z_out[o] = z;
    out = s1 * h2 + s2;
  } else out = acc + (p1 ? e.alpha1 * s1 : 0.f) + (p2 ? e.alpha2 * s2 : 0.f);
  e.C[o] = out;
}

// ---------------------------------------------------------------- the kernel
template <int VER, int KIND, int ACT>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const __grid_constant__ CUtensorMap tmap_blo, const TcArgs g) {
  constexpr int ACC_BN = Cfg<VER>::ACC_BN;
  constexpr int SA = Cfg<VER>::SA;
  constexpr int KSUB = Cfg<VER>::KSUB;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // carve: stages first (1024-aligned), then barriers
  // align by OFFSET (not by integer round trip) so the compiler keeps the shared address space (LDS/STS,
  // not generic LD/ST — ncu on the first version showed generic accesses in the converter)
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int b_bytes = g.bn * BK * 4;
  const int raw_bytes = A_BYTES + b_bytes;
  const int sub_bytes = VER == 2 ? raw_bytes + b_bytes : 2 * raw_bytes;     // one 16-k sub-block: A_raw | B_raw | B_lo (| A_lo in VER 1)
  const int stage_bytes = KSUB * sub_bytes;
  const int blo_off = raw_bytes;               // B_lo directly after B_raw: [B_hi | B_lo] is one 2*bn-row operand (fuse_n)
  const int alo_off = raw_bytes + b_bytes;     // VER 1 only
  const int cross_off = g.bn;                  // cross-term accumulator = main + bn columns
  const int STAGES = g.stages;
  float* epi_stage = reinterpret_cast<float*>(smem + (size_t)STAGES * stage_bytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * stage_bytes + EPI_SMEM_BYTES);
  uint64_t* full_bar = bars;                     // [STAGES] TMA landed
  uint64_t* conv_bar = bars + MAX_STAGES;        // [STAGES] hi/lo tiles ready
  uint64_t* empty_bar = bars + 2 * MAX_STAGES;   // [STAGES] MMAs that read the stage have completed
  uint64_t* tmem_full = bars + 3 * MAX_STAGES;   // [1]
  uint64_t* tmem_empty = bars + 3 * MAX_STAGES + 2;  // [1]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 3 * MAX_STAGES + 4);
  uint64_t* afree_bar = bars + 3 * MAX_STAGES + 5;   // [SA] VER 2: MMAs that read the TMEM A stage have completed

  // warp index through a shuffle broadcast: ptxas then KNOWS it is warp-uniform, so the role branches are uniform
  // and descriptors / addresses of the MMA warp live in uniform registers.  With a plain threadIdx.x >> 5 the issue
  // loop was compiled as a divergent region: every UTCHMMA / UTCBAR got R2UR moves and an ELECT + BRA.U.ANY waterfall
  // loop and cost ~120-160 clocks to issue, a commit ~200 (benchmarks/umma_probe.cu: 79-107 and ~7 when uniform).
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  int tcount = 0;   // debug trace cursor of this thread's role

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&conv_bar[s], VER == 2 ? 4 : NUM_CONV_THREADS / 32);   // one arrival per converter warp (VER 2: per group)
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 4);    // one arrival per epilogue warp
    }
    if (VER == 2)
      for (int a = 0; a < SA; ++a) mbar_init(&afree_bar[a], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == W_MMA) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);
  const int64_t tiles_mn = (int64_t)g.tiles_m * g.tiles_n;
  const int64_t total_tiles = tiles_mn * g.splits;

  // register budget per role (warpgroup granular; 512 threads start at 128 registers each): the epilogue needs
  // room for a chunk of operand prefetches, converters / TMA / MMA need little.  Each role's branch begins with
  // its own setmaxnreg so that ptxas allocates that branch against the adjusted budget.
  if (warp >= 12) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(VER == 2 ? 72 : 56));
  if (warp == W_TMA) {
    // ======================= TMA producer =======================
    // every lane waits for the free stage; lane 0 arms the transaction count, then the boxes of the stage
    // (1 per K-major operand, bn/32 or 4 per MN-major operand) are issued by different lanes in parallel
    int stage = 0;
    uint32_t phase = 0;
    const int nA = (g.a_mn_major && !g.a_3d) ? BM / 32 : 1;
    const int nB = (g.b_mn_major && !g.b_3d) ? g.bn / 32 : 1;
    for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int split = (int)(tile / tiles_mn);
      const int64_t rem = tile - (int64_t)split * tiles_mn;
      const int tm = (int)(rem / g.tiles_n);
      const int tn = (int)(rem - (int64_t)tm * g.tiles_n);
      const int64_t kb0 = (int64_t)split * g.kblocks_per_split;
      const int64_t kb1 = imin<int64_t>(g.kblocks_total, kb0 + g.kblocks_per_split);
      for (int64_t kb = kb0; kb < kb1; ++kb) {
        // the whole (converged) warp polls, ONE elected lane arms the transaction count and issues every box of
        // the stage with warp-uniform coordinates (UTMALDG takes uniform registers: per-lane boxes were compiled
        // into R2UR + ELECT / BRA.U.ANY waterfall loops)
        mbar_wait_uniform(&empty_bar[stage], phase ^ 1, g.wait_ns);
        if (elect_one()) {
          mbar_expect_tx(&full_bar[stage], (uint32_t)(KSUB * (raw_bytes + (g.b_lo_tma ? b_bytes : 0))));
#pragma unroll
          for (int sub = 0; sub < KSUB; ++sub) {
          unsigned char* st = smem + (size_t)stage * stage_bytes + (size_t)sub * sub_bytes;
          const int k0 = (int)(kb * (BK * KSUB)) + sub * BK;
          if (!g.a_mn_major) tma_load_2d(st, &tmap_a, k0, tm * BM, &full_bar[stage]);                       // box {16 k, 128 m}
          else if (g.a_3d) tma_load_3d(st, &tmap_a, 0, k0, tm * (BM / 32), &full_bar[stage]);              // box {32 m, 16 k, 4}
          else
            for (int i = 0; i < nA; ++i) tma_load_2d(st + i * 2048, &tmap_a, tm * BM + i * 32, k0, &full_bar[stage]);   // box {32 m, 16 k}
          unsigned char* sb = st + A_BYTES;
          if (!g.b_mn_major) tma_load_2d(sb, &tmap_b, k0, tn * g.bn, &full_bar[stage]);                     // box {16 k, bn n}
          else if (g.b_3d) tma_load_3d(sb, &tmap_b, 0, k0, tn * (g.bn / 32), &full_bar[stage]);            // box {32 n, 16 k, bn/32}
          else
            for (int i = 0; i < nB; ++i) tma_load_2d(sb + i * 2048, &tmap_b, tn * g.bn + i * 32, k0, &full_bar[stage]);
          if (g.b_lo_tma) {       // same boxes from the precomputed lo plane, landing right behind B_raw
            unsigned char* sl = st + blo_off;
            if (!g.b_mn_major) tma_load_2d(sl, &tmap_blo, k0, tn * g.bn, &full_bar[stage]);
            else if (g.b_3d) tma_load_3d(sl, &tmap_blo, 0, k0, tn * (g.bn / 32), &full_bar[stage]);
            else
              for (int i = 0; i < nB; ++i) tma_load_2d(sl + i * 2048, &tmap_blo, tn * g.bn + i * 32, k0, &full_bar[stage]);
          }
          }
          trace_ev(g.trace, 0, tcount, 1, (unsigned)kb);
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == W_MMA) {
    // ======================= MMA issuer =======================
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)g.a_mn_major << 15) |
                           ((uint32_t)g.b_mn_major << 16) | ((uint32_t)(g.bn >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    const uint32_t s0 = smem_u32(smem);
    const uint64_t dA_hi = g.a_mn_major ? desc_mnmajor(s0, 0, g) : desc_kmajor(s0, 0);
    const uint64_t dA_lo = g.a_mn_major ? desc_mnmajor(s0 + alo_off, 0, g) : desc_kmajor(s0 + alo_off, 0);
    const uint64_t dB_hi = g.b_mn_major ? desc_mnmajor(s0 + A_BYTES, 0, g) : desc_kmajor(s0 + A_BYTES, 0);
    const uint64_t dB_lo = g.b_mn_major ? desc_mnmajor(s0 + blo_off, 0, g) : desc_kmajor(s0 + blo_off, 0);
    const uint32_t idesc_ts = idesc & ~(1u << 15);                                           // TMEM A is K-major
    const uint32_t idesc_ts2 = (idesc_ts & ~(0x3Fu << 17)) | ((uint32_t)((2 * g.bn) >> 3) << 17);   // N = 2*bn
    const uint32_t idesc2 = (idesc & ~(0x3Fu << 17)) | ((uint32_t)((2 * g.bn) >> 3) << 17);
    int astage = 0;
    const uint32_t a_kstep16 = (uint32_t)(g.a_mn_major ? g.mn_kstep : 32) >> 4;
    const uint32_t b_kstep16 = (uint32_t)(g.b_mn_major ? g.mn_kstep : 32) >> 4;
    for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int split = (int)(tile / tiles_mn);
      const int64_t kb0 = (int64_t)split * g.kblocks_per_split;
      const int64_t kb1 = imin<int64_t>(g.kblocks_total, kb0 + g.kblocks_per_split);
      mbar_wait_relaxed(&tmem_empty[acc], acc_phase ^ 1);
      __syncwarp();
      tc_fence_after();
      const uint32_t d_main = tmem_base + (uint32_t)(acc * 2 * ACC_BN);
      const uint32_t d_small = d_main + (uint32_t)cross_off;
      for (int64_t kb = kb0; kb < kb1; ++kb) {
        if (KRS_TC_TRACE && g.trace != nullptr && elect_one()) trace_ev(g.trace, 1, tcount, 15, (unsigned)kb);
        mbar_wait_uniform(&conv_bar[stage], phase, g.wait_ns);
        if (KRS_TC_TRACE && g.trace != nullptr && elect_one()) trace_ev(g.trace, 1, tcount, 16, (unsigned)kb);
        tc_fence_after();
        if (elect_one()) {
          trace_ev(g.trace, 1, tcount, 4, (unsigned)kb);
          // descriptors = per-launch base (stage 0, k-step 0) + (stage offset + k-step offset) >> 4 in the
          // 14-bit start-address field: one 32-bit add each instead of rebuilding the bit fields
          const uint32_t so = (uint32_t)(stage * stage_bytes) >> 4;
          if constexpr (VER == 2) {
#pragma unroll
            for (int sk = 0; sk < KSUB * (BK / 8); ++sk) {
              const int sub = sk / (BK / 8), ks = sk % (BK / 8);
              const uint32_t ta_hi = tmem_base + (uint32_t)(Cfg<VER>::A_COL0 + astage * (32 * KSUB) + sub * 32);   // lanes 0..127
              const uint32_t ta_lo = ta_hi + 16;
              const uint32_t sso = so + ((uint32_t)(sub * sub_bytes) >> 4);
              const uint64_t dbh = dB_hi + sso + (ks ? b_kstep16 : 0u);
              const uint64_t dbl = dB_lo + sso + (ks ? b_kstep16 : 0u);
              const uint32_t first = (kb == kb0 && sk == 0) ? 0u : 1u;
              if (g.fuse_n) {
                // B_lo follows B_hi in the stage: one N = 2*bn MMA yields [hi*hi | hi*lo] in [d_main, d_main + 2*bn)
                tc_mma_tf32_ts(d_main, ta_hi + 8 * ks, dbh, idesc_ts2, first);
                tc_mma_tf32_ts(d_small, ta_lo + 8 * ks, dbh, idesc_ts, 1u);
              } else {
                tc_mma_tf32_ts(d_small, ta_lo + 8 * ks, dbh, idesc_ts, first);
                tc_mma_tf32_ts(d_small, ta_hi + 8 * ks, dbl, idesc_ts, 1u);
                tc_mma_tf32_ts(d_main, ta_hi + 8 * ks, dbh, idesc_ts, first);
              }
            }
            tc_commit(&afree_bar[astage]);                    // frees the TMEM A stage
          } else {
#pragma unroll
            for (int ks = 0; ks < BK / 8; ++ks) {
              const uint64_t dah = dA_hi + so + (ks ? a_kstep16 : 0u);
              const uint64_t dal = dA_lo + so + (ks ? a_kstep16 : 0u);
              const uint64_t dbh = dB_hi + so + (ks ? b_kstep16 : 0u);
              const uint64_t dbl = dB_lo + so + (ks ? b_kstep16 : 0u);
              const uint32_t first = (kb == kb0 && ks == 0) ? 0u : 1u;
              if (g.fuse_n) {
                tc_mma_tf32(d_main, dah, dbh, idesc2, first);    // [hi*hi | hi*lo] -> [main | cross], N = 2*bn
                tc_mma_tf32(d_small, dal, dbh, idesc, 1u);       // lo*hi -> cross
              } else {
                tc_mma_tf32(d_small, dal, dbh, idesc, first);    // cross terms: their own accumulator
                tc_mma_tf32(d_small, dah, dbl, idesc, 1u);
                tc_mma_tf32(d_main, dah, dbh, idesc, first);     // main term
              }
            }
          }
          tc_commit(&empty_bar[stage]);                       // frees the stage when these MMAs retire
          if (kb == kb1 - 1) tc_commit(&tmem_full[acc]);      // accumulators complete
          trace_ev(g.trace, 1, tcount, 5, (unsigned)kb);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
        if (++astage == SA) astage = 0;
      }
      if (kb1 <= kb0 && elect_one()) tc_commit(&tmem_full[acc]);  // empty K range: nothing accumulated
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= 4 && warp < 12) {
    // ======================= converters (warps 4..11) =======================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(VER == 2 ? 104 : 80));
    const int ct = threadIdx.x - 128;       // 0..255
    int stage = 0;
    uint32_t phase = 0;
    const int nvec = (g.b_lo_tma ? A_BYTES : raw_bytes) / 16;
    if constexpr (VER == 2) {
      // Two groups of 4 warps take alternate k-blocks.  Thread = one tile row (its warp's TMEM lane quadrant):
      // 16 raw fp32 of the row -> hi (masked bits) and lo = tf32_rn(x - hi) -> two tcgen05.st.x16 into the TMEM
      // A ring.  The group's 128 threads also write the B_lo tile (flat float4 pass, layout-agnostic).
      const int grp = (warp - 4) >> 2;
      const int q = warp & 3;
      const int row = q * 32 + lane;
      const int gt = ct - grp * 128;          // 0..127 inside the group
      const int nvb = g.b_lo_tma ? 0 : b_bytes / 16;
      constexpr int MAXVB = ACC_BN * BK * 4 / 16 / 128;   // 3
      int astage = 0;
      uint32_t aphase = 0;
      uint32_t cnt = 0;
      const uint32_t a_row_off = (uint32_t)row * 64u;
      const uint32_t a_sw = (uint32_t)(row >> 1) & 3u;     // SWIZZLE_64B: 16-byte chunk index ^= address bits [7:8]
      for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int split = (int)(tile / tiles_mn);
        const int64_t kb0 = (int64_t)split * g.kblocks_per_split;
        const int64_t kb1 = imin<int64_t>(g.kblocks_total, kb0 + g.kblocks_per_split);
        for (int64_t kb = kb0; kb < kb1; ++kb) {
          if ((int)(cnt & 1u) == grp) {
            mbar_wait(&full_bar[stage], phase, g.wait_ns);
            if (ct == 0) trace_ev(g.trace, 2, tcount, 2, (unsigned)kb);
            const unsigned char* st0 = smem + (size_t)stage * stage_bytes;
            // B tiles first (only when the lo plane is not streamed by TMA): short-lived registers
            if (nvb > 0) {
#pragma unroll
              for (int sub = 0; sub < KSUB; ++sub) {
                const float4* braw = reinterpret_cast<const float4*>(st0 + (size_t)sub * sub_bytes + A_BYTES);
                float4* blo = reinterpret_cast<float4*>(smem + (size_t)stage * stage_bytes + (size_t)sub * sub_bytes + blo_off);
                float4 bx[MAXVB];
#pragma unroll
                for (int j = 0; j < MAXVB; ++j) {
                  const int i = gt + j * 128;
                  if (i < nvb) bx[j] = braw[i];
                }
#pragma unroll
                for (int j = 0; j < MAXVB; ++j) {
                  const int i = gt + j * 128;
                  if (i < nvb) {
                    const float4 x = bx[j];
                    float4 l;
                    l.x = tf32_lo_of(x.x);
                    l.y = tf32_lo_of(x.y);
                    l.z = tf32_lo_of(x.z);
                    l.w = tf32_lo_of(x.w);
                    blo[i] = l;
                  }
                }
              }
            }
            // A rows of every sub-block: all loads, then the splits, then ONE wait for the TMEM stage
            uint32_t hi[KSUB][16], lo[KSUB][16];
#pragma unroll
            for (int sub = 0; sub < KSUB; ++sub) {
              const unsigned char* st = st0 + (size_t)sub * sub_bytes;
              if (!g.a_mn_major) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                  const uint4 v = *reinterpret_cast<const uint4*>(st + a_row_off + ((((uint32_t)c) ^ a_sw) << 4));
                  hi[sub][4 * c + 0] = v.x; hi[sub][4 * c + 1] = v.y; hi[sub][4 * c + 2] = v.z; hi[sub][4 * c + 3] = v.w;
                }
              } else {
                // MN-major raw tile, unswizzled: 4 blocks of [16 k rows x 32 m]; lanes read consecutive words
#pragma unroll
                for (int k = 0; k < 16; ++k)
                  hi[sub][k] = *reinterpret_cast<const uint32_t*>(st + q * 2048 + k * 128 + lane * 4);
              }
            }
#pragma unroll
            for (int sub = 0; sub < KSUB; ++sub)
#pragma unroll
              for (int k = 0; k < 16; ++k) {
                const uint32_t raw = hi[sub][k];
                const uint32_t h = raw & 0xFFFFE000u;
                lo[sub][k] = __float_as_uint(__uint_as_float(raw) - __uint_as_float(h)) + 0x1000u;   // = tf32_lo_of(raw)
                hi[sub][k] = h;
              }
            mbar_wait(&afree_bar[astage], aphase ^ 1, g.wait_ns);     // MMAs of the stage that used this A stage are done
            tc_fence_after();
            if (ct == 0) trace_ev(g.trace, 2, tcount, 9, (unsigned)kb);
            const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(Cfg<VER>::A_COL0 + astage * (32 * KSUB));
#pragma unroll
            for (int sub = 0; sub < KSUB; ++sub) {
              tc_st16(ta + (uint32_t)(sub * 32), hi[sub]);
              tc_st16(ta + (uint32_t)(sub * 32 + 16), lo[sub]);
            }
            if (ct == 0) trace_ev(g.trace, 2, tcount, 8, (unsigned)kb);
            tc_wait_st();                                                     // TMEM stores complete
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // B_lo visible to the async proxy
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&conv_bar[stage]);
            if (ct == 0) trace_ev(g.trace, 2, tcount, 3, (unsigned)kb);
          }
          ++cnt;
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
          if (++astage == SA) { astage = 0; aphase ^= 1; }
        }
      }
    } else
    for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int split = (int)(tile / tiles_mn);
      const int64_t kb0 = (int64_t)split * g.kblocks_per_split;
      const int64_t kb1 = imin<int64_t>(g.kblocks_total, kb0 + g.kblocks_per_split);
      for (int64_t kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase, g.wait_ns);
        if (ct == 0) trace_ev(g.trace, 2, tcount, 2, (unsigned)kb);
        float4* raw = reinterpret_cast<float4*>(smem + (size_t)stage * stage_bytes);
        float4* lo = reinterpret_cast<float4*>(smem + (size_t)stage * stage_bytes + raw_bytes);
        constexpr int A_VECS = A_BYTES / 16;
        const int b_vecs = b_bytes / 16;
        // all loads first (<= 6 float4 per thread at bn = 256), then split + store: one shared-memory
        // latency per k-block instead of one per element
        constexpr int MAXV = (A_BYTES + MAX_BN * BK * 4) / 16 / NUM_CONV_THREADS;   // 6
        float4 xv[MAXV];
#pragma unroll
        for (int j = 0; j < MAXV; ++j) {
          const int i = ct + j * NUM_CONV_THREADS;
          if (i < nvec) xv[j] = raw[i];
        }
#pragma unroll
        for (int j = 0; j < MAXV; ++j) {
          const int i = ct + j * NUM_CONV_THREADS;
          if (i < nvec) {
            const float4 x = xv[j];
            float4 h, l;
            h.x = __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u);
            h.y = __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u);
            h.z = __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u);
            h.w = __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u);
            l.x = __uint_as_float(__float_as_uint(x.x - h.x) + 0x1000u);
            l.y = __uint_as_float(__float_as_uint(x.y - h.y) + 0x1000u);
            l.z = __uint_as_float(__float_as_uint(x.z - h.z) + 0x1000u);
            l.w = __uint_as_float(__float_as_uint(x.w - h.w) + 0x1000u);
            if (!g.no_mask) raw[i] = h;
            lo[i < A_VECS ? i + b_vecs : i - A_VECS] = l;      // lo region = B_lo | A_lo
          }
        }
        if (ct == 0) trace_ev(g.trace, 2, tcount, 8, (unsigned)kb);
        // generic-proxy writes must be visible to the tensor core (async proxy) before the MMA reads them
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (ct == 0) trace_ev(g.trace, 2, tcount, 9, (unsigned)kb);
        __syncwarp();
        if (lane == 0) mbar_arrive(&conv_bar[stage]);
        if (ct == 0) trace_ev(g.trace, 2, tcount, 3, (unsigned)kb);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp < 4) {
    // ======================= epilogue (warps 0..3) =======================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(224));
    const int quad = warp & 3;              // TMEM lane quadrant this warp may read
    float* T = epi_stage + warp * (32 * EPI_LD);           // this warp's 32 x 32 staging tile
    const int rr = lane >> 3, c4 = lane & 7;               // coalesced domain: 4 rows x 8 float4 per instruction
    const float* p1 = (KIND == TK_CROSS || KIND == TK_ADD2) ? epi_stream1(g.epi) : nullptr;
    const float* p2 = (KIND == TK_CROSS || KIND == TK_ADD2) ? epi_stream2(g.epi) : nullptr;
    const float* bias = (KIND == TK_CROSS || KIND == TK_BIAS_ACT) ? g.epi.bias : nullptr;
    EpiRegs er;
    er.C = g.C; er.h2_out = g.epi.h2_out; er.z_out = g.epi.z_out; er.ldc = g.ldc;
    er.diag = g.epi.diag; er.alpha1 = g.epi.alpha1; er.alpha2 = g.epi.alpha2; er.act = g.epi.act;
    er.has1 = p1 != nullptr; er.has2 = p2 != nullptr;
    // L2 prefetch of the epilogue operand tiles (x0 / x, add1 / add2) of a FUTURE tile: issued one tile ahead so the
    // epilogue's loads hit L2 instead of paying DRAM latency with only 4 warps' worth of requests in flight
    auto prefetch_tile_l2 = [&](int64_t t) {
      if (t >= total_tiles || (p1 == nullptr && p2 == nullptr)) return;
      const int sp = (int)(t / tiles_mn);
      const int64_t rm = t - (int64_t)sp * tiles_mn;
      const int ptm = (int)(rm / g.tiles_n);
      const int ptn = (int)(rm - (int64_t)ptm * g.tiles_n);
      const int64_t m = (int64_t)ptm * BM + quad * 32 + lane;
      if (m >= g.M) return;
      const int64_t n0 = (int64_t)ptn * g.bn;
      const int64_t ncols = imin<int64_t>(g.bn, g.N - n0);
      for (int64_t c = 0; c < ncols; c += 32) {
        if (p1) asm volatile("prefetch.global.L2 [%0];" ::"l"(p1 + m * g.ldc + n0 + c));
        if (p2) asm volatile("prefetch.global.L2 [%0];" ::"l"(p2 + m * g.ldc + n0 + c));
      }
    };
    prefetch_tile_l2(blockIdx.x);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int split = (int)(tile / tiles_mn);
      const int64_t rem = tile - (int64_t)split * tiles_mn;
      const int tm = (int)(rem / g.tiles_n);
      const int tn = (int)(rem - (int64_t)tm * g.tiles_n);
      const int64_t m_base = (int64_t)tm * BM + quad * 32;
      const int64_t nbase = (int64_t)tn * g.bn;
      prefetch_tile_l2(tile + gridDim.x);
      // operand streams of chunk c for this lane's 8 (row, float4) slots; requested one chunk ahead
      float4 a1[8], a2[8];
      auto prefetch = [&](int c, float4 (&u1)[8], float4 (&u2)[8]) {
        const int64_t n = nbase + c + c4 * 4;
        const bool col_ok = g.epi_vec && (c + c4 * 4 < g.bn) && (n + 3 < g.N);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int64_t m = m_base + rr + 4 * i;
          const bool ok = col_ok && m < g.M;
          u1[i] = (ok && p1) ? __ldg(reinterpret_cast<const float4*>(p1 + m * g.ldc + n)) : make_float4(0.f, 0.f, 0.f, 0.f);
          u2[i] = (ok && p2) ? __ldg(reinterpret_cast<const float4*>(p2 + m * g.ldc + n)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      };
      mbar_wait_relaxed(&tmem_full[acc], acc_phase);
      tc_fence_after();
      if (warp == 0 && lane == 0) trace_ev(g.trace, 3, tcount, 6, (unsigned)tile);
      const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * 2 * ACC_BN);
      const bool empty_k = ((int64_t)split * g.kblocks_per_split) >= g.kblocks_total;
      for (int c = 0; c < g.bn; c += 32) {
        // operand streams of this chunk are requested first; the TMEM loads + transpose below (~1K clocks)
        // cover most of their latency without a second register buffer
        prefetch(c, a1, a2);
        const int64_t nq = nbase + c + c4 * 4;
        const bool vec_ok = g.epi_vec && (c + c4 * 4 < g.bn) && (nq + 3 < g.N);
        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (vec_ok && bias) bv = __ldg(reinterpret_cast<const float4*>(bias + nq));
        if (warp == 0 && lane == 0) trace_ev(g.trace, 3, tcount, 10, (unsigned)c);
        // TMEM -> registers (this lane = one row), main + cross-term accumulators, fp32 RN add
        uint32_t r[16], r2[16];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          if (c + 16 * half < g.bn) {
            tc_ld16(t_row + (uint32_t)(c + 16 * half), r);
            tc_ld16(t_row + (uint32_t)(cross_off + c + 16 * half), r2);
            tc_wait_ld();
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float4 v;
              v.x = empty_k ? 0.f : __fadd_rn(__uint_as_float(r[4 * q + 0]), __uint_as_float(r2[4 * q + 0]));
              v.y = empty_k ? 0.f : __fadd_rn(__uint_as_float(r[4 * q + 1]), __uint_as_float(r2[4 * q + 1]));
              v.z = empty_k ? 0.f : __fadd_rn(__uint_as_float(r[4 * q + 2]), __uint_as_float(r2[4 * q + 2]));
              v.w = empty_k ? 0.f : __fadd_rn(__uint_as_float(r[4 * q + 3]), __uint_as_float(r2[4 * q + 3]));
              *reinterpret_cast<float4*>(T + lane * EPI_LD + 16 * half + 4 * q) = v;
            }
          }
        }
        __syncwarp();
        if (warp == 0 && lane == 0) trace_ev(g.trace, 3, tcount, 11, (unsigned)c);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int row = rr + 4 * i;
          const int64_t m = m_base + row;
          if (m < g.M && (c + c4 * 4 < g.bn)) {
            const float4 v = *reinterpret_cast<const float4*>(T + row * EPI_LD + c4 * 4);
            const int64_t o = m * er.ldc + nq;
            if (vec_ok) {
              epilogue_f4<KIND, ACT>(er, o, v, a1[i], a2[i], bv);
            } else {
              const float vv[4] = {v.x, v.y, v.z, v.w};
              for (int j = 0; j < 4; ++j)
                if (nq + j < g.N) epilogue_scalar<KIND, ACT>(er, p1, p2, bias, o + j, nq + j, vv[j]);
            }
          }
          if (i == 0 && warp == 0 && lane == 0) trace_ev(g.trace, 3, tcount, 13, (unsigned)c);
          if (i == 3 && warp == 0 && lane == 0) trace_ev(g.trace, 3, tcount, 14, (unsigned)c);
        }
        __syncwarp();                        // staging tile is rewritten by the next chunk
        if (warp == 0 && lane == 0) trace_ev(g.trace, 3, tcount, 12, (unsigned)c);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (warp == 0 && lane == 0) trace_ev(g.trace, 3, tcount, 7, (unsigned)tile);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
  }
}

// lo plane of a whole matrix: dst = tf32_rn(x - trunc_tf32(x)), bit-identical to the in-kernel converters
__global__ void __launch_bounds__(256) split_lo_kernel(const float4* __restrict__ src, float4* __restrict__ dst, int64_t nvec) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 x = src[i];
    float4 l;
    l.x = tf32_lo_of(x.x);
    l.y = tf32_lo_of(x.y);
    l.z = tf32_lo_of(x.z);
    l.w = tf32_lo_of(x.w);
    dst[i] = l;
  }
}
// caller-owned scratch for the precomputed B_lo plane (krs_gemm_set_workspace); calls that use it must be issued
// on one stream at a time
struct Workspace { void* ptr; size_t bytes; int device; };
constexpr int MAX_DEVICES = 64;
std::mutex g_ws_mu;
Workspace g_ws_val[MAX_DEVICES];           // one registration per device (a process may drive several GPUs)
Workspace get_ws(int dev) {
  std::lock_guard<std::mutex> l(g_ws_mu);
  return (dev >= 0 && dev < MAX_DEVICES) ? g_ws_val[dev] : Workspace{nullptr, 0, -1};
}

// debug / tuning overrides, read ONCE per process (they used to cost ~10 getenv calls per GEMM launch)
struct EnvKnobs {
  int no_mask = 1, mn_layout = 1, mn_swz = (int)CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, mn_lbo = 2048, mn_sbo = 512, mn_kstep = 1024;
  int fuse_n = 1, b_lo_tma = 1, allow3d = 1;
  uint32_t wait_ns = 0x400;
};
const EnvKnobs& env_knobs() {
  static EnvKnobs k;
  static std::once_flag once;
  std::call_once(once, [] {
    auto geti = [](const char* name, int& v) { if (const char* e = getenv(name)) v = atoi(e); };
    geti("KRS_TC_NO_MASK", k.no_mask);
    geti("KRS_TC_MN_LAYOUT", k.mn_layout);
    geti("KRS_TC_MN_SWZ", k.mn_swz);
    geti("KRS_TC_MN_LBO", k.mn_lbo);
    geti("KRS_TC_MN_SBO", k.mn_sbo);
    geti("KRS_TC_MN_KSTEP", k.mn_kstep);
    geti("KRS_TC_FUSE_N", k.fuse_n);
    geti("KRS_TC_B_LO_TMA", k.b_lo_tma);
    if (const char* e = getenv("KRS_TC_WAIT_NS")) k.wait_ns = (uint32_t)atoi(e);
    if (getenv("KRS_TC_NO_3D") != nullptr) k.allow3d = 0;
  });
  return k;
}

// ---------------------------------------------------------------- host side
// MN-contiguous matrix [rows = K][cols = MN] viewed as 3-D {32, K, MN/32}: one box {32, BK, blocks} lands in
// shared memory as `blocks` consecutive [BK rows x 128 B] sub-tiles — exactly the MN-major UMMA layout.
bool make_map_3d(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld, int blocks, CUtensorMapSwizzle swz) {
  auto fn = encode_fn();
  if (!fn || (cols % 32) != 0) return false;
  cuuint64_t dims[3] = {32, (cuuint64_t)rows, (cuuint64_t)(cols / 32)};
  cuuint64_t strides[2] = {(cuuint64_t)ld * sizeof(float), 128};
  cuuint32_t box[3] = {32, (cuuint32_t)BK, (cuuint32_t)blocks};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

using TcKernel = void (*)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const TcArgs);
template <int VER>
TcKernel kernel_table_v(int kind, int act) {
  switch (kind) {
    case TK_STORE: return gemm_tc_kernel<VER, TK_STORE, TA_LINEAR>;
    case TK_ACCUM: return gemm_tc_kernel<VER, TK_ACCUM, TA_LINEAR>;
    case TK_ATOMIC: return gemm_tc_kernel<VER, TK_ATOMIC, TA_LINEAR>;
    case TK_ADD2: return gemm_tc_kernel<VER, TK_ADD2, TA_LINEAR>;
    case TK_BIAS_ACT:
      return act == TA_LINEAR ? gemm_tc_kernel<VER, TK_BIAS_ACT, TA_LINEAR>
                              : (act == TA_RELU ? gemm_tc_kernel<VER, TK_BIAS_ACT, TA_RELU> : gemm_tc_kernel<VER, TK_BIAS_ACT, TA_GENERIC>);
    case TK_CROSS:
      return act == TA_LINEAR ? gemm_tc_kernel<VER, TK_CROSS, TA_LINEAR>
                              : (act == TA_RELU ? gemm_tc_kernel<VER, TK_CROSS, TA_RELU> : gemm_tc_kernel<VER, TK_CROSS, TA_GENERIC>);
  }
  return nullptr;
}
// nullptr for (kind, act) pairs that are never launched (activation classes of kinds without an activation)
TcKernel kernel_table(int ver, int kind, int act) {
  const bool has_act = kind == TK_BIAS_ACT || kind == TK_CROSS;
  if (!has_act && act != TA_LINEAR) return nullptr;
  return ver == 2 ? kernel_table_v<2>(kind, act) : kernel_table_v<1>(kind, act);
}

int pick_bn(int64_t N, bool b_mn_major, int max_bn) {
  const int gran = b_mn_major ? 32 : 16;
  const int64_t tiles = ceil_div<int64_t>(N, max_bn);
  int64_t bn = ceil_div<int64_t>(ceil_div<int64_t>(N, tiles), gran) * gran;
  if (bn > max_bn) bn = max_bn;
  if (bn < gran) bn = gran;
  return (int)bn;
}

}  // namespace

long long gemm_tc_launches() { return g_tc_launches.load(); }

int gemm_tc(const float* A, int64_t lda, bool transA, const float* B, int64_t ldb, bool transB, float* C, int64_t ldc,
            int64_t M, int64_t N, int64_t K, const Epilogue& epi, int split_k, bool accumulate, cudaStream_t stream, int ver) {
  // shapes the tensor-core path does not cover fall back to the exact FFMA engine
  if (M < 64 || N < 16 || K < 16) return KRS_EUNSUPPORTED;
  if (!aligned16(A) || !aligned16(B) || (lda % 4) != 0 || (ldb % 4) != 0) return KRS_EUNSUPPORTED;
  if (M >= ((int64_t)1 << 31) || N >= ((int64_t)1 << 31) || K >= ((int64_t)1 << 31)) return KRS_EUNSUPPORTED;
  if (split_k < 1) split_k = 1;
  KRS_REQUIRE(split_k == 1 || epi.kind == EPI_NONE, "gemm_tc: split-K only with the plain epilogue");
  if (encode_fn() == nullptr) return KRS_EUNSUPPORTED;
  // bound the number of truncating accumulations per TMEM tile (see the header): plain-epilogue GEMMs
  // with a long reduction are split so that each partial tile covers at most KC_MAX of K
  if (epi.kind == EPI_NONE && !accumulate && K > KC_MAX) split_k = (int)imax<int64_t>(split_k, ceil_div<int64_t>(K, KC_MAX));

  TcArgs g;
  g.C = C; g.ldc = ldc; g.M = M; g.N = N; g.K = K;
  g.a_mn_major = transA ? 1 : 0;       // A stored (K,M): M contiguous
  g.b_mn_major = transB ? 0 : 1;       // B stored (K,N): N contiguous ; transB: stored (N,K): K contiguous
  g.bn = pick_bn(N, g.b_mn_major != 0, ver == 2 ? Cfg<2>::ACC_BN : Cfg<1>::ACC_BN);
  g.tiles_m = (int)ceil_div<int64_t>(M, BM);
  g.tiles_n = (int)ceil_div<int64_t>(N, g.bn);
  const int ksub = ver == 2 ? Cfg<2>::KSUB : Cfg<1>::KSUB;
  g.kblocks_total = ceil_div<int64_t>(K, BK * ksub);          // ring stages along K (KSUB sub-blocks of 16 each)
  g.kblocks_per_split = ceil_div<int64_t>(g.kblocks_total, split_k);
  g.splits = (int)ceil_div<int64_t>(g.kblocks_total, g.kblocks_per_split);
  g.atomic_out = g.splits > 1 ? 1 : 0;
  g.accumulate = accumulate ? 1 : 0;
  g.epi = epi;
  {
    auto ok16 = [](const void* q) { return q == nullptr || aligned16(q); };
    g.epi_vec = ((ldc % 4) == 0 && aligned16(C) && ok16(epi.x0) && ok16(epi.x) && ok16(epi.h2_out) && ok16(epi.z_out) &&
                 ok16(epi.add1) && ok16(epi.add2) && ok16(epi.bias) && (g.bn % 16) == 0 && ((int64_t)g.bn % 4) == 0) ? 1 : 0;
  }
  // the tensor core ignores the low 13 mantissa bits of a tf32 operand: leaving hi = raw fp32 is bit-identical
  // to masking it (tests/tc_stress.py, both modes) and saves a third of the converter's shared-memory stores
  g.trace = g_trace.load();
  const EnvKnobs& env = env_knobs();
  g.no_mask = env.no_mask;
  g.mn_lbo = env.mn_lbo; g.mn_sbo = env.mn_sbo; g.mn_kstep = env.mn_kstep; g.mn_layout = env.mn_layout;
  const int mn_swz = env.mn_swz;
  g.fuse_n = env.fuse_n;
  if (2 * g.bn > 256) g.fuse_n = 0;
  // B_lo plane precomputed once per call into the caller-registered workspace when B is small and re-read by many
  // m-tiles (weights): the converters then touch only the A tile (28 -> 16 elements and 10 -> 4 shared-memory
  // instructions per thread and k-block in VER 2)
  g.wait_ns = env.wait_ns;
  g.b_lo_tma = 0;
  const float* B_lo = nullptr;
  {
    const int64_t b_rows = transB ? N : K;
    const size_t b_plane = (size_t)b_rows * (size_t)ldb * sizeof(float);
    int cur_dev = -1;
    if (cudaGetDevice(&cur_dev) != cudaSuccess) cur_dev = -1;
    Workspace w = get_ws(cur_dev);
    const bool want = env.b_lo_tma != 0 && w.ptr != nullptr && b_plane <= w.bytes && M >= 4 * N && cur_dev == w.device;
    if (want) {
      split_lo_kernel<<<(unsigned)imin<int64_t>(2 * sm_count(), ceil_div<int64_t>((int64_t)(b_plane / 16), 256)), 256, 0, stream>>>(
          reinterpret_cast<const float4*>(B), reinterpret_cast<float4*>(w.ptr), (int64_t)(b_plane / 16));
      KRS_LAUNCH_CHECK();
      g_split_launches.fetch_add(1);
      B_lo = reinterpret_cast<const float*>(w.ptr);
      g.b_lo_tma = 1;
    }
  }
  // VER 2: the MN-major A tile is only read by converter threads (consecutive lanes = consecutive words): no swizzle
  const int a_mn_swz = ver == 2 ? (int)CU_TENSOR_MAP_SWIZZLE_NONE : mn_swz;

  CUtensorMap ma, mb, mblo;
  bool ok;
  g.a_3d = g.b_3d = 0;
  const bool allow3d = env.allow3d != 0;
  if (!g.a_mn_major) ok = make_map(&ma, A, M, K, lda, BK, BM, CU_TENSOR_MAP_SWIZZLE_64B);        // [M][K]
  else {
    ok = allow3d && make_map_3d(&ma, A, K, M, lda, BM / 32, (CUtensorMapSwizzle)a_mn_swz);
    if (ok) g.a_3d = 1;
    else ok = make_map(&ma, A, K, M, lda, 32, BK, (CUtensorMapSwizzle)a_mn_swz);                 // [K][M]
  }
  if (!ok) return KRS_EUNSUPPORTED;
  if (!g.b_mn_major) ok = make_map(&mb, B, N, K, ldb, BK, g.bn, CU_TENSOR_MAP_SWIZZLE_64B);      // [N][K]
  else {
    ok = allow3d && make_map_3d(&mb, B, K, N, ldb, g.bn / 32, (CUtensorMapSwizzle)mn_swz);
    if (ok) g.b_3d = 1;
    else ok = make_map(&mb, B, K, N, ldb, 32, BK, (CUtensorMapSwizzle)mn_swz);                   // [K][N]
  }
  if (!ok) return KRS_EUNSUPPORTED;
  if (g.b_lo_tma) {
    if (!g.b_mn_major) ok = make_map(&mblo, B_lo, N, K, ldb, BK, g.bn, CU_TENSOR_MAP_SWIZZLE_64B);
    else if (g.b_3d) ok = make_map_3d(&mblo, B_lo, K, N, ldb, g.bn / 32, (CUtensorMapSwizzle)mn_swz);
    else ok = make_map(&mblo, B_lo, K, N, ldb, 32, BK, (CUtensorMapSwizzle)mn_swz);
    if (!ok) return KRS_EUNSUPPORTED;
  } else {
    mblo = mb;
  }

  if (g.atomic_out && !accumulate)
    KRS_CUDA(cudaMemset2DAsync(C, (size_t)ldc * sizeof(float), 0, (size_t)N * sizeof(float), (size_t)M, stream));

  const size_t stage_bytes = ver == 2 ? (size_t)ksub * (size_t)(A_BYTES + 2 * g.bn * BK * 4) : 2 * (size_t)(A_BYTES + g.bn * BK * 4);
  const size_t budget = 227 * 1024 - 1024 - 512 - EPI_SMEM_BYTES;   // alignment slack + barriers + epilogue staging
  g.stages = (int)imin<int64_t>(MAX_STAGES, (int64_t)(budget / stage_bytes));
  if (g.stages < 3) return KRS_EUNSUPPORTED;
  const size_t smem = 1024 + g.stages * stage_bytes + EPI_SMEM_BYTES + 512;
  // the dynamic shared memory opt-in is a PER-DEVICE function attribute: once per device, not once per process
  static std::once_flag once[MAX_DEVICES];
  static cudaError_t attr_err[MAX_DEVICES];
  {
    int dev = 0;
    KRS_CUDA(cudaGetDevice(&dev));
    KRS_REQUIRE(dev >= 0 && dev < MAX_DEVICES, "gemm_tc: device index %d out of range", dev);
    std::call_once(once[dev], [dev] {
      attr_err[dev] = cudaSuccess;
      for (int v = 0; v < 2 && attr_err[dev] == cudaSuccess; ++v)
        for (int k = 0; k < TK_COUNT && attr_err[dev] == cudaSuccess; ++k)
          for (int a = 0; a < TA_COUNT && attr_err[dev] == cudaSuccess; ++a)
            if (kernel_table(v + 1, k, a))
              attr_err[dev] = cudaFuncSetAttribute(kernel_table(v + 1, k, a), cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    });
    KRS_CUDA(attr_err[dev]);
  }
  const int64_t total_tiles = (int64_t)g.tiles_m * g.tiles_n * g.splits;
  const unsigned grid = (unsigned)imax<int64_t>(1, imin<int64_t>(total_tiles, sm_count()));
  int kind;
  switch (epi.kind) {
    case EPI_NONE: kind = g.atomic_out ? TK_ATOMIC : (g.accumulate ? TK_ACCUM : TK_STORE); break;
    case EPI_BIAS_ACT: kind = TK_BIAS_ACT; break;
    case EPI_CROSS: kind = TK_CROSS; break;
    default: kind = TK_ADD2; break;
  }
  const bool has_act = kind == TK_BIAS_ACT || kind == TK_CROSS;
  const int actc = !has_act ? TA_LINEAR : (epi.act == KRS_ACT_LINEAR ? TA_LINEAR : (epi.act == KRS_ACT_RELU ? TA_RELU : TA_GENERIC));
  TcKernel kfn = kernel_table(ver == 2 ? 2 : 1, kind, actc);
  KRS_REQUIRE(kfn != nullptr, "gemm_tc: no kernel for kind %d act %d", kind, actc);
  kfn<<<grid, NUM_THREADS, smem, stream>>>(ma, mb, mblo, g);
  KRS_LAUNCH_CHECK();
  g_tc_launches.fetch_add(1);
  return KRS_OK;
}

void gemm_tc_set_trace(unsigned long long* p) { g_trace.store(p); }
}  // namespace krs

extern "C" int krs_gemm_set_workspace(void* dev_buf, size_t bytes) {
  if (dev_buf != nullptr && !krs::aligned16(dev_buf)) {
    krs::set_error("krs_gemm_set_workspace: buffer must be 16-byte aligned");
    return KRS_EINVAL;
  }
  int dev = -1;
  if (dev_buf != nullptr) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, dev_buf) != cudaSuccess || at.type != cudaMemoryTypeDevice) {
      (void)cudaGetLastError();
      krs::set_error("krs_gemm_set_workspace: not a device pointer");
      return KRS_EINVAL;
    }
    dev = at.device;
  }
  if (dev_buf == nullptr && cudaGetDevice(&dev) != cudaSuccess) dev = -1;     // NULL unregisters the current device's buffer
  if (dev < 0 || dev >= krs::MAX_DEVICES) {
    krs::set_error("krs_gemm_set_workspace: device index %d out of range", dev);
    return KRS_EINVAL;
  }
  { std::lock_guard<std::mutex> l(krs::g_ws_mu); krs::g_ws_val[dev] = krs::Workspace{dev_buf, dev_buf ? bytes : 0, dev_buf ? dev : -1}; }
  return KRS_OK;
}
extern "C" long long krs_gemm_split_launch_count(void) { return krs::g_split_launches.load(); }
extern "C" long long krs_gemm_tc_launch_count(void) { return krs::gemm_tc_launches(); }
extern "C" int krs_gemm_tc_set_trace(void* dev_buf) {
  krs::gemm_tc_set_trace(reinterpret_cast<unsigned long long*>(dev_buf));
  return KRS_OK;
}
