// dot.cu — DLRM DotInteraction forward / backward.
//
// Replaces DotInteraction.call (dot_interaction.py:134-205): stack -> batched F F^T -> lower
// triangle take (or mask).  Per sample this is a 27x27x128 Gram — tiny, and the op is HBM-bound
// (998 MB at C3), so the design goal is: read every feature row exactly once straight
// from its source tensor (no stack copy, strided views of the concatenated gather output are
// fine), never touch shared memory for operands, write only the selected triangle.
//
// One warp per sample; the Gram runs on warp-level mma.sync m16n8k8 TF32 with a 3-term split
// (hi*hi + hi*lo + lo*hi, fp32 accumulate) so results keep fp32-level accuracy (|err| ~1e-6 rel).
// tcgen05 is deliberately NOT used here: a 32x32 per-sample product cannot fill a 128-row UMMA
// tile and the kernel is bandwidth-bound anyway.  Because A and B fragments come from the SAME
// registers, the k-index <-> memory mapping is free, which lets each lane fetch 16-byte vectors.
#include "common.cuh"

namespace krs {
namespace {

constexpr int MAXN = 32;

struct DotParams {
  const float* feat[MAXN];
  int64_t stride[MAXN];
  float* dfeat[MAXN];
  int64_t dstride[MAXN];
  int N, E;
  int64_t B;
  int self_interaction, skip_gather;
  int out_dim;
  float* out;          // fwd output
  const float* gout;   // bwd incoming gradient
};

__device__ __forceinline__ uint32_t tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
// 3xTF32 split with plain integer / fp32 ops: hi = x truncated to TF32, lo = (x - hi) rounded to TF32 by adding half an
// ulp of the 13 dropped bits.  cvt.rna.tf32.f32 is emulated with ~4 instructions on sm_100a and this kernel splits every
// element it reads (the split was ~40 % of its issue slots); the residual x - hi is exact, so the pair keeps ~21 bits.
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(x) & 0xFFFFE000u;
  lo = (__float_as_uint(x - __uint_as_float(hi)) + 0x1000u) & 0xFFFFE000u;
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__device__ __forceinline__ int tri_index(int i, int j, int self) { return self ? i * (i + 1) / 2 + j : i * (i - 1) / 2 + j; }

// ------------------------------------------------------------------ forward (mma path)
// N <= 32, E % 16 == 0, every feature pointer 16-byte aligned with stride % 4 == 0.
__global__ void __launch_bounds__(128) dot_fwd_mma_kernel(const __grid_constant__ DotParams p) {
  // per-warp staging of one sample's outputs: the selected triangle is written to global memory as ONE contiguous run
  // (out_dim floats, full 128-byte lines) instead of 351 scattered 4-byte stores (the forward was store-bound: 48 % of HBM)
  __shared__ float stage[4][32 * 32];
  float* st = stage[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int N = p.N, E = p.E, self = p.self_interaction;
  for (int64_t b = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; b < p.B; b += nwarps) {
    const float* rp[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = g + 8 * i;
      rp[i] = row < N ? p.feat[row] + b * p.stride[row] + 4 * t : nullptr;
    }
    float acc[2][4][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[mt][nt][q] = 0.f;
    // software pipeline: the next 16-wide k chunk is in flight while the current one feeds the tensor core
    // (the mma asm blocks are volatile, so the compiler would not hoist the loads by itself; with one chunk in
    // flight per warp the kernel reached only 46 % of the HBM roofline)
    float4 x[4], xn[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = rp[i] ? ldg_nc_f4(rp[i]) : make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k0 = 0; k0 < E; k0 += 16) {
      const bool more = k0 + 16 < E;
#pragma unroll
      for (int i = 0; i < 4; ++i) xn[i] = (more && rp[i]) ? ldg_nc_f4(rp[i] + k0 + 16) : make_float4(0.f, 0.f, 0.f, 0.f);
      // two k8 steps per 16-wide chunk: step 0 uses (.x,.y) as k slots (t, t+4); step 1 uses (.z,.w)
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        uint32_t hi[4][2], lo[4][2];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          split_tf32(s == 0 ? x[i].x : x[i].z, hi[i][0], lo[i][0]);
          split_tf32(s == 0 ? x[i].y : x[i].w, hi[i][1], lo[i][1]);
        }
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) {
            if (nt > 2 * mt + 1) continue;   // upper-triangle tiles are never needed
            // A(16x8): a0=(g,t) a1=(g+8,t) a2=(g,t+4) a3=(g+8,t+4) ; B(8x8): b0=(k=t,n=g) b1=(k=t+4,n=g)
            mma_tf32(acc[mt][nt], lo[2 * mt][0], lo[2 * mt + 1][0], lo[2 * mt][1], lo[2 * mt + 1][1], hi[nt][0], hi[nt][1]);
            mma_tf32(acc[mt][nt], hi[2 * mt][0], hi[2 * mt + 1][0], hi[2 * mt][1], hi[2 * mt + 1][1], lo[nt][0], lo[nt][1]);
            mma_tf32(acc[mt][nt], hi[2 * mt][0], hi[2 * mt + 1][0], hi[2 * mt][1], hi[2 * mt + 1][1], hi[nt][0], hi[nt][1]);
          }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) x[i] = xn[i];
    }
    float* o = p.out + b * (int64_t)p.out_dim;
    __syncwarp();                              // the previous sample's staged row has been read
    if (p.skip_gather)                         // N x N layout: entries above the kept triangle are exact zeros
      for (int idx = lane; idx < p.out_dim; idx += 32) st[idx] = 0.f;
    __syncwarp();
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        if (nt > 2 * mt + 1) continue;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int i = 16 * mt + g + ((q & 2) ? 8 : 0);
          const int j = 8 * nt + 2 * t + (q & 1);
          if (i < N && j < N && (j < i || (self && j == i))) {
            if (p.skip_gather) st[i * N + j] = acc[mt][nt][q];
            else st[tri_index(i, j, self)] = acc[mt][nt][q];
          }
        }
      }
    __syncwarp();
    for (int idx = lane; idx < p.out_dim; idx += 32) __stcs(o + idx, st[idx]);
  }
}

// ------------------------------------------------------------------ backward (mma path)
// dF = S F with S = G + G^T (32x32, zero padded) staged per warp in shared memory.
// Requirements: N <= 32, E % 32 == 0, 16-byte aligned rows.
__global__ void __launch_bounds__(128) dot_bwd_mma_kernel(const __grid_constant__ DotParams p) {
  __shared__ float S[4][32][33];
  __shared__ float G[4][32 * 32];          // the sample's incoming gradient, loaded with coalesced 128-byte reads
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int N = p.N, E = p.E, self = p.self_interaction;
  float(*Sw)[33] = S[warp];
  for (int64_t b = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; b < p.B; b += nwarps) {
    const float* gsrc = p.gout + b * (int64_t)p.out_dim;
    float* go = G[warp];
    __syncwarp();
    for (int idx = lane; idx < p.out_dim; idx += 32) go[idx] = __ldcs(gsrc + idx);
    __syncwarp();
    for (int idx = lane; idx < 32 * 32; idx += 32) {
      const int i = idx >> 5, j = idx & 31;
      float v = 0.f;
      if (i < N && j < N) {
        const int hi_ = max(i, j), lo_ = min(i, j);
        if (hi_ != lo_) v = p.skip_gather ? go[hi_ * N + lo_] : go[tri_index(hi_, lo_, self)];
        else if (self) v = 2.f * (p.skip_gather ? go[i * N + i] : go[tri_index(i, i, 1)]);
      }
      Sw[i][j] = v;
    }
    __syncwarp();
    // A fragments of S (split once per sample): [mt][kt] -> a0..a3
    uint32_t ah[2][4][4], al[2][4][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int kt = 0; kt < 4; ++kt) {
        split_tf32(Sw[16 * mt + g][8 * kt + t], ah[mt][kt][0], al[mt][kt][0]);
        split_tf32(Sw[16 * mt + g + 8][8 * kt + t], ah[mt][kt][1], al[mt][kt][1]);
        split_tf32(Sw[16 * mt + g][8 * kt + t + 4], ah[mt][kt][2], al[mt][kt][2]);
        split_tf32(Sw[16 * mt + g + 8][8 * kt + t + 4], ah[mt][kt][3], al[mt][kt][3]);
      }
    // B operands (feature rows) of one 32-wide chunk: 2 x float4 per k-tile, all 8 loads issued together, and the NEXT
    // chunk's loads are in flight while this one feeds the tensor core.  (Loading f0 / f1 inside the k-tile loop, right
    // before their split, exposed one global-load latency per k-tile — 16 per sample — and held the backward at ~21 % of
    // the HBM roofline; the mma asm is volatile, so the compiler cannot hoist the loads itself.)
    float4 fc[4][2], fn[4][2];
    auto load_chunk = [&](float4 (&dst)[4][2], int e0) {
#pragma unroll
      for (int kt = 0; kt < 4; ++kt) {
        const int r0 = 8 * kt + t, r1 = 8 * kt + t + 4;
        dst[kt][0] = r0 < N ? ldg_nc_f4(p.feat[r0] + b * p.stride[r0] + e0 + 4 * g) : make_float4(0.f, 0.f, 0.f, 0.f);
        dst[kt][1] = r1 < N ? ldg_nc_f4(p.feat[r1] + b * p.stride[r1] + e0 + 4 * g) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    load_chunk(fc, 0);
    for (int e0 = 0; e0 < E; e0 += 32) {
      if (e0 + 32 < E) load_chunk(fn, e0 + 32);
      float acc[2][4][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[mt][nt][q] = 0.f;
#pragma unroll
      for (int kt = 0; kt < 4; ++kt) {
        // B(k = feature, n): column n=g of n-tile nt maps to e = e0 + 4g + nt  => one float4 per row
        const float4 f0 = fc[kt][0], f1 = fc[kt][1];
        const float b0v[4] = {f0.x, f0.y, f0.z, f0.w};
        const float b1v[4] = {f1.x, f1.y, f1.z, f1.w};
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          uint32_t bh0, bl0, bh1, bl1;
          split_tf32(b0v[nt], bh0, bl0);
          split_tf32(b1v[nt], bh1, bl1);
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            mma_tf32(acc[mt][nt], al[mt][kt][0], al[mt][kt][1], al[mt][kt][2], al[mt][kt][3], bh0, bh1);
            mma_tf32(acc[mt][nt], ah[mt][kt][0], ah[mt][kt][1], ah[mt][kt][2], ah[mt][kt][3], bl0, bl1);
            mma_tf32(acc[mt][nt], ah[mt][kt][0], ah[mt][kt][1], ah[mt][kt][2], ah[mt][kt][3], bh0, bh1);
          }
        }
      }
      // C(row i, col n): c0=(g,2t) c1=(g,2t+1) c2=(g+8,2t) c3=(g+8,2t+1); e = e0 + 4n + nt
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int i = 16 * mt + g + 8 * half;
          if (i < N) {
            float* d = p.dfeat[i] + b * p.dstride[i] + e0 + 8 * t;
            *reinterpret_cast<float4*>(d) = make_float4(acc[mt][0][2 * half], acc[mt][1][2 * half], acc[mt][2][2 * half],
                                                        acc[mt][3][2 * half]);
            *reinterpret_cast<float4*>(d + 4) = make_float4(acc[mt][0][2 * half + 1], acc[mt][1][2 * half + 1],
                                                            acc[mt][2][2 * half + 1], acc[mt][3][2 * half + 1]);
          }
        }
#pragma unroll
      for (int kt = 0; kt < 4; ++kt) { fc[kt][0] = fn[kt][0]; fc[kt][1] = fn[kt][1]; }
    }
  }
}

// ------------------------------------------------------------------ generic fallbacks (any N <= 32, any E)
__global__ void dot_fwd_naive_kernel(const __grid_constant__ DotParams p) {
  const int64_t total = p.B * (int64_t)p.out_dim;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = idx / p.out_dim;
    const int o = (int)(idx - b * p.out_dim);
    int i, j;
    bool keep = true;
    if (p.skip_gather) {
      i = o / p.N;
      j = o - i * p.N;
      keep = (j < i) || (p.self_interaction && j == i);
    } else {
      // invert tri_index: largest i with base(i) <= o
      i = 0;
      if (p.self_interaction) { while ((i + 1) * (i + 2) / 2 <= o) ++i; j = o - i * (i + 1) / 2; }
      else { i = 1; while ((i + 1) * i / 2 <= o) ++i; j = o - i * (i - 1) / 2; }
    }
    float acc = 0.f;
    if (keep) {
      const float* a = p.feat[i] + b * p.stride[i];
      const float* c = p.feat[j] + b * p.stride[j];
      for (int e = 0; e < p.E; ++e) acc = fmaf(a[e], c[e], acc);
    }
    p.out[idx] = acc;
  }
}

__global__ void dot_bwd_naive_kernel(const __grid_constant__ DotParams p) {
  const int64_t total = p.B * (int64_t)p.N * p.E;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int e = (int)(idx % p.E);
    const int i = (int)((idx / p.E) % p.N);
    const int64_t b = idx / ((int64_t)p.E * p.N);
    const float* go = p.gout + b * (int64_t)p.out_dim;
    float acc = 0.f;
    for (int j = 0; j < p.N; ++j) {
      const int hi_ = max(i, j), lo_ = min(i, j);
      float s = 0.f;
      if (hi_ != lo_) s = p.skip_gather ? go[hi_ * p.N + lo_] : go[tri_index(hi_, lo_, p.self_interaction)];
      else if (p.self_interaction) s = 2.f * (p.skip_gather ? go[i * p.N + i] : go[tri_index(i, i, 1)]);
      acc = fmaf(s, p.feat[j][b * p.stride[j] + e], acc);
    }
    p.dfeat[i][b * p.dstride[i] + e] = acc;
  }
}

int fill(DotParams& p, const float* const* feats, const int64_t* strides, int N, int E, int64_t B, int self_i, int skip) {
  KRS_REQUIRE(feats && strides, "dot: null feature table");
  KRS_REQUIRE(N >= 1 && N <= MAXN, "dot: number of features must be in 1..%d, got %d", MAXN, N);
  KRS_REQUIRE(E >= 1 && B >= 0, "dot: bad E/B");
  for (int i = 0; i < N; ++i) {
    KRS_REQUIRE(feats[i] != nullptr, "dot: feature %d is null", i);
    p.feat[i] = feats[i];
    p.stride[i] = strides[i];
    p.dfeat[i] = nullptr;
    p.dstride[i] = 0;
  }
  p.N = N; p.E = E; p.B = B;
  p.self_interaction = self_i ? 1 : 0;
  p.skip_gather = skip ? 1 : 0;
  p.out_dim = skip ? N * N : (self_i ? N * (N + 1) / 2 : N * (N - 1) / 2);   // dot_interaction.py:207-222
  p.out = nullptr;
  p.gout = nullptr;
  return KRS_OK;
}

bool rows_vec_ok(const DotParams& p, bool grads) {
  for (int i = 0; i < p.N; ++i) {
    if (!aligned16(p.feat[i]) || (p.stride[i] % 4) != 0) return false;
    if (grads && (!aligned16(p.dfeat[i]) || (p.dstride[i] % 4) != 0)) return false;
  }
  return true;
}

}  // namespace
}  // namespace krs

using namespace krs;

extern "C" int krs_dot_fwd(const float* const* feats, const int64_t* strides, int N, int E, int64_t B,
                           int self_interaction, int skip_gather, float* out, void* stream) {
  DotParams p;
  int rc = fill(p, feats, strides, N, E, B, self_interaction, skip_gather);
  if (rc) return rc;
  KRS_REQUIRE(out != nullptr, "krs_dot_fwd: null output");
  p.out = out;
  if (B == 0 || p.out_dim == 0) return KRS_OK;
  cudaStream_t s = as_stream(stream);
  if ((E % 16 == 0) && rows_vec_ok(p, false)) {
    const unsigned grid = (unsigned)krs::imax<int64_t>(1, krs::imin<int64_t>(ceil_div<int64_t>(B, 4), (int64_t)sm_count() * 16));
    dot_fwd_mma_kernel<<<grid, 128, 0, s>>>(p);
  } else {
    const int64_t total = B * (int64_t)p.out_dim;
    const unsigned grid = (unsigned)krs::imax<int64_t>(1, krs::imin<int64_t>(ceil_div<int64_t>(total, 256), (int64_t)sm_count() * 32));
    dot_fwd_naive_kernel<<<grid, 256, 0, s>>>(p);
  }
  KRS_LAUNCH_CHECK();
  return KRS_OK;
}

extern "C" int krs_dot_bwd(const float* const* feats, const int64_t* strides, const float* gout, float* const* dfeats,
                           const int64_t* dstrides, int N, int E, int64_t B, int self_interaction, int skip_gather,
                           void* stream) {
  DotParams p;
  int rc = fill(p, feats, strides, N, E, B, self_interaction, skip_gather);
  if (rc) return rc;
  KRS_REQUIRE(gout && dfeats && dstrides, "krs_dot_bwd: null argument");
  for (int i = 0; i < N; ++i) {
    KRS_REQUIRE(dfeats[i] != nullptr, "krs_dot_bwd: gradient buffer %d is null", i);
    p.dfeat[i] = dfeats[i];
    p.dstride[i] = dstrides[i];
  }
  p.gout = gout;
  if (B == 0) return KRS_OK;
  cudaStream_t s = as_stream(stream);
  if (p.out_dim == 0) {   // single feature without self interaction: gradient is zero
    for (int i = 0; i < N; ++i)
      KRS_CUDA(cudaMemset2DAsync(p.dfeat[i], sizeof(float) * (size_t)p.dstride[i], 0, sizeof(float) * (size_t)E, (size_t)B, s));
    return KRS_OK;
  }
  if ((E % 32 == 0) && rows_vec_ok(p, true)) {
    const unsigned grid = (unsigned)krs::imax<int64_t>(1, krs::imin<int64_t>(ceil_div<int64_t>(B, 4), (int64_t)sm_count() * 8));
    dot_bwd_mma_kernel<<<grid, 128, 0, s>>>(p);
  } else {
    const int64_t total = B * (int64_t)N * E;
    const unsigned grid = (unsigned)krs::imax<int64_t>(1, krs::imin<int64_t>(ceil_div<int64_t>(total, 256), (int64_t)sm_count() * 32));
    dot_bwd_naive_kernel<<<grid, 256, 0, s>>>(p);
  }
  KRS_LAUNCH_CHECK();
  return KRS_OK;
}
