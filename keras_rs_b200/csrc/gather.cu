// gather.cu — fused multi-table embedding gather (forward) and scatter-add (backward).
//
// Replaces, in ONE launch for all features: keras.layers.Embedding.call == ops.take(table, ids, 0)
// (examples/dcn.py:430-435), EmbedReduce.call's weights/sum/mean/sqrtn (embed_reduce.py:162-274),
// the per-feature Python loop of DistributedEmbedding._default_device_call
// (base_distributed_embedding.py:910-928) and the ops.concatenate that follows
// (examples/dcn.py:437): rows land directly in the concatenated (B, sum E) activation.
//
// HBM-bound integer/byte work: no tensor cores.  What matters is bytes in flight per SM and
// full-sector accesses:
//   * fast path (1-hot, uniform E, E/4 a power of two): E/4 lanes per row, 16-byte
//     ld.global.nc.L1::no_allocate loads, UNROLL independent rows per lane group in flight,
//     streaming (st.global.cs) stores; the output of consecutive (b,f) items is one contiguous
//     stream, so stores are perfectly coalesced.
//   * bulk path (variant 2): rows are staged into shared memory with cp.async.bulk (TMA engine,
//     mbarrier complete_tx) and each tile leaves with ONE bulk smem->global store.
//   * generic path: any E / hotness / weights / combiner / MOD-sharded tables.
// Backward: warp handles 32 consecutive samples of one feature; duplicate ids inside the warp are
// found with __match_any_sync and their gradient rows summed in registers before one atomic per
// (row, lane) — the scatter-add oracle is jax/test_utils.py:395-417.
#include <stdlib.h>

#include "common.cuh"

namespace krs {
namespace {

constexpr int MAXF = 96;

struct GatherParams {
  krs_feature_t f[MAXF];
  int F;
  int64_t B;
  float* out;          // fwd: output ; bwd: gout (const in practice)
  int64_t out_ld;
};

template <typename IdT>
__device__ __forceinline__ int64_t load_id(const void* ids, int64_t idx) {
  return (int64_t)reinterpret_cast<const IdT*>(ids)[idx];
}
// Index rule of the reference's backend (keras.ops.take on JAX == jnp.take, default mode "fill"; oracle/np_oracle.py
// embedding_lookup): ids < 0 count from the end of the table; ids still outside [0, vocab) address no row — the forward
// returns a NaN row for them, the backward drops them.  Returns the row, or -1.
__device__ __forceinline__ int64_t resolve_id(int64_t id, int64_t vocab) {
  id += id < 0 ? vocab : 0;
  return (uint64_t)id < (uint64_t)vocab ? id : -1;
}
__device__ __forceinline__ const float* nan_row() { return reinterpret_cast<const float*>(uintptr_t(1)); }   // sentinel "row of NaN"
__device__ __forceinline__ float4 nan4() {
  const float q = __int_as_float(0x7fc00000);
  return make_float4(q, q, q, q);
}
__device__ __forceinline__ const float* row_ptr(const krs_feature_t& f, int64_t id) {
  if (f.num_shards > 1) {
    const int s = (int)(id % f.num_shards);
    return f.shard_tables[s] + (id / f.num_shards) * (int64_t)f.dim;
  }
  return f.table + id * (int64_t)f.dim;
}

// ------------------------------------------------------------------ forward, fast path
// Requirements (checked on the host): every feature 1-hot, no weights, same dim E = 4*LPR,
// out_offset = f*E, out_ld = F*E.  Rows are numbered r = b*F + f, so the output is one contiguous
// stream of R = B*F rows.
//
// ncu on the first version (profiles/r1_gather_fast_v1.md) showed it bound by the XU pipe (87%: a
// 64-bit division per row per lane) and the indexed constant cache (78%: p.f[f] per row), with DRAM
// at 41%.  This version does ONE (b,f) division per 32-row chunk, keeps the per-feature descriptors
// in shared memory, has each lane resolve ONE row address (ids read coalesced) and hands the row
// pointers to the LPR-lane groups with shuffles.
struct FeatLite {
  const float* table;
  const void* ids;
  long long vocab;
  long long stride;
  const float* const* shards;
  int nshards;
  int shard_shift;   // log2(nshards) when it is a power of two, else -1
};

template <int LPR, typename IdT, bool SHARDED>
__global__ void __launch_bounds__(256) gather_fast_kernel(const __grid_constant__ GatherParams p, int chunks_per_warp) {
  constexpr int RPW = 32 / LPR;              // rows per warp-wide load instruction
  constexpr int STEPS = LPR;                 // sub-steps to cover a 32-row chunk
  constexpr int U = STEPS < 8 ? STEPS : 8;   // loads in flight per lane
  constexpr int E = LPR * 4;
  __shared__ FeatLite sf[MAXF];
  for (int i = threadIdx.x; i < p.F; i += blockDim.x) {
    FeatLite t;
    t.table = p.f[i].table;
    t.ids = p.f[i].ids;
    t.vocab = p.f[i].vocab;
    t.stride = p.f[i].ids_stride;
    t.shards = p.f[i].shard_tables;
    t.nshards = p.f[i].num_shards;
    t.shard_shift = (t.nshards > 0 && (t.nshards & (t.nshards - 1)) == 0) ? (31 - __clz(t.nshards)) : -1;
    sf[i] = t;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane % LPR, rsub = lane / LPR;
  const unsigned F = (unsigned)p.F;
  const long long R = p.B * (long long)p.F;
  const long long chunk0 = ((long long)blockIdx.x * 8 + warp) * chunks_per_warp;
  for (int c = 0; c < chunks_per_warp; ++c) {
    const long long r0 = (chunk0 + c) * 32;
    if (r0 >= R) break;
    // (b,f) of this lane's row: one warp-uniform division, then a carry loop (F may be < 32)
    long long b = (R < 0x7fffffffLL) ? (long long)((unsigned)r0 / F) : r0 / F;
    unsigned f = (unsigned)(r0 - b * F) + lane;
    while (f >= F) { f -= F; ++b; }
    const float* src = nullptr;
    if (r0 + lane < R) {
      const FeatLite& ft = sf[f];
      const long long id = resolve_id(load_id<IdT>(ft.ids, b * ft.stride), ft.vocab);
      if (id < 0) {
        src = nan_row();
      } else if (SHARDED) {
        const int sh = ft.shard_shift;
        const int owner = sh >= 0 ? (int)(id & (ft.nshards - 1)) : (int)(id % ft.nshards);
        const long long local = sh >= 0 ? (id >> sh) : (id / ft.nshards);
        src = ft.shards[owner] + local * E;      // direct peer-mapped table (random remote rows are slow over NVLink:
                                                 // the training path routes through exchange.cu instead)
      } else {
        src = ft.table + id * E;
      }
    }
#pragma unroll
    for (int s0 = 0; s0 < STEPS; s0 += U) {
      float4 v[U];
      const float* q[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int j = (s0 + u) * RPW + rsub;
        q[u] = reinterpret_cast<const float*>(__shfl_sync(0xffffffffu, (unsigned long long)src, j));
        if (q[u] == nan_row()) v[u] = nan4();
        else if (q[u]) v[u] = ldg_nc_f4(q[u] + sub * 4);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int j = (s0 + u) * RPW + rsub;
        if (q[u]) stg_cs_f4(p.out + (r0 + j) * E + sub * 4, v[u]);
      }
    }
  }
}

// ------------------------------------------------------------------ forward, bulk-copy (TMA engine) path
// Same requirements as the fast path.  A CTA owns tiles of TILE_ROWS consecutive output rows; every
// thread issues cp.async.bulk global->shared for its rows (row = E*4 bytes, multiple of 16), all
// signalling one mbarrier; after the wait one thread issues a single bulk shared->global store of
// the whole contiguous tile.  Two tiles are in flight (double buffer).
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  const uint32_t addr = smem_u32(bar);
  while (!ok) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)),
               "r"(bytes)
               : "memory");
}

template <typename IdT>
__global__ void __launch_bounds__(256) gather_bulk_kernel(const __grid_constant__ GatherParams p, int E, int tile_rows) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t bars[2];
  float* tile[2] = {reinterpret_cast<float*>(smem_raw), reinterpret_cast<float*>(smem_raw) + (size_t)tile_rows * E};
  const int64_t R = p.B * p.F;
  const int64_t ntiles = (R + tile_rows - 1) / tile_rows;
  const uint32_t F = (uint32_t)p.F;
  const uint32_t row_bytes = (uint32_t)E * 4u;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  uint32_t phase[2] = {0, 0};
  int it = 0;
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
    const int bsel = it & 1;
    const int64_t r0 = t * tile_rows;
    const int nrows = (int)krs::imin<int64_t>(tile_rows, R - r0);
    // rows whose id addresses no table row are not copied: they are written as NaN by their thread and leave the
    // transaction count (first pass: count them)
    int bad = 0;
    for (int i = threadIdx.x; i < nrows; i += blockDim.x) {
      const int64_t r = r0 + i;
      const int64_t b = r / F;
      const int f = (int)(r - b * F);
      const krs_feature_t& ft = p.f[f];
      bad += resolve_id(load_id<IdT>(ft.ids, b * ft.ids_stride), ft.vocab) < 0;
    }
    // the bulk store that last read this buffer (2 tiles ago) must have finished reading smem
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
    const int nbad = __syncthreads_count(bad);
    if (threadIdx.x == 0) mbar_expect_tx(&bars[bsel], (uint32_t)(nrows - nbad) * row_bytes);
    __syncthreads();
    for (int i = threadIdx.x; i < nrows; i += blockDim.x) {
      const int64_t r = r0 + i;
      const int64_t b = r / F;
      const int f = (int)(r - b * F);
      const krs_feature_t& ft = p.f[f];
      const int64_t id = resolve_id(load_id<IdT>(ft.ids, b * ft.ids_stride), ft.vocab);
      if (id >= 0) {
        bulk_g2s(tile[bsel] + (size_t)i * E, ft.table + id * E, row_bytes, &bars[bsel]);
      } else {
        for (int c = 0; c < E; ++c) tile[bsel][(size_t)i * E + c] = __int_as_float(0x7fc00000);
      }
    }
    if (nbad) __syncthreads();          // generic-proxy NaN rows are visible before thread 0's async-proxy fence
    mbar_wait(&bars[bsel], phase[bsel]);
    phase[bsel] ^= 1;
    if (threadIdx.x == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      bulk_s2g(p.out + r0 * E, tile[bsel], (uint32_t)nrows * row_bytes);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// ------------------------------------------------------------------ forward, multi-hot "sample" path
// One lane group (lpr lanes, one float4 column each) per SAMPLE.  The lookups of all the sample's features are walked as ONE
// flattened sequence l = 0 .. sum(H)-1 with HU row loads in flight at any time, whatever the hotness of the individual
// feature: in the ml_perf feature list (examples/ml_perf/configs/v6e_8.py:15-172) 14 of the 26 features are one-hot and one
// has 100 ids, and a group-per-(sample, feature) kernel keeps a single 512-byte load in flight for most of its items
// (measured 2.06 TB/s = 31 % of the HBM roofline).  Per feature the sum still runs h = 0 .. H-1 in order without FMA
// contraction, so results are bit-identical to gather_generic_kernel and to the oracle (embed_reduce.py:253-274).
// Requirements (host-checked): every dim a multiple of 4 and <= 128, 16-byte aligned tables / output, sum(H) <= SAMPLE_MAXL.
constexpr int SAMPLE_MAXL = 2048;
struct FeatSample {
  const float* table;
  const void* ids;
  const float* weights;      // nullptr unless honoured (embed_reduce.py:224)
  const float* const* shards;
  long long vocab, stride;
  int nshards, i64, nchunk, hot, out_off, div_kind;   // div_kind: 0 none, 1 mean, 2 sqrtn
};

// (A cp.async ring — 24 rows per warp in flight in shared memory — was tried and measured SLOWER, 4.31 ms against 2.45 ms
// for register-resident loads with the same per-lookup address work: the walk was instruction-bound, not latency-bound.)
// Per-lookup work is split like in gather_fast_kernel: ONE lane per lookup does the address work (id load, index rule,
// MOD shard split, row address, weight), the row pointer and weight reach the other lanes of the group by shuffle, and the
// per-lookup metadata every lane needs for the accumulate / finalize logic is one packed 8-byte shared-memory word.  (With
// every lane redoing the address math — ~50 instructions and a dozen shared-memory reads per lookup — the walk sat at
// 3.3-3.8 TB/s whatever the number of loads in flight.)
struct LookupMeta {          // per flattened lookup l, built once per CTA
  unsigned short feat;
  unsigned char nchunk;      // float4 columns of the feature
  unsigned char flags;       // 1: first lookup of its feature, 2: last, 4: hotness 1 (no add), 8: mean divisor, 16: sqrtn divisor
  int out_off;
};

template <int HU, int MINB>
__global__ void __launch_bounds__(256, MINB) gather_sample_kernel(const __grid_constant__ GatherParams p, int lpr, int total_hot) {
  __shared__ FeatSample sf[MAXF];
  __shared__ LookupMeta meta[SAMPLE_MAXL];
  __shared__ int first[MAXF + 1];
  for (int i = threadIdx.x; i < p.F; i += blockDim.x) {
    const krs_feature_t& f = p.f[i];
    FeatSample t;
    t.table = f.table;
    t.ids = f.ids;
    t.weights = (f.weights != nullptr && (f.reduce || f.combiner == KRS_COMBINER_SUM)) ? f.weights : nullptr;
    t.shards = f.shard_tables;
    t.vocab = f.vocab;
    t.stride = f.ids_stride;
    t.nshards = f.num_shards;
    t.i64 = f.ids_i64;
    t.nchunk = f.dim / 4;
    t.hot = f.hotness;
    t.out_off = f.out_offset;
    t.div_kind = (f.reduce && f.combiner != KRS_COMBINER_SUM) ? (f.combiner == KRS_COMBINER_MEAN ? 1 : 2) : 0;
    sf[i] = t;
  }
  if (threadIdx.x == 0) {
    int run = 0;
    for (int i = 0; i < p.F; ++i) { first[i] = run; run += p.f[i].hotness; }
    first[p.F] = run;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < p.F; i += blockDim.x) {
    const FeatSample& ft = sf[i];
    for (int l = first[i]; l < first[i + 1]; ++l) {
      LookupMeta m;
      m.feat = (unsigned short)i;
      m.nchunk = (unsigned char)ft.nchunk;
      m.flags = (unsigned char)((l == first[i] ? 1 : 0) | (l == first[i + 1] - 1 ? 2 : 0) | (ft.hot == 1 ? 4 : 0) |
                                (ft.div_kind == 1 ? 8 : 0) | (ft.div_kind == 2 ? 16 : 0));
      m.out_off = ft.out_off;
      meta[l] = m;
    }
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int sub = lane % lpr, gsub = lane / lpr;
  const int groups = 32 / lpr;
  const int gbase = gsub * lpr;                               // first lane of this group
  const int RES = lpr < HU ? lpr : HU;                        // lookups resolved (one per lane of the group) per pass
  const int64_t ngroups = (((int64_t)gridDim.x * blockDim.x) >> 5) * groups;
  // all lanes of a warp iterate together (shuffles below): the warp-level loop runs while ANY group has a sample
  for (int64_t b0 = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * groups; b0 < p.B; b0 += ngroups) {
    const int64_t b = b0 + gsub;
    const bool live = b < p.B;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float dsum = 0.f;
    for (int l0 = 0; l0 < total_hot; l0 += RES) {
      // ---- one lane per lookup: address work
      const float* my_src = nullptr;
      float my_w = 1.f;
      if (live && sub < RES && l0 + sub < total_hot) {
        const int l = l0 + sub;
        const int f = meta[l].feat;
        const FeatSample& ft = sf[f];
        const int64_t idx = b * ft.stride + (l - first[f]);
        const int64_t id = resolve_id(ft.i64 ? load_id<int64_t>(ft.ids, idx) : load_id<int32_t>(ft.ids, idx), ft.vocab);
        if (ft.weights) my_w = ft.weights[idx];
        if (id < 0) my_src = nan_row();
        else {
          const int dim = ft.nchunk * 4;
          my_src = ft.nshards > 1 ? ft.shards[(int)(id % ft.nshards)] + (id / ft.nshards) * (int64_t)dim : ft.table + id * (int64_t)dim;
        }
      }
      // ---- every lane: its float4 column of the RES rows, all loads issued before the first use
      const float* src[HU];
      float w[HU];
      float4 v[HU];
#pragma unroll
      for (int u = 0; u < HU; ++u) {
        src[u] = reinterpret_cast<const float*>(__shfl_sync(0xffffffffu, (unsigned long long)my_src, gbase + (u < RES ? u : 0)));
        w[u] = __shfl_sync(0xffffffffu, my_w, gbase + (u < RES ? u : 0));
        if (u >= RES || l0 + u >= total_hot) src[u] = nullptr;
      }
#pragma unroll
      for (int u = 0; u < HU; ++u) {
        if (src[u] == nullptr) continue;
        if (src[u] == nan_row()) v[u] = nan4();            // no such row: NaN (jnp.take "fill") poisons the reduced sample
        else if (sub < meta[l0 + u].nchunk) v[u] = ldg_nc_f4(src[u] + sub * 4);
        else v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < HU; ++u) {
        if (src[u] == nullptr) continue;
        const LookupMeta m = meta[l0 + u];
        if (m.flags & 1) {
          acc = make_float4(0.f, 0.f, 0.f, 0.f);
          dsum = 0.f;
        }
        // x = x * w ; sum over axis -2 in order h = 0..H-1 (no FMA contraction: embed_reduce.py:253,261)
        if (m.flags & 4) {
          acc.x = __fmul_rn(v[u].x, w[u]); acc.y = __fmul_rn(v[u].y, w[u]); acc.z = __fmul_rn(v[u].z, w[u]); acc.w = __fmul_rn(v[u].w, w[u]);
        } else {
          acc.x = __fadd_rn(acc.x, __fmul_rn(v[u].x, w[u])); acc.y = __fadd_rn(acc.y, __fmul_rn(v[u].y, w[u]));
          acc.z = __fadd_rn(acc.z, __fmul_rn(v[u].z, w[u])); acc.w = __fadd_rn(acc.w, __fmul_rn(v[u].w, w[u]));
        }
        if (m.flags & 8) dsum = __fadd_rn(dsum, w[u]);
        else if (m.flags & 16) dsum = __fadd_rn(dsum, __fmul_rn(w[u], w[u]));
        if ((m.flags & 2) && sub < m.nchunk) {
          float4 r = acc;
          if (m.flags & 24) {
            const float div = (m.flags & 8) ? dsum : sqrtf(dsum);
            r.x = (div != 0.f) ? __fdiv_rn(r.x, div) : 0.f;   // divide_no_nan
            r.y = (div != 0.f) ? __fdiv_rn(r.y, div) : 0.f;
            r.z = (div != 0.f) ? __fdiv_rn(r.z, div) : 0.f;
            r.w = (div != 0.f) ? __fdiv_rn(r.w, div) : 0.f;
          }
          stg_cs_f4(p.out + b * p.out_ld + m.out_off + sub * 4, r);
        }
      }
    }
  }
}

// ------------------------------------------------------------------ forward, generic path
// One lane group of `lpr` lanes per (b,f) item; element unit = float4 (VEC) or float.
template <bool VEC>
__global__ void __launch_bounds__(256) gather_generic_kernel(const __grid_constant__ GatherParams p, int lpr) {
  const int lane = threadIdx.x & 31;
  const int sub = lane % lpr, gsub = lane / lpr;
  const int groups = 32 / lpr;
  const int64_t items = p.B * p.F;
  const int64_t ngroups = (((int64_t)gridDim.x * blockDim.x) >> 5) * groups;
  const int64_t gid = ((((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * groups) + gsub;
  constexpr int W = VEC ? 4 : 1;
  for (int64_t it = gid; it < items; it += ngroups) {
    const int64_t b = it / p.F;
    const int f = (int)(it - b * p.F);
    const krs_feature_t& ft = p.f[f];
    const int H = ft.hotness;
    const int nchunk = ft.dim / W;
    const bool use_w = ft.weights != nullptr && (ft.reduce || ft.combiner == KRS_COMBINER_SUM);  // embed_reduce.py:224
    // divisor (mean: sum w ; sqrtn: sqrt(sum w^2)); every lane of the group computes it redundantly
    float div = 1.f;
    if (ft.reduce && ft.combiner != KRS_COMBINER_SUM) {
      float d = 0.f;
      for (int h = 0; h < H; ++h) {
        const float w = use_w ? ft.weights[b * ft.ids_stride + h] : 1.f;
        d = (ft.combiner == KRS_COMBINER_MEAN) ? __fadd_rn(d, w) : __fadd_rn(d, __fmul_rn(w, w));
      }
      div = (ft.combiner == KRS_COMBINER_MEAN) ? d : sqrtf(d);
    }
    float* dst = p.out + b * p.out_ld + ft.out_offset;
    for (int c = sub; c < nchunk; c += lpr) {
      float acc[W];
#pragma unroll
      for (int j = 0; j < W; ++j) acc[j] = 0.f;
      // Multi-hot rows (ml_perf hotness up to 100, examples/ml_perf/configs/v6e_8.py): HU row loads are in flight per lane
      // group before the first is consumed; the sum itself stays sequential in h (bit-exact against the oracle's order).
      constexpr int HU = 8;
      for (int h0 = 0; h0 < H; h0 += HU) {
        const float* src[HU];
        float w[HU];
        float v[HU][W];
#pragma unroll
        for (int u = 0; u < HU; ++u) {
          src[u] = nullptr;
          w[u] = 1.f;
          if (h0 + u < H) {
            const int64_t idx = b * ft.ids_stride + h0 + u;
            const int64_t id = resolve_id(ft.ids_i64 ? load_id<int64_t>(ft.ids, idx) : load_id<int32_t>(ft.ids, idx), ft.vocab);
            if (use_w) w[u] = ft.weights[idx];
            src[u] = id < 0 ? nan_row() : row_ptr(ft, id) + c * W;
          }
        }
#pragma unroll
        for (int u = 0; u < HU; ++u) {
          if (src[u] == nullptr) continue;
          if (src[u] == nan_row()) {               // no such row: NaN (jnp.take "fill"), which poisons the reduced sample
#pragma unroll
            for (int j = 0; j < W; ++j) v[u][j] = __int_as_float(0x7fc00000);
          } else if (VEC) {
            const float4 t = ldg_nc_f4(src[u]);
            v[u][0] = t.x; v[u][1] = t.y; v[u][2] = t.z; v[u][3] = t.w;
          } else {
            v[u][0] = __ldg(src[u]);
          }
        }
        // x = x * w ; sum over axis -2 in order h = 0..H-1 (no FMA contraction: embed_reduce.py:253,261)
#pragma unroll
        for (int u = 0; u < HU; ++u) {
          if (src[u] == nullptr) continue;
#pragma unroll
          for (int j = 0; j < W; ++j) acc[j] = (H == 1) ? __fmul_rn(v[u][j], w[u]) : __fadd_rn(acc[j], __fmul_rn(v[u][j], w[u]));
        }
      }
      if (ft.reduce && ft.combiner != KRS_COMBINER_SUM) {
#pragma unroll
        for (int j = 0; j < W; ++j) acc[j] = (div != 0.f) ? __fdiv_rn(acc[j], div) : 0.f;   // divide_no_nan
      }
      if (VEC) *reinterpret_cast<float4*>(dst + c * 4) = make_float4(acc[0], acc[1], acc[2], acc[3]);
      else dst[c] = acc[0];
    }
  }
}

// ------------------------------------------------------------------ backward
// Fast path: 1-hot, weight 1; warp = 32 consecutive samples of one feature.  Duplicate ids inside the warp
// are found with __match_any_sync; each distinct id ("leader") gets ONE accumulated row.
//   VEC variant (E = 4*LPR, LPR a power of two <= 32): LPR lanes own a row, RPW = 32/LPR leader rows are
//   processed per step, each lane sums its float4 over the duplicate samples and issues ONE 16-byte
//   RED.E.ADD.F32x4 — 4x fewer (and 4x wider) reductions than the scalar form, which matters most for
//   row-sharded tables where the reduction crosses NVLink to the owning GPU.
template <typename IdT, int LPR>
__global__ void __launch_bounds__(256) scatter_fast_kernel(const __grid_constant__ GatherParams p) {
  constexpr bool VEC = LPR > 0;
  constexpr int RPW = VEC ? 32 / (LPR > 0 ? LPR : 1) : 1;
  const int lane = threadIdx.x & 31;
  const int64_t nblk = (p.B + 31) >> 5;                 // sample blocks
  const int64_t items = nblk * p.F;                     // item = (sample block, feature), feature fastest
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const float* __restrict__ gout = p.out;
  for (int64_t it = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; it < items; it += nwarps) {
    const int64_t blk = it / p.F;
    const int f = (int)(it - blk * p.F);
    const krs_feature_t& ft = p.f[f];
    const int64_t b0 = blk << 5;
    const int64_t b = b0 + lane;
    bool valid = b < p.B;
    int64_t id = -1 - lane;                              // unique negative sentinel for tail lanes
    if (valid) {
      id = resolve_id(load_id<IdT>(ft.ids, b * ft.ids_stride), ft.vocab);
      if (id < 0) { valid = false; id = -1 - lane; }        // ids that address no row receive no gradient
    }
    const int S = ft.num_shards;
    const unsigned peers = __match_any_sync(0xffffffffu, id);
    const bool leader = valid && ((__ffs(peers) - 1) == lane);
    if (leader) {
      if (S > 1) {
        if (ft.shard_touched) {
          const int64_t lr = id / S;
          atomicOr(ft.shard_touched[(int)(id % S)] + (lr >> 5), 1u << (lr & 31));
        }
      } else if (ft.touched) {
        atomicOr(ft.touched + (id >> 5), 1u << (id & 31));
      }
    }
    const unsigned leaders = __ballot_sync(0xffffffffu, leader);
    const int E = ft.dim;
    if (VEC) {
      const int sub = lane % (LPR > 0 ? LPR : 1), rsub = lane / (LPR > 0 ? LPR : 1);
      const int nlead = __popc(leaders);
      // U leaders per lane group and pass: their first gradient rows are in flight together (one load per lane and pass
      // left the kernel latency-bound at 52 % of the HBM roofline); duplicates of a row, which are rare, follow serially.
      constexpr int U = 4;
      for (int base = 0; base < nlead; base += RPW * U) {
        float4 acc[U];
        int r[U];
        int64_t rid[U];
        unsigned rp[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int nth = base + u * RPW + rsub;                 // this lane group's leader (nth set bit)
          const bool act = nth < nlead;
          r[u] = act ? (int)__fns(leaders, 0, nth + 1) : 0;
          rid[u] = __shfl_sync(0xffffffffu, id, r[u]);
          rp[u] = __shfl_sync(0xffffffffu, peers, r[u]);
          if (!act) r[u] = -1;
          else acc[u] = ldg_nc_f4(gout + (b0 + r[u]) * p.out_ld + ft.out_offset + sub * 4);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (r[u] < 0) continue;
          for (unsigned q = rp[u] & ~(1u << r[u]); q; q &= q - 1) {
            const int j = __ffs(q) - 1;
            const float4 v = ldg_nc_f4(gout + (b0 + j) * p.out_ld + ft.out_offset + sub * 4);
            acc[u].x += v.x; acc[u].y += v.y; acc[u].z += v.z; acc[u].w += v.w;
          }
          float* drow = (S > 1) ? ft.shard_grads[(int)(rid[u] % S)] + (rid[u] / S) * (int64_t)E : ft.grad + rid[u] * (int64_t)E;
          atomicAdd(reinterpret_cast<float4*>(drow + sub * 4), acc[u]);
        }
      }
    } else {
      for (unsigned m = leaders; m; m &= m - 1) {
        const int r = __ffs(m) - 1;
        const int64_t rid = __shfl_sync(0xffffffffu, id, r);
        const unsigned rp = __shfl_sync(0xffffffffu, peers, r);
        float* drow = (S > 1) ? ft.shard_grads[(int)(rid % S)] + (rid / S) * (int64_t)E : ft.grad + rid * (int64_t)E;
        for (int c = lane; c < E; c += 32) {
          float acc = 0.f;
          for (unsigned q = rp; q; q &= q - 1) {
            const int j = __ffs(q) - 1;
            acc += gout[(b0 + j) * p.out_ld + ft.out_offset + c];
          }
          atomicAdd(drow + c, acc);
        }
      }
    }
  }
}


// Generic: warp per (b,f) item, loops over hotness; coefficient = w / divisor.
__global__ void __launch_bounds__(256) scatter_generic_kernel(const __grid_constant__ GatherParams p, bool vec) {
  const int lane = threadIdx.x & 31;
  const int64_t items = p.B * p.F;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const float* __restrict__ gout = p.out;
  for (int64_t it = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; it < items; it += nwarps) {
    const int64_t b = it / p.F;
    const int f = (int)(it - b * p.F);
    const krs_feature_t& ft = p.f[f];
    const int H = ft.hotness, E = ft.dim;
    const bool use_w = ft.weights != nullptr && (ft.reduce || ft.combiner == KRS_COMBINER_SUM);
    float scale = 1.f;
    if (ft.reduce && ft.combiner != KRS_COMBINER_SUM) {
      float d = 0.f;
      for (int h = 0; h < H; ++h) {
        const float w = use_w ? ft.weights[b * ft.ids_stride + h] : 1.f;
        d += (ft.combiner == KRS_COMBINER_MEAN) ? w : w * w;
      }
      if (ft.combiner == KRS_COMBINER_SQRTN) d = sqrtf(d);
      scale = d != 0.f ? 1.f / d : 0.f;
    }
    const float* g = gout + b * p.out_ld + ft.out_offset;
    for (int h = 0; h < H; ++h) {
      const int64_t idx = b * ft.ids_stride + h;
      const int64_t id = resolve_id(ft.ids_i64 ? load_id<int64_t>(ft.ids, idx) : load_id<int32_t>(ft.ids, idx), ft.vocab);
      const float coef = (use_w ? ft.weights[idx] : 1.f) * scale;
      if (coef == 0.f || id < 0) continue;
      float* drow;
      if (ft.num_shards > 1) drow = ft.shard_grads[(int)(id % ft.num_shards)] + (id / ft.num_shards) * (int64_t)E;
      else drow = ft.grad + id * (int64_t)E;
      if (lane == 0) {
        if (ft.num_shards > 1) {
          if (ft.shard_touched) {
            const int64_t lr = id / ft.num_shards;
            atomicOr(ft.shard_touched[(int)(id % ft.num_shards)] + (lr >> 5), 1u << (lr & 31));
          }
        } else if (ft.touched) {
          atomicOr(ft.touched + (id >> 5), 1u << (id & 31));
        }
      }
      if (vec) {                                         // one 16-byte reduction per lane instead of four scalar ones
        for (int c = lane * 4; c < E; c += 128) {
          const float4 gv = *reinterpret_cast<const float4*>(g + c);
          atomicAdd(reinterpret_cast<float4*>(drow + c), make_float4(coef * gv.x, coef * gv.y, coef * gv.z, coef * gv.w));
        }
      } else {
        for (int c = lane; c < E; c += 32) atomicAdd(drow + c, coef * g[c]);
      }
    }
  }
}

// Multi-hot backward, the mirror of gather_sample_kernel: lane group per SAMPLE, the sample's lookups walked as one flattened
// sequence, ONE lane per lookup does the address work (id, index rule, coefficient w / divisor, destination row, touched
// bit), pointer + coefficient reach the group by shuffle, every lane issues one 16-byte reduction per lookup; the sample's
// gradient chunk of a feature is loaded once and reused for all its lookups.  (The warp-per-(sample, feature) kernel above
// keeps one row in flight per warp for the one-hot majority of the ml_perf list and repeats the address math in 32 lanes.)
struct FeatScatter {
  float* grad;
  float* const* shard_grads;
  uint32_t* touched;
  uint32_t* const* shard_touched;
  const void* ids;
  const float* weights;      // nullptr unless honoured
  long long vocab, stride;
  int nshards, i64, nchunk, hot, out_off, div_kind;
};

#ifndef KRS_SCAT_HU
#define KRS_SCAT_HU 4
#endif
#ifndef KRS_SCAT_MINB
#define KRS_SCAT_MINB 4
#endif
template <int HU, int MINB>
__global__ void __launch_bounds__(256, MINB) scatter_sample_kernel(const __grid_constant__ GatherParams p, int lpr, int total_hot) {
  __shared__ FeatScatter sf[MAXF];
  __shared__ LookupMeta meta[SAMPLE_MAXL];
  __shared__ int first[MAXF + 1];
  for (int i = threadIdx.x; i < p.F; i += blockDim.x) {
    const krs_feature_t& f = p.f[i];
    FeatScatter t;
    t.grad = f.grad;
    t.shard_grads = f.shard_grads;
    t.touched = f.touched;
    t.shard_touched = f.shard_touched;
    t.ids = f.ids;
    t.weights = (f.weights != nullptr && (f.reduce || f.combiner == KRS_COMBINER_SUM)) ? f.weights : nullptr;
    t.vocab = f.vocab;
    t.stride = f.ids_stride;
    t.nshards = f.num_shards;
    t.i64 = f.ids_i64;
    t.nchunk = f.dim / 4;
    t.hot = f.hotness;
    t.out_off = f.out_offset;
    t.div_kind = (f.reduce && f.combiner != KRS_COMBINER_SUM) ? (f.combiner == KRS_COMBINER_MEAN ? 1 : 2) : 0;
    sf[i] = t;
  }
  if (threadIdx.x == 0) {
    int run = 0;
    for (int i = 0; i < p.F; ++i) { first[i] = run; run += p.f[i].hotness; }
    first[p.F] = run;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < p.F; i += blockDim.x) {
    const FeatScatter& ft = sf[i];
    for (int l = first[i]; l < first[i + 1]; ++l) {
      LookupMeta m;
      m.feat = (unsigned short)i;
      m.nchunk = (unsigned char)ft.nchunk;
      m.flags = (unsigned char)(l == first[i] ? 1 : 0);
      m.out_off = ft.out_off;
      meta[l] = m;
    }
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int sub = lane % lpr, gsub = lane / lpr;
  const int groups = 32 / lpr;
  const int gbase = gsub * lpr;
  const int RES = lpr < HU ? lpr : HU;
  const float* __restrict__ gout = p.out;
  const int64_t ngroups = (((int64_t)gridDim.x * blockDim.x) >> 5) * groups;
  for (int64_t b0 = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * groups; b0 < p.B; b0 += ngroups) {
    const int64_t b = b0 + gsub;
    const bool live = b < p.B;
    float4 gcur = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int l0 = 0; l0 < total_hot; l0 += RES) {
      float* my_dst = nullptr;
      float my_coef = 0.f;
      if (live && sub < RES && l0 + sub < total_hot) {
        const int l = l0 + sub;
        const int f = meta[l].feat;
        const FeatScatter& ft = sf[f];
        const int64_t base = b * ft.stride;
        const int64_t idx = base + (l - first[f]);
        const int64_t id = resolve_id(ft.i64 ? load_id<int64_t>(ft.ids, idx) : load_id<int32_t>(ft.ids, idx), ft.vocab);
        float coef = ft.weights ? ft.weights[idx] : 1.f;
        if (ft.div_kind) {                                  // mean / sqrtn: the feature's divisor (rare on this path)
          float d = 0.f;
          for (int h = 0; h < ft.hot; ++h) {
            const float w = ft.weights ? ft.weights[base + h] : 1.f;
            d += ft.div_kind == 1 ? w : w * w;
          }
          if (ft.div_kind == 2) d = sqrtf(d);
          coef *= d != 0.f ? 1.f / d : 0.f;
        }
        if (id >= 0 && coef != 0.f) {                       // ids that address no row receive no gradient
          const int dim = ft.nchunk * 4;
          if (ft.nshards > 1) {
            const int o = (int)(id % ft.nshards);
            const int64_t lr = id / ft.nshards;
            my_dst = ft.shard_grads[o] + lr * (int64_t)dim;
            if (ft.shard_touched) atomicOr(ft.shard_touched[o] + (lr >> 5), 1u << (lr & 31));
          } else {
            my_dst = ft.grad + id * (int64_t)dim;
            if (ft.touched) atomicOr(ft.touched + (id >> 5), 1u << (id & 31));
          }
          my_coef = coef;
        }
      }
#pragma unroll
      for (int u = 0; u < HU; ++u) {
        float* dst = reinterpret_cast<float*>(__shfl_sync(0xffffffffu, (unsigned long long)my_dst, gbase + (u < RES ? u : 0)));
        const float coef = __shfl_sync(0xffffffffu, my_coef, gbase + (u < RES ? u : 0));
        if (u >= RES || l0 + u >= total_hot || !live) continue;
        const LookupMeta m = meta[l0 + u];
        if (sub >= m.nchunk) continue;
        if (m.flags & 1) gcur = ldg_nc_f4(gout + b * p.out_ld + m.out_off + sub * 4);     // the feature's gradient chunk, once
        if (dst != nullptr)
          atomicAdd(reinterpret_cast<float4*>(dst + sub * 4), make_float4(coef * gcur.x, coef * gcur.y, coef * gcur.z, coef * gcur.w));
      }
    }
  }
}

int fill_params(GatherParams& p, const krs_feature_t* features, int F, int64_t B, float* out, int64_t out_ld) {
  KRS_REQUIRE(features != nullptr && F > 0 && F <= MAXF, "gather: need 1..%d features per call, got %d", MAXF, F);
  KRS_REQUIRE(B >= 0, "gather: negative batch");
  for (int i = 0; i < F; ++i) {
    const krs_feature_t& f = features[i];
    KRS_REQUIRE(f.ids != nullptr || B == 0, "gather: feature %d has null ids", i);
    KRS_REQUIRE(f.dim > 0 && f.hotness > 0 && f.vocab > 0, "gather: feature %d has bad dim/hotness/vocab", i);
    KRS_REQUIRE(f.combiner >= 0 && f.combiner <= 2, "gather: feature %d has unknown combiner %d", i, f.combiner);
    KRS_REQUIRE(f.out_offset >= 0 && f.out_offset + f.dim <= out_ld, "gather: feature %d columns exceed out_ld", i);
    KRS_REQUIRE(f.num_shards <= 1 || f.shard_tables != nullptr || f.shard_grads != nullptr,
                "gather: feature %d is sharded but has no shard pointer array", i);
    KRS_REQUIRE(f.num_shards <= 1 || (f.shard_mode & 0xff) == KRS_SHARD_DIRECT,
                "gather: feature %d: only KRS_SHARD_DIRECT addressing exists here (the routed exchange is krs_xchg_*)", i);
    p.f[i] = f;
  }
  p.F = F;
  p.B = B;
  p.out = out;
  p.out_ld = out_ld;
  return KRS_OK;
}

bool uniform_onehot(const GatherParams& p, int* E_out, bool* i64, bool* sharded) {
  const int E = p.f[0].dim;
  const int is64 = p.f[0].ids_i64;
  const bool sh = p.f[0].num_shards > 1;
  for (int i = 0; i < p.F; ++i) {
    const krs_feature_t& f = p.f[i];
    if (f.hotness != 1 || f.dim != E || f.ids_i64 != is64 || ((f.num_shards > 1) != sh)) return false;
    if (f.weights != nullptr && (f.reduce || f.combiner == KRS_COMBINER_SUM)) return false;
    if (f.out_offset != i * E) return false;
    if (!sh && (f.table == nullptr || !aligned16(f.table))) return false;
  }
  *sharded = sh;
  if (p.out_ld != (int64_t)p.F * E || !aligned16(p.out)) return false;
  *E_out = E;
  *i64 = is64 != 0;
  return true;
}

template <int LPR, typename IdT, bool SHARDED>
int launch_fast(const GatherParams& p, cudaStream_t s) {
  const int64_t R = p.B * p.F;
  const int64_t chunks = ceil_div<int64_t>(R, 32);
  // two 32-row chunks per warp keeps ~7 waves of CTAs at C2 (small tail) while amortising the
  // per-CTA descriptor staging
  const int chunks_per_warp = chunks >= (int64_t)sm_count() * 8 * 6 * 4 ? 2 : 1;
  const int64_t blocks_needed = ceil_div<int64_t>(chunks, 8 * chunks_per_warp);
  const unsigned grid = (unsigned)krs::imax<int64_t>(1, krs::imin<int64_t>(blocks_needed, 0x7fffffff));
  gather_fast_kernel<LPR, IdT, SHARDED><<<grid, 256, 0, s>>>(p, chunks_per_warp);
  KRS_LAUNCH_CHECK();
  return KRS_OK;
}

template <typename IdT>
int launch_bulk(const GatherParams& p, int E, cudaStream_t s) {
  // two tiles of ~48 KB each
  int tile_rows = (48 * 1024) / (E * 4);
  if (tile_rows < 1) return KRS_EUNSUPPORTED;
  const size_t smem = (size_t)2 * tile_rows * E * 4;
  // per device and cheap: set before every launch (a process may drive several GPUs)
  KRS_CUDA(cudaFuncSetAttribute(gather_bulk_kernel<IdT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const int64_t R = p.B * p.F;
  const int64_t ntiles = ceil_div<int64_t>(R, tile_rows);
  const unsigned grid = (unsigned)krs::imax<int64_t>(1, krs::imin<int64_t>(ntiles, (int64_t)sm_count() * 2));
  gather_bulk_kernel<IdT><<<grid, 256, smem, s>>>(p, E, tile_rows);
  KRS_LAUNCH_CHECK();
  return KRS_OK;
}

}  // namespace
}  // namespace krs

using namespace krs;

extern "C" int krs_gather_fwd(const krs_feature_t* features, int F, int64_t B, float* out, int64_t out_ld, int variant,
                              void* stream) {
  GatherParams p;
  int rc = fill_params(p, features, F, B, out, out_ld);
  if (rc) return rc;
  KRS_REQUIRE(out != nullptr || B == 0, "krs_gather_fwd: null output");
  for (int i = 0; i < F; ++i) {
    KRS_REQUIRE(p.f[i].num_shards <= 1 ? p.f[i].table != nullptr : p.f[i].shard_tables != nullptr,
                "krs_gather_fwd: feature %d has no table", i);
  }
  if (B == 0) return KRS_OK;
  cudaStream_t s = as_stream(stream);
  int E = 0;
  bool i64 = false, sharded = false;
  const bool fast_ok = uniform_onehot(p, &E, &i64, &sharded) && (E % 4 == 0) && ((E / 4) & (E / 4 - 1)) == 0 && E <= 128;
  KRS_REQUIRE(variant >= 0 && variant <= 3, "krs_gather_fwd: unknown variant %d", variant);
  if ((variant == 1 || variant == 2) && !fast_ok) {
    set_error("krs_gather_fwd: variant %d needs 1-hot features of one dim E in {4,8,16,32,64,128}", variant);
    return KRS_EUNSUPPORTED;
  }
  if (variant == 2) {
    if (sharded) {
      set_error("krs_gather_fwd: the bulk-copy variant does not address sharded tables");
      return KRS_EUNSUPPORTED;
    }
    return i64 ? launch_bulk<int64_t>(p, E, s) : launch_bulk<int32_t>(p, E, s);
  }
  if (fast_ok && variant != 3) {
#define KRS_FAST(L)                                                              \
  case L:                                                                        \
    if (sharded) return i64 ? launch_fast<L, int64_t, true>(p, s) : launch_fast<L, int32_t, true>(p, s); \
    return i64 ? launch_fast<L, int64_t, false>(p, s) : launch_fast<L, int32_t, false>(p, s);
    switch (E / 4) {
      KRS_FAST(1)
      KRS_FAST(2)
      KRS_FAST(4)
      KRS_FAST(8)
      KRS_FAST(16)
      KRS_FAST(32)
    }
#undef KRS_FAST
  }
  // generic
  int maxE = 0;
  bool vec = aligned16(out) && (out_ld % 4 == 0);
  for (int i = 0; i < F; ++i) {
    maxE = max(maxE, p.f[i].dim);
    if (p.f[i].dim % 4 != 0 || p.f[i].out_offset % 4 != 0) vec = false;
    if (p.f[i].num_shards <= 1 && !aligned16(p.f[i].table)) vec = false;
  }
  const int chunks = vec ? maxE / 4 : maxE;
  int lpr = 1;
  while (lpr < chunks && lpr < 32) lpr <<= 1;
  int64_t total_hot = 0;
  for (int i = 0; i < F; ++i) total_hot += p.f[i].hotness;
  if (vec && variant != 3 && maxE <= 128 && total_hot <= SAMPLE_MAXL && F <= 255) {
    // flattened per-sample walk: a constant number of row loads in flight whatever the per-feature hotness
    const int64_t warps_needed = ceil_div<int64_t>(B, 32 / lpr);
    const unsigned grid = (unsigned)krs::imax<int64_t>(1, krs::imin<int64_t>(ceil_div<int64_t>(warps_needed, 8), (int64_t)sm_count() * 64));
    // 4 lookups per pass, <= 40 registers, 6 CTAs per SM: the best of the measured variants (1.58 ms = 5.16 TB/s at the
    // ml_perf shape; 8 lookups per pass at 3 or 4 CTAs per SM: 1.63 / 1.68 ms — profiles/r2_multihot_gather_variants.txt)
    gather_sample_kernel<4, 6><<<grid, 256, 0, s>>>(p, lpr, (int)total_hot);
    KRS_LAUNCH_CHECK();
    return KRS_OK;
  }
  const int64_t items = B * F;
  const int64_t warps_needed = ceil_div<int64_t>(items, 32 / lpr);
  const unsigned grid = (unsigned)krs::imax<int64_t>(1, krs::imin<int64_t>(ceil_div<int64_t>(warps_needed, 8), (int64_t)sm_count() * 32));
  if (vec) gather_generic_kernel<true><<<grid, 256, 0, s>>>(p, lpr);
  else gather_generic_kernel<false><<<grid, 256, 0, s>>>(p, lpr);
  KRS_LAUNCH_CHECK();
  return KRS_OK;
}

extern "C" int krs_gather_bwd(const krs_feature_t* features, int F, int64_t B, const float* gout, int64_t gout_ld,
                              void* stream) {
  GatherParams p;
  int rc = fill_params(p, features, F, B, const_cast<float*>(gout), gout_ld);
  if (rc) return rc;
  KRS_REQUIRE(gout != nullptr, "krs_gather_bwd: null gradient");
  bool fast = true;
  const int is64 = p.f[0].ids_i64;
  for (int i = 0; i < F; ++i) {
    const krs_feature_t& f = p.f[i];
    KRS_REQUIRE(f.num_shards > 1 ? f.shard_grads != nullptr : f.grad != nullptr,
                "krs_gather_bwd: feature %d has no gradient arena", i);
    if (f.hotness != 1 || f.ids_i64 != is64) fast = false;
    if (f.weights != nullptr && (f.reduce || f.combiner == KRS_COMBINER_SUM)) fast = false;
  }
  if (B == 0) return KRS_OK;
  cudaStream_t s = as_stream(stream);
  if (fast) {
    const int64_t items = ceil_div<int64_t>(B, 32) * F;
    const unsigned grid = (unsigned)krs::imax<int64_t>(1, krs::imin<int64_t>(ceil_div<int64_t>(items, 8), 0x7fffffff));
    // vector variant: every feature has the same E = 4*LPR (LPR in {1,2,4,8,16,32}) and 16-byte aligned rows
    int E0 = p.f[0].dim;
    bool vec = (E0 % 4 == 0) && ((E0 / 4) & (E0 / 4 - 1)) == 0 && E0 <= 128 && aligned16(gout) && (gout_ld % 4 == 0);
    for (int i = 0; i < F && vec; ++i) {
      const krs_feature_t& f = p.f[i];
      if (f.dim != E0 || (f.out_offset % 4) != 0) vec = false;
      if (f.num_shards <= 1 && !aligned16(f.grad)) vec = false;
    }
#define KRS_SCAT(L)                                                                                   \
  case L:                                                                                             \
    if (is64) scatter_fast_kernel<int64_t, L><<<grid, 256, 0, s>>>(p);                                \
    else scatter_fast_kernel<int32_t, L><<<grid, 256, 0, s>>>(p);                                     \
    break;
    switch (vec ? E0 / 4 : 0) {
      KRS_SCAT(1) KRS_SCAT(2) KRS_SCAT(4) KRS_SCAT(8) KRS_SCAT(16) KRS_SCAT(32)
      default:
        if (is64) scatter_fast_kernel<int64_t, 0><<<grid, 256, 0, s>>>(p);
        else scatter_fast_kernel<int32_t, 0><<<grid, 256, 0, s>>>(p);
    }
#undef KRS_SCAT
  } else {
    const int64_t items = B * F;
    const unsigned grid = (unsigned)krs::imax<int64_t>(1, krs::imin<int64_t>(ceil_div<int64_t>(items, 8), (int64_t)sm_count() * 32));
    bool vec = aligned16(gout) && (gout_ld % 4 == 0);
    for (int i = 0; i < F && vec; ++i) {
      const krs_feature_t& f = p.f[i];
      if (f.dim % 4 != 0 || f.out_offset % 4 != 0 || (f.num_shards <= 1 && !aligned16(f.grad))) vec = false;
    }
    int maxE = 0;
    int64_t total_hot = 0;
    for (int i = 0; i < F; ++i) { maxE = max(maxE, p.f[i].dim); total_hot += p.f[i].hotness; }
    if (vec && maxE <= 128 && total_hot <= SAMPLE_MAXL) {
      int lpr = 1;
      while (lpr < maxE / 4 && lpr < 32) lpr <<= 1;
      const int64_t warps_needed = ceil_div<int64_t>(B, 32 / lpr);
      #ifndef KRS_SCAT_GRIDMUL
#define KRS_SCAT_GRIDMUL 64
#endif
      const unsigned g2 = (unsigned)krs::imax<int64_t>(1, krs::imin<int64_t>(ceil_div<int64_t>(warps_needed, 8), (int64_t)sm_count() * KRS_SCAT_GRIDMUL));
      scatter_sample_kernel<KRS_SCAT_HU, KRS_SCAT_MINB><<<g2, 256, 0, s>>>(p, lpr, (int)total_hot);
    } else {
      scatter_generic_kernel<<<grid, 256, 0, s>>>(p, vec);
    }
  }
  KRS_LAUNCH_CHECK();
  return KRS_OK;
}
