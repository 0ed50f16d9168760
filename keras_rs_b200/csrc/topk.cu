// topk.cu — BruteForceRetrieval: streaming score + exact top-k.
//
// Replaces Retrieval.compute_score (retrieval.py:101-117: matmul(q, transpose(c))) followed by
// keras.ops.top_k and the optional ops.take(candidate_ids, top_ids) (brute_force_retrieval.py:139-143).
// The (nq, nc) score matrix (164 GB at C4) is never materialised.  Two score engines share the selection rule and the
// merge kernel:
//   * topk_tc_kernel (tensor pipe, the default for large problems; see the block comment above it): tcgen05 3xTF32 score
//     tiles with the 128-query tile parked in tensor memory, 41 ms at C4;
//   * topk_partial_kernel (exact-fp32 FMA; small problems and shapes the tensor-pipe kernel does not take: d % 4 != 0,
//     d > 64): each CTA owns a 64-query tile and a slice of the candidates, computes 64x128 score tiles in registers,
//     filters them against the running k-th best score of each query and merges the few survivors into per-query
//     sorted lists held in shared memory, 257 ms at C4.
// topk_merge_kernel then merges the per-slice lists of every query.
// Ordering is total and deterministic: score descending, ties -> lowest candidate index first
// (jax.lax.top_k rule adopted in SURVEY.md Appendix A.3), independent of thread scheduling.
#include <limits.h>
#include <math.h>

#include <atomic>

#include "common.cuh"
#include "tc_common.cuh"

namespace krs {
namespace {
using namespace tcx;

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


// =====================================================================================================================
// Tensor-pipe variant (tcgen05): the score tiles Q C^T run as 3xTF32 UMMAs (fp32-level accuracy, like gemm_tc.cu).
//
// Work item = (candidate slice s, 128-query tile qt), items ordered slice-major so that the CTAs running at the same
// time stream the same slice (one DRAM read, the other query tiles hit L2).  Per item:
//   * the query tile is loaded ONCE, split into hi / lo in registers and parked in TENSOR MEMORY (columns 384..511:
//     k-block kb -> 16 hi + 16 lo columns), where it serves as the A operand of every MMA of the item (TS form);
//   * candidate tiles (96 candidates x 16 dims per ring entry) arrive by TMA, converter threads add the lo plane
//     right behind the raw tile, and per k-step two MMAs produce  Q_hi x [C_hi | C_lo] -> [main | cross]  (N = 192)
//     and  Q_lo x C_hi -> cross  (N = 96) into a double-buffered pair of TMEM accumulators;
//   * four selection warps (lane = query = TMEM lane) read the 96 scores of a tile, compare them with the running
//     k-th best of their query and insert the rare survivors, warp-cooperatively, into per-query sorted lists in shared
//     memory (score desc, index asc — the same total order as the FFMA kernel);
//   * at the end of the item the lists go to the workspace and topk_merge_kernel merges the S lists per query.
// Roles (512 threads): warps 0-3 selection | warps 4-11 converters (two groups, alternate ring entries) |
// warp 12 TMA producer | warp 13 MMA issuer (warp-uniform issue, elect.sync).
constexpr int TQ = 128;                 // queries per item == TMEM lanes
constexpr int TNC = 96;                 // candidates per tile
constexpr int TKB = 16;                 // dims per ring entry (64-byte rows, SWIZZLE_64B)
constexpr int T_MAXKB = 4;              // d <= 64
constexpr int T_STAGE = 12288;          // ring entry: query k-block 8 KB | candidate raw 6 KB + lo 6 KB
constexpr int T_QBYTES = TQ * TKB * 4;  // 8192
constexpr int T_CBYTES = TNC * TKB * 4; // 6144
constexpr int T_QCOL0 = 4 * TNC;        // TMEM columns of the parked query tile
constexpr int T_THREADS = 512;
constexpr int T_MAX_STAGES = 16;

// order-preserving float <-> int key (involution), so that atomicMax on ints orders scores
__device__ __forceinline__ int score_key(float x) {
  const int b = __float_as_int(x);
  return b >= 0 ? b : (b ^ 0x7fffffff);
}
__device__ __forceinline__ float key_score(int key) { return __int_as_float(key >= 0 ? key : (key ^ 0x7fffffff)); }

struct TopkTcArgs {
  float* part_s;
  int32_t* part_i;
  int* g_bound;          // [nq] best k-th score any slice has reached so far (score_key), a lower bound of the final k-th
  int64_t nq, nc;
  int d, k, S, KB, stages;
  int q_tiles;
  int64_t tiles_per_slice, ntiles, n_items;
};

__global__ void __launch_bounds__(T_THREADS, 1)
topk_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_c, const TopkTcArgs a) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int STAGES = a.stages;
  const int k = a.k;
  long long* list_k = reinterpret_cast<long long*>(smem + (size_t)STAGES * T_STAGE);   // [TQ][k] packed (score, index) keys
  uint64_t* bars = reinterpret_cast<uint64_t*>(list_k + (size_t)TQ * k);
  uint64_t* full_bar = bars;                       // [STAGES] TMA landed
  uint64_t* conv_bar = bars + T_MAX_STAGES;        // [STAGES] converter group done
  uint64_t* empty_bar = bars + 2 * T_MAX_STAGES;   // [STAGES] MMAs that read the entry have completed
  uint64_t* tmem_full = bars + 3 * T_MAX_STAGES;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;            // [2]
  uint64_t* qfree_bar = tmem_empty + 2;            // [1] all MMAs of the item have completed (query columns reusable)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(qfree_bar + 1);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&conv_bar[s], 4);
      mbar_init(&empty_bar[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 4);
    }
    mbar_init(qfree_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 13) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);
  const int KB = a.KB;

  if (warp == 12) {
    // ======================= TMA producer =======================
    int stage = 0;
    uint32_t phase = 0;
    for (int64_t item = blockIdx.x; item < a.n_items; item += gridDim.x) {
      const int64_t sl = item / a.q_tiles;
      const int qt = (int)(item - sl * a.q_tiles);
      const int64_t t0 = sl * a.tiles_per_slice;
      const int64_t t1 = imin<int64_t>(a.ntiles, t0 + a.tiles_per_slice);
      for (int kb = 0; kb < KB; ++kb) {
        mbar_wait_uniform(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&full_bar[stage], (uint32_t)T_QBYTES);
          tma_load_2d(smem + (size_t)stage * T_STAGE, &tmap_q, kb * TKB, qt * TQ, &full_bar[stage]);
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      for (int64_t tile = t0; tile < t1; ++tile)
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait_uniform(&empty_bar[stage], phase ^ 1);
          if (elect_one()) {
            mbar_expect_tx(&full_bar[stage], (uint32_t)T_CBYTES);
            tma_load_2d(smem + (size_t)stage * T_STAGE, &tmap_c, kb * TKB, (int)(tile * TNC), &full_bar[stage]);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
    }
  } else if (warp == 13) {
    // ======================= MMA issuer =======================
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TNC >> 3) << 17) | ((uint32_t)(TQ >> 4) << 24);
    const uint32_t idesc2 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)((2 * TNC) >> 3) << 17) | ((uint32_t)(TQ >> 4) << 24);
    const uint64_t dB = desc_kmajor(smem_u32(smem), 0);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int64_t item = blockIdx.x; item < a.n_items; item += gridDim.x) {
      const int64_t sl = item / a.q_tiles;
      const int64_t t0 = sl * a.tiles_per_slice;
      const int64_t t1 = imin<int64_t>(a.ntiles, t0 + a.tiles_per_slice);
      // query entries: nothing to multiply, the entry is released as soon as the converters have parked it in TMEM
      for (int kb = 0; kb < KB; ++kb) {
        mbar_wait_uniform(&conv_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) tc_commit(&empty_bar[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      for (int64_t tile = t0; tile < t1; ++tile) {
        mbar_wait_uniform(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_main = tmem_base + (uint32_t)(acc * 2 * TNC);
        const uint32_t d_cross = d_main + (uint32_t)TNC;
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait_uniform(&conv_bar[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t so = (uint32_t)(stage * T_STAGE) >> 4;
            const uint32_t q_hi = tmem_base + (uint32_t)(T_QCOL0 + kb * 32);
            const uint32_t q_lo = q_hi + 16;
#pragma unroll
            for (int ks = 0; ks < TKB / 8; ++ks) {
              const uint64_t db = dB + so + (ks ? 2u : 0u);          // k-step = +32 bytes
              const uint32_t first = (kb == 0 && ks == 0) ? 0u : 1u;
              tc_mma_tf32_ts(d_main, q_hi + 8 * ks, db, idesc2, first);    // [hi*hi | hi*lo] -> [main | cross]
              tc_mma_tf32_ts(d_cross, q_lo + 8 * ks, db, idesc, 1u);       // lo*hi -> cross
            }
            tc_commit(&empty_bar[stage]);
            if (kb == KB - 1) tc_commit(&tmem_full[acc]);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      if (elect_one()) tc_commit(qfree_bar);      // completes when every MMA of this item has retired
    }
  } else if (warp >= 4 && warp < 12) {
    // ======================= converters =======================
    const int grp = (warp - 4) >> 2;
    const int q = warp & 3;                 // TMEM lane quadrant of this warp
    const int row = q * 32 + lane;
    const int gt = (int)threadIdx.x - 128 - grp * 128;
    const uint32_t a_row_off = (uint32_t)row * 64u;
    const uint32_t a_sw = (uint32_t)(row >> 1) & 3u;
    int stage = 0;
    uint32_t phase = 0;
    uint32_t cnt = 0;
    uint32_t items_done = 0;
    for (int64_t item = blockIdx.x; item < a.n_items; item += gridDim.x) {
      const int64_t sl = item / a.q_tiles;
      const int64_t t0 = sl * a.tiles_per_slice;
      const int64_t t1 = imin<int64_t>(a.ntiles, t0 + a.tiles_per_slice);
      for (int kb = 0; kb < KB; ++kb) {
        if ((int)(cnt & 1u) == grp) {
          mbar_wait(&full_bar[stage], phase);
          const unsigned char* st = smem + (size_t)stage * T_STAGE;
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const uint4 v = *reinterpret_cast<const uint4*>(st + a_row_off + ((((uint32_t)c) ^ a_sw) << 4));
            hi[4 * c + 0] = v.x; hi[4 * c + 1] = v.y; hi[4 * c + 2] = v.z; hi[4 * c + 3] = v.w;
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const uint32_t h = hi[i] & 0xFFFFE000u;
            lo[i] = __float_as_uint(__uint_as_float(hi[i]) - __uint_as_float(h)) + 0x1000u;       // = tf32_lo_of(raw)
            hi[i] = h;
          }
          // the query columns are still read by the previous item's MMAs until qfree completes
          if (items_done > 0) mbar_wait(qfree_bar, (items_done - 1) & 1u);
          tc_fence_after();
          const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(T_QCOL0 + kb * 32);
          tc_st16(ta, hi);
          tc_st16(ta + 16, lo);
          tc_wait_st();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&conv_bar[stage]);
        }
        ++cnt;
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      for (int64_t tile = t0; tile < t1; ++tile)
        for (int kb = 0; kb < KB; ++kb) {
          if ((int)(cnt & 1u) == grp) {
            mbar_wait(&full_bar[stage], phase);
            const float4* raw = reinterpret_cast<const float4*>(smem + (size_t)stage * T_STAGE);
            float4* lo = reinterpret_cast<float4*>(smem + (size_t)stage * T_STAGE + T_CBYTES);
            float4 x[3];
#pragma unroll
            for (int j = 0; j < 3; ++j) x[j] = raw[gt + j * 128];          // 384 float4 per entry
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              float4 l;
              l.x = tf32_lo_of(x[j].x);
              l.y = tf32_lo_of(x[j].y);
              l.z = tf32_lo_of(x[j].z);
              l.w = tf32_lo_of(x[j].w);
              lo[gt + j * 128] = l;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&conv_bar[stage]);
          }
          ++cnt;
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      ++items_done;
    }
  } else if (warp < 4) {
    // ======================= selection (warps 0..3; lane = query row = TMEM lane) =======================
    // A list entry is ONE 64-bit key: (score_key << 32) | (0xFFFFFFFF - candidate index); a larger key is a better
    // entry (higher score, ties -> lower index), so the total order is a single integer comparison and an insert
    // moves half as many shared-memory words as separate score / index arrays.
    long long* lk_w = list_k + (size_t)(warp * 32) * k;
    constexpr long long EMPTY = LLONG_MIN;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int64_t item = blockIdx.x; item < a.n_items; item += gridDim.x) {
      const int64_t sl = item / a.q_tiles;
      const int qt = (int)(item - sl * a.q_tiles);
      const int64_t t0 = sl * a.tiles_per_slice;
      const int64_t t1 = imin<int64_t>(a.ntiles, t0 + a.tiles_per_slice);
      for (int idx = lane; idx < 32 * k; idx += 32) lk_w[idx] = EMPTY;
      __syncwarp();
      // th = max(k-th best of this slice's list, best k-th ANY slice of this query has published): a candidate below
      // either bound cannot be in the final top-k, so the slices prune each other (without this every slice warms
      // up from -inf and the list inserts, each several dependent shared-memory round trips, dominate)
      float th = -INFINITY;
      const int64_t q_me = (int64_t)qt * TQ + warp * 32 + lane;
      int published = INT_MIN;
      for (int64_t tile = t0; tile < t1; ++tile) {
        if (q_me < a.nq) th = fmaxf(th, key_score(*reinterpret_cast<volatile int*>(a.g_bound + q_me)));
        mbar_wait_relaxed(&tmem_full[acc], acc_phase);
        tc_fence_after();
        const uint32_t t_row = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(acc * 2 * TNC);
        const int64_t c0 = tile * TNC;
        const int nvalid = (int)imin<int64_t>(TNC, a.nc - c0);       // candidates of this tile that exist
        for (int ch = 0; ch < TNC / 16; ++ch) {
          uint32_t r[16], r2[16];
          tc_ld16(t_row + (uint32_t)(ch * 16), r);
          tc_ld16(t_row + (uint32_t)(TNC + ch * 16), r2);
          tc_wait_ld();
          float v[16];
          uint32_t mk = 0;                                 // bit j: score j of this lane's query passes the threshold
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            v[j] = __fadd_rn(__uint_as_float(r[j]), __uint_as_float(r2[j]));
            mk |= (v[j] >= th) ? (1u << j) : 0u;
          }
          const int nv = nvalid - ch * 16;
          if (nv < 16) mk &= nv <= 0 ? 0u : ((1u << nv) - 1u);
          unsigned any = __ballot_sync(0xffffffffu, mk != 0u);
          while (any) {                                    // survivors are rare: one warp-cooperative insert each
            const int src = __ffs(any) - 1;
            const int jj = __ffs(mk) - 1;                  // meaningful on lane src
            float sc_l = 0.f;
#pragma unroll
            for (int j = 0; j < 16; ++j) sc_l = (j == jj) ? v[j] : sc_l;
            const float sc = __shfl_sync(0xffffffffu, sc_l, src);
            const int jb = __shfl_sync(0xffffffffu, jj, src);
            const int ci = (int)(c0 + ch * 16 + jb);
            const long long key = ((long long)score_key(sc) << 32) | (long long)(0xFFFFFFFFu - (uint32_t)ci);
            long long* lk = lk_w + (size_t)src * k;
            long long cur[4];
            int pos = 0;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const int x = lane + 32 * c;
              cur[c] = x < k ? lk[x] : EMPTY;
              pos += __popc(__ballot_sync(0xffffffffu, x < k && cur[c] > key));
            }
            float kth = -INFINITY;
            if (pos < k) {                                 // warp-uniform
              __syncwarp();
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                const int x = lane + 32 * c;
                if (x >= pos && x + 1 < k) lk[x + 1] = cur[c];
              }
              if (lane == 0) lk[pos] = key;
              __syncwarp();
              const long long kk = lk[k - 1];
              if (kk != EMPTY) kth = key_score((int)(kk >> 32));
            }
            if (lane == src) {
              mk &= mk - 1;                                // this candidate is done
              if (kth > th) {
                th = kth;
#pragma unroll
                for (int j = 0; j < 16; ++j) mk &= (v[j] >= th) ? 0xFFFFFFFFu : ~(1u << j);   // re-filter what is left
                const int pkey = score_key(kth);
                if (pkey > published && q_me < a.nq) {
                  atomicMax(a.g_bound + q_me, pkey);
                  published = pkey;
                }
              }
            }
            any = __ballot_sync(0xffffffffu, mk != 0u);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      // partial lists of this item -> workspace [(q * S + slice) * k + r]
      __syncwarp();
      for (int qq = 0; qq < 32; ++qq) {
        const int64_t qg = (int64_t)qt * TQ + warp * 32 + qq;
        if (qg < a.nq) {
          const int64_t o = (qg * a.S + sl) * k;
          for (int r = lane; r < k; r += 32) {
            const long long kk = lk_w[(size_t)qq * k + r];
            a.part_s[o + r] = kk == EMPTY ? -INFINITY : key_score((int)(kk >> 32));
            a.part_i[o + r] = kk == EMPTY ? INT_MAX : (int)(0xFFFFFFFFu - (uint32_t)(kk & 0xFFFFFFFFll));
          }
        }
      }
      __syncwarp();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 13) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
  }
}

std::atomic<int> g_topk_engine{0};        // 0 auto, 1 FFMA tiles only, 2 tensor pipe whenever eligible
std::atomic<long long> g_topk_tc_launches{0};

struct TcPlan {
  bool ok;
  int S, stages, KB;
  int64_t ntiles, tiles_per_slice, n_items;
  int q_tiles;
  size_t smem;
};
TcPlan plan_tc(int64_t nq, int64_t nc, int d, int k) {
  TcPlan p{};
  p.ok = false;
  if ((d & 3) != 0 || d > T_MAXKB * TKB || k > KP || nc < TNC || nc >= (int64_t)INT_MAX - TNC) return p;
  p.KB = (d + TKB - 1) / TKB;
  p.q_tiles = (int)ceil_div<int64_t>(nq, TQ);
  p.ntiles = ceil_div<int64_t>(nc, TNC);
  int64_t S = ceil_div<int64_t>((int64_t)sm_count() * 8, p.q_tiles);     // ~8 items per CTA
  S = imin<int64_t>(S, MERGE_MAX / k);
  S = imin<int64_t>(S, p.ntiles);
  S = imax<int64_t>(1, S);
  p.tiles_per_slice = ceil_div<int64_t>(p.ntiles, S);
  p.S = (int)ceil_div<int64_t>(p.ntiles, p.tiles_per_slice);
  p.n_items = (int64_t)p.S * p.q_tiles;
  const size_t lists = (size_t)TQ * k * 8;
  const size_t fixed = 1024 + lists + 512;
  const size_t budget = 227 * 1024;
  if (fixed + 4 * (size_t)T_STAGE > budget) return p;
  p.stages = (int)imin<int64_t>(T_MAX_STAGES, (int64_t)((budget - fixed) / T_STAGE));
  p.smem = fixed + (size_t)p.stages * T_STAGE;
  p.ok = encode_fn() != nullptr;
  return p;
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

// tensor-pipe path: engine 2 = whenever the shape is eligible; auto = when there is enough work to fill the machine
static bool use_tc(const TcPlan& p, int64_t nq, int64_t nc) {
  const int e = g_topk_engine.load();
  if (!p.ok || e == 1) return false;
  if (e == 2) return true;
  return nq * nc >= ((int64_t)1 << 26);
}

extern "C" int krs_set_topk_engine(int engine) {
  if (engine < 0 || engine > 2) {
    krs::set_error("krs_set_topk_engine: 0 (auto), 1 (fp32 FMA tiles) or 2 (tcgen05 whenever eligible), got %d", engine);
    return KRS_EINVAL;
  }
  g_topk_engine.store(engine);
  return KRS_OK;
}
extern "C" long long krs_topk_tc_launch_count(void) { return g_topk_tc_launches.load(); }

extern "C" size_t krs_topk_workspace_bytes(int64_t nq, int64_t nc, int d, int k) {
  if (nq <= 0 || nc <= 0 || k <= 0 || k > KP) return 0;
  int S = choose_splits(nq, nc, k);
  const TcPlan p = plan_tc(nq, nc, d, k);
  if (p.ok && p.S > S) S = p.S;       // large enough for either engine
  return (size_t)nq * S * k * (sizeof(float) + sizeof(int32_t)) + (p.ok ? (size_t)nq * sizeof(int) : 0);
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
  const TcPlan tp = plan_tc(nq, nc, d, k);
  if (use_tc(tp, nq, nc) && aligned16(Q) && aligned16(C)) {
    const size_t need_tc = (size_t)nq * tp.S * k * (sizeof(float) + sizeof(int32_t)) + (size_t)nq * sizeof(int);
    KRS_REQUIRE(workspace && workspace_bytes >= need_tc, "krs_topk: workspace too small (%zu < %zu)", workspace_bytes, need_tc);
    CUtensorMap mq, mc;
    if (make_map(&mq, Q, nq, d, d, TKB, TQ, CU_TENSOR_MAP_SWIZZLE_64B) &&
        make_map(&mc, C, nc, d, d, TKB, TNC, CU_TENSOR_MAP_SWIZZLE_64B)) {
      TopkTcArgs t;
      t.part_s = reinterpret_cast<float*>(workspace);
      t.part_i = reinterpret_cast<int32_t*>(t.part_s + (size_t)nq * tp.S * k);
      t.g_bound = reinterpret_cast<int*>(t.part_i + (size_t)nq * tp.S * k);
      KRS_CUDA(cudaMemsetAsync(t.g_bound, 0x80, (size_t)nq * sizeof(int), s));     // key 0x80808080 = below every score
      t.nq = nq; t.nc = nc; t.d = d; t.k = k; t.S = tp.S; t.KB = tp.KB; t.stages = tp.stages;
      t.q_tiles = tp.q_tiles; t.tiles_per_slice = tp.tiles_per_slice; t.ntiles = tp.ntiles; t.n_items = tp.n_items;
      KRS_CUDA(cudaFuncSetAttribute(topk_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      const unsigned grid = (unsigned)imax<int64_t>(1, imin<int64_t>(tp.n_items, sm_count()));
      topk_tc_kernel<<<grid, T_THREADS, tp.smem, s>>>(mq, mc, t);
      KRS_LAUNCH_CHECK();
      g_topk_tc_launches.fetch_add(1);
      int P = 1;
      while (P < tp.S * k) P <<= 1;
      const size_t smem2 = (size_t)P * 8;
      KRS_CUDA(cudaFuncSetAttribute(topk_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
      topk_merge_kernel<<<(unsigned)nq, 256, smem2, s>>>(t.part_s, t.part_i, cand_ids, top_scores, top_ids, tp.S, k, P);
      KRS_LAUNCH_CHECK();
      return KRS_OK;
    }
  }
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
