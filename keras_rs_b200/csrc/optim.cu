// optim.cu — optimizer sweeps for the hot path's parameters (Keras 3 update rules).
//
// AdamW (examples/dcn.py:127) has to visit every row of every table each step (moment decay and
// decoupled weight decay touch untouched rows too — that is what the reference's dense-gradient
// path does), so it is a pure HBM streaming kernel: read p,m,v (+g for touched rows), write p,m,v.
// The gradient "arena" + touched bitmap written by krs_gather_bwd lets the sweep skip reading the
// (V,E) gradient for rows that received none (g = 0 there), and re-zeroes what it consumed, so no
// separate 3.3 GB zero-fill or dense-gradient read is ever paid.
// SGD / Adagrad (examples/ml_perf/main.py:203) have zero update where g = 0, so with an arena they
// walk only the touched rows — exactly equal to the dense update
// (precedent: jax/embedding_lookup.py:174-273; oracle jax/test_utils.py:474-497).
#include <math.h>

#include "common.cuh"

namespace krs {
namespace {

struct AdamArgs {
  float lr, b1, b2, eps, wd, alpha;   // alpha = lr * sqrt(1-b2^t)/(1-b1^t)
};

__device__ __forceinline__ void adamw_one(float& p, float& m, float& v, float g, const AdamArgs& a) {
  p = p - p * a.wd * a.lr;                 // decoupled decay (Keras: variable -= variable * wd * lr)
  m = m + (g - m) * (1.f - a.b1);
  v = v + (g * g - v) * (1.f - a.b2);
  p = p - (m * a.alpha) / (sqrtf(v) + a.eps);
}

// One thread per float4.  VEC4 requires n % 4 == 0, row_len % 4 == 0 and 16-byte alignment.
// `ever` (nullable, arena mode only): one bit per row, set once the row has EVER received a gradient.  A row that never
// has still holds m = v = 0 exactly, so the full update reduces — bit for bit — to the decoupled decay of p alone
// (m' = 0, v' = 0, p' = p - p*wd*lr - (0*alpha)/(0+eps)): such rows move 8 instead of 24 bytes per parameter.
template <bool ARENA>
__global__ void __launch_bounds__(256) adamw_vec_kernel(float* __restrict__ p, float* __restrict__ m,
                                                        float* __restrict__ v, float* __restrict__ g,
                                                        const uint32_t* __restrict__ touched,
                                                        const uint32_t* __restrict__ ever,
                                                        int64_t n4, int row_len4,
                                                        AdamArgs a, const float* __restrict__ hyper) {
  if (hyper) { a.lr = hyper[0]; a.b1 = hyper[1]; a.b2 = hyper[2]; a.eps = hyper[3]; a.wd = hyper[4]; a.alpha = hyper[5]; }
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    // ALL loads first, the re-zeroing store of the gradient row last: with `g = load; store 0` ahead of the
    // p/m/v loads the same-address store serialised a second DRAM round trip per thread (probe: 3.2 ms with no
    // touched rows, 5.9 ms at 6.5 %, 13.9 ms at 100 % — benchmarks/adamw_probe.py)
    bool hit = !ARENA;
    if (ARENA) {
      const int64_t row = i / row_len4;
      hit = (touched[row >> 5] >> (row & 31)) & 1u;
      if (ever != nullptr && !hit && !((ever[row >> 5] >> (row & 31)) & 1u)) {     // cold row: decay only
        float4 pc = reinterpret_cast<float4*>(p)[i];
        pc.x = pc.x - pc.x * a.wd * a.lr;
        pc.y = pc.y - pc.y * a.wd * a.lr;
        pc.z = pc.z - pc.z * a.wd * a.lr;
        pc.w = pc.w - pc.w * a.wd * a.lr;
        reinterpret_cast<float4*>(p)[i] = pc;
        continue;
      }
    }
    float4 gp = make_float4(0.f, 0.f, 0.f, 0.f);
    if (hit) gp = reinterpret_cast<const float4*>(g)[i];
    float4 pp = reinterpret_cast<float4*>(p)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    adamw_one(pp.x, mm.x, vv.x, gp.x, a);
    adamw_one(pp.y, mm.y, vv.y, gp.y, a);
    adamw_one(pp.z, mm.z, vv.z, gp.z, a);
    adamw_one(pp.w, mm.w, vv.w, gp.w, a);
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
    if (ARENA && hit) reinterpret_cast<float4*>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

template <bool ARENA>
__global__ void __launch_bounds__(256) adamw_scalar_kernel(float* __restrict__ p, float* __restrict__ m,
                                                           float* __restrict__ v, float* __restrict__ g,
                                                           const uint32_t* __restrict__ touched, int64_t n, int row_len,
                                                           AdamArgs a, const float* __restrict__ hyper) {
  if (hyper) { a.lr = hyper[0]; a.b1 = hyper[1]; a.b2 = hyper[2]; a.eps = hyper[3]; a.wd = hyper[4]; a.alpha = hyper[5]; }
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    bool hit = !ARENA;
    if (ARENA) {
      const int64_t row = i / row_len;
      hit = (touched[row >> 5] >> (row & 31)) & 1u;
    }
    float gp = hit ? g[i] : 0.f;
    float pp = p[i], mm = m[i], vv = v[i];
    adamw_one(pp, mm, vv, gp, a);
    p[i] = pp;
    m[i] = mm;
    v[i] = vv;
    if (ARENA && hit) g[i] = 0.f;
  }
}

// Row-sparse SGD / Adagrad over the touched bitmap.  A warp takes `wpw` (1..32) consecutive bitmap words per pass — lane j < wpw
// fetches word j — then walks their set bits together; a row is updated by the whole warp (lanes stride the row).  The word
// is cleared afterwards by its owner lane.  Rows of one warp are updated one after the other (latency-bound), so the host
// picks `wpw` small enough to spread the bitmap over ~16 CTAs per SM: with 32 words per warp a 1M-row table at 6 % density
// ran on 123 CTAs, 63 serial rows per warp, 245 us per table (profiles/r2_launches_c3_dlrm.md).
__global__ void __launch_bounds__(256) sparse_rows_kernel(float* __restrict__ p, float* __restrict__ acc,
                                                          float* __restrict__ g, uint32_t* __restrict__ touched,
                                                          int64_t nwords, int64_t nrows, int row_len, float lr,
                                                          float eps, int kind, int wpw) {
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t w0 = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * wpw; w0 < nwords; w0 += nwarps * wpw) {
    const int64_t wi = w0 + lane;
    const bool mine = lane < wpw && wi < nwords;
    uint32_t word = mine ? touched[wi] : 0u;
    unsigned nz = __ballot_sync(0xffffffffu, word != 0u);
    while (nz) {
      const int src = __ffs(nz) - 1;
      nz &= nz - 1;
      uint32_t bits = __shfl_sync(0xffffffffu, word, src);
      while (bits) {
        const int bit = __ffs(bits) - 1;
        bits &= bits - 1;
        const int64_t row = (w0 + src) * 32 + bit;
        if (row >= nrows) break;
        const int64_t base = row * (int64_t)row_len;
        for (int c = lane; c < row_len; c += 32) {
          const float gg = g[base + c];                 // loads first, re-zeroing store last (see adamw_vec_kernel)
          const float pv = p[base + c];
          if (kind == 1) {
            const float a2 = acc[base + c] + gg * gg;
            acc[base + c] = a2;
            p[base + c] = pv - lr * gg / sqrtf(a2 + eps);
          } else {
            p[base + c] = pv - lr * gg;
          }
          g[base + c] = 0.f;
        }
      }
    }
    if (mine && word != 0u) touched[wi] = 0u;
  }
}

__global__ void __launch_bounds__(256) dense_sgd_adagrad_kernel(float* __restrict__ p, float* __restrict__ acc,
                                                                const float* __restrict__ g, int64_t n, float lr,
                                                                float eps, int kind) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gg = g[i];
    if (kind == 1) {
      const float a2 = acc[i] + gg * gg;
      acc[i] = a2;
      p[i] = p[i] - lr * gg / sqrtf(a2 + eps);
    } else {
      p[i] = p[i] - lr * gg;
    }
  }
}

// ------------------------------------------------------------------ optimizers on COMPACT gradient rows
// The row-sharded backward (exchange.cu) leaves one gradient row per distinct touched table row in `compact`
// (slot order = arena row order) plus uniq_rows[slot]; nothing table-sized exists besides the table and its slots.
struct RowsHyper { float h[8]; };

template <int KIND>
__device__ __forceinline__ void rows_update_one(float& p, float& s1, float& s2, float g, const RowsHyper& H) {
  if (KIND == KRS_OPT_SGD) {
    p = p - H.h[0] * g;
  } else if (KIND == KRS_OPT_ADAGRAD) {            // keras Adagrad: acc += g^2 ; p -= lr * g / sqrt(acc + eps)
    s1 = s1 + g * g;
    p = p - H.h[0] * g / sqrtf(s1 + H.h[1]);
  } else if (KIND == KRS_OPT_ADAM) {               // keras Adam on the touched rows only (lazy / SparseCore form)
    s1 = s1 + (g - s1) * (1.f - H.h[1]);
    s2 = s2 + (g * g - s2) * (1.f - H.h[2]);
    p = p - (s1 * H.h[4]) / (sqrtf(s2) + H.h[3]);
  } else {                                         // keras Ftrl (l2_shrinkage = 0, the only form the reference accepts)
    const float lr = H.h[0], lrp = H.h[1], l1 = H.h[2], l2 = H.h[3] + H.h[4] / (2.f * lr);
    const float na = s1 + g * g;
    const float pa = (lrp == -0.5f) ? sqrtf(s1) : powf(s1, -lrp);
    const float pn = (lrp == -0.5f) ? sqrtf(na) : powf(na, -lrp);
    s2 = s2 + (g - (pn - pa) / lr * p);
    const float quad = pn / lr + 2.f * l2;
    const float lc = fminf(fmaxf(s2, -l1), l1);
    p = (lc - s2) / quad;
    s1 = na;
  }
}

// LPR lanes per compact row (float4 each per pass): slots are consecutive, table rows increase with the slot.
template <int KIND>
__global__ void __launch_bounds__(256) rows_apply_kernel(float* __restrict__ p, float* __restrict__ s1,
                                                         float* __restrict__ s2, float* __restrict__ compact,
                                                         const int32_t* __restrict__ uniq_rows,
                                                         const uint32_t* __restrict__ n_unique, int64_t cap_rows, int E4,
                                                         int lpr, const RowsHyper H) {
  const int64_t n = krs::imin<int64_t>((int64_t)*n_unique, cap_rows);
  const int sub = threadIdx.x % lpr;
  const int64_t ngroups = ((int64_t)gridDim.x * blockDim.x) / lpr;
  for (int64_t slot = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / lpr; slot < n; slot += ngroups) {
    const int64_t row = uniq_rows[slot];
    for (int c = sub; c < E4; c += lpr) {
      const int64_t gi = slot * E4 + c, pi = row * E4 + c;
      const float4 g = reinterpret_cast<const float4*>(compact)[gi];
      float4 pp = reinterpret_cast<float4*>(p)[pi];
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
      if (KIND != KRS_OPT_SGD) a = reinterpret_cast<float4*>(s1)[pi];
      if (KIND == KRS_OPT_ADAM || KIND == KRS_OPT_FTRL) b = reinterpret_cast<float4*>(s2)[pi];
      rows_update_one<KIND>(pp.x, a.x, b.x, g.x, H);
      rows_update_one<KIND>(pp.y, a.y, b.y, g.y, H);
      rows_update_one<KIND>(pp.z, a.z, b.z, g.z, H);
      rows_update_one<KIND>(pp.w, a.w, b.w, g.w, H);
      reinterpret_cast<float4*>(p)[pi] = pp;
      if (KIND != KRS_OPT_SGD) reinterpret_cast<float4*>(s1)[pi] = a;
      if (KIND == KRS_OPT_ADAM || KIND == KRS_OPT_FTRL) reinterpret_cast<float4*>(s2)[pi] = b;
      reinterpret_cast<float4*>(compact)[gi] = make_float4(0.f, 0.f, 0.f, 0.f);      // loads first, re-zeroing store last
    }
  }
}

// The same four rules straight from a gradient ARENA + touched bitmap (single-GPU tables, TableConfig.optimizer): the
// warp walks 32 bitmap words, LPR-lane groups take the set rows one float4 column each, the arena row is re-zeroed and the
// word cleared.  touched == nullptr: dense gradient, every element is visited (FTRL / Adam on ordinary variables).
template <int KIND>
__global__ void __launch_bounds__(256) rows_apply_arena_kernel(float* __restrict__ p, float* __restrict__ s1, float* __restrict__ s2,
                                                               float* __restrict__ g, uint32_t* __restrict__ touched, int64_t nwords,
                                                               int64_t nrows, int row_len, const RowsHyper H) {
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t w0 = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 32; w0 < nwords; w0 += nwarps * 32) {
    const int64_t wi = w0 + lane;
    const uint32_t word = wi < nwords ? touched[wi] : 0u;
    unsigned nz = __ballot_sync(0xffffffffu, word != 0u);
    while (nz) {
      const int src = __ffs(nz) - 1;
      nz &= nz - 1;
      uint32_t bits = __shfl_sync(0xffffffffu, word, src);
      while (bits) {
        const int bit = __ffs(bits) - 1;
        bits &= bits - 1;
        const int64_t row = (w0 + src) * 32 + bit;
        if (row >= nrows) break;
        const int64_t base = row * (int64_t)row_len;
        for (int c = lane; c < row_len; c += 32) {
          const float gg = g[base + c];
          float pv = p[base + c];
          float a = KIND != KRS_OPT_SGD ? s1[base + c] : 0.f;
          float b = (KIND == KRS_OPT_ADAM || KIND == KRS_OPT_FTRL) ? s2[base + c] : 0.f;
          rows_update_one<KIND>(pv, a, b, gg, H);
          p[base + c] = pv;
          if (KIND != KRS_OPT_SGD) s1[base + c] = a;
          if (KIND == KRS_OPT_ADAM || KIND == KRS_OPT_FTRL) s2[base + c] = b;
          g[base + c] = 0.f;
        }
      }
    }
    if (wi < nwords && word != 0u) touched[wi] = 0u;
  }
}
template <int KIND>
__global__ void __launch_bounds__(256) dense_apply_kernel(float* __restrict__ p, float* __restrict__ s1, float* __restrict__ s2,
                                                          const float* __restrict__ g, int64_t n, const RowsHyper H) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float pv = p[i];
    float a = KIND != KRS_OPT_SGD ? s1[i] : 0.f;
    float b = (KIND == KRS_OPT_ADAM || KIND == KRS_OPT_FTRL) ? s2[i] : 0.f;
    rows_update_one<KIND>(pv, a, b, g[i], H);
    p[i] = pv;
    if (KIND != KRS_OPT_SGD) s1[i] = a;
    if (KIND == KRS_OPT_ADAM || KIND == KRS_OPT_FTRL) s2[i] = b;
  }
}

// adamw_vec_kernel<ARENA> with the gradient row fetched through the slot numbering of the touched bitmap.
__global__ void __launch_bounds__(256) adamw_compact_kernel(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v,
                                                            float* __restrict__ compact, const uint32_t* __restrict__ touched,
                                                            const uint32_t* __restrict__ wordprefix,
                                                            const uint32_t* __restrict__ blockbase,
                                                            const uint32_t* __restrict__ ever, int64_t n4, int row_len4,
                                                            AdamArgs a) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / row_len4;
    const uint32_t word = touched[row >> 5];
    const bool hit = (word >> (row & 31)) & 1u;
    if (ever != nullptr && !hit && !((ever[row >> 5] >> (row & 31)) & 1u)) {     // never-touched row: m = v = 0, decay only
      float4 pc = reinterpret_cast<float4*>(p)[i];
      pc.x = pc.x - pc.x * a.wd * a.lr;
      pc.y = pc.y - pc.y * a.wd * a.lr;
      pc.z = pc.z - pc.z * a.wd * a.lr;
      pc.w = pc.w - pc.w * a.wd * a.lr;
      reinterpret_cast<float4*>(p)[i] = pc;
      continue;
    }
    int64_t gi = 0;
    float4 gp = make_float4(0.f, 0.f, 0.f, 0.f);
    if (hit) {
      const int64_t w = row >> 5;
      const int64_t slot = (int64_t)blockbase[w >> 10] + wordprefix[w] + __popc(word & ((1u << (row & 31)) - 1u));
      gi = slot * row_len4 + (i - row * row_len4);
      gp = reinterpret_cast<const float4*>(compact)[gi];
    }
    float4 pp = reinterpret_cast<float4*>(p)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    adamw_one(pp.x, mm.x, vv.x, gp.x, a);
    adamw_one(pp.y, mm.y, vv.y, gp.y, a);
    adamw_one(pp.z, mm.z, vv.z, gp.z, a);
    adamw_one(pp.w, mm.w, vv.w, gp.w, a);
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
    if (hit) reinterpret_cast<float4*>(compact)[gi] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// ever |= touched ; touched = 0   (after the sweep; replaces the memset of the touched bitmap)
__global__ void __launch_bounds__(256) fold_touched_kernel(uint32_t* __restrict__ ever, uint32_t* __restrict__ touched, int64_t nwords) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t t = touched[i];
    if (t != 0u) {
      ever[i] |= t;
      touched[i] = 0u;
    }
  }
}

// hyper = [lr, b1, b2, eps, wd, alpha, step]: advances the step counter and refreshes the folded bias
// correction ON THE DEVICE, so a CUDA-graph replay of the training step needs no per-step host parameters.
__global__ void adam_hyper_advance_kernel(float* hyper) {
  const double t = (double)hyper[6] + 1.0;
  hyper[6] = (float)t;
  hyper[5] = (float)((double)hyper[0] * sqrt(1.0 - pow((double)hyper[2], t)) / (1.0 - pow((double)hyper[1], t)));
}

}  // namespace
}  // namespace krs

using namespace krs;

extern "C" int krs_adamw_cold(float* p, float* m, float* v, float* g, uint32_t* touched, uint32_t* ever, int64_t n, int row_len,
                              float lr, float b1, float b2, float eps, float wd, int64_t step, const float* hyper_dev,
                              void* stream);
extern "C" int krs_adamw(float* p, float* m, float* v, float* g, uint32_t* touched, int64_t n, int row_len, float lr,
                         float b1, float b2, float eps, float wd, int64_t step, const float* hyper_dev, void* stream) {
  return krs_adamw_cold(p, m, v, g, touched, nullptr, n, row_len, lr, b1, b2, eps, wd, step, hyper_dev, stream);
}

extern "C" int krs_adamw_cold(float* p, float* m, float* v, float* g, uint32_t* touched, uint32_t* ever, int64_t n, int row_len,
                              float lr, float b1, float b2, float eps, float wd, int64_t step, const float* hyper_dev,
                              void* stream) {
  KRS_REQUIRE(p && m && v && g, "krs_adamw: null argument");
  KRS_REQUIRE(ever == nullptr || touched != nullptr, "krs_adamw_cold: the ever-touched bitmap needs the gradient arena");
  KRS_REQUIRE(n >= 0 && (step >= 1 || hyper_dev != nullptr), "krs_adamw: bad n/step");
  KRS_REQUIRE(touched == nullptr || (row_len > 0 && n % row_len == 0), "krs_adamw: arena needs n %% row_len == 0");
  if (n == 0) return KRS_OK;
  cudaStream_t s = as_stream(stream);
  AdamArgs a;
  a.lr = lr; a.b1 = b1; a.b2 = b2; a.eps = eps; a.wd = wd;
  a.alpha = hyper_dev ? 0.f
                      : (float)((double)lr * sqrt(1.0 - pow((double)b2, (double)step)) / (1.0 - pow((double)b1, (double)step)));
  const bool vec = (n % 4 == 0) && aligned16(p) && aligned16(m) && aligned16(v) && aligned16(g) &&
                   (touched == nullptr || row_len % 4 == 0);
  const int64_t work = vec ? n / 4 : n;
  const unsigned grid = (unsigned)krs::imax<int64_t>(1, krs::imin<int64_t>(ceil_div<int64_t>(work, 256), (int64_t)sm_count() * 32));
  if (vec) {
    if (touched) adamw_vec_kernel<true><<<grid, 256, 0, s>>>(p, m, v, g, touched, ever, work, row_len / 4, a, hyper_dev);
    else adamw_vec_kernel<false><<<grid, 256, 0, s>>>(p, m, v, g, nullptr, nullptr, work, 1, a, hyper_dev);
  } else {
    if (touched) adamw_scalar_kernel<true><<<grid, 256, 0, s>>>(p, m, v, g, touched, n, row_len, a, hyper_dev);
    else adamw_scalar_kernel<false><<<grid, 256, 0, s>>>(p, m, v, g, nullptr, n, 1, a, hyper_dev);
  }
  KRS_LAUNCH_CHECK();
  if (touched) {
    const int64_t rows = n / row_len;
    const int64_t nwords = ceil_div<int64_t>(rows, 32);
    if (ever != nullptr) {     // (the scalar kernel ignores `ever` and does the full update; the fold keeps it valid)
      fold_touched_kernel<<<(unsigned)krs::imax<int64_t>(1, krs::imin<int64_t>(ceil_div<int64_t>(nwords, 256), (int64_t)sm_count() * 8)), 256, 0, s>>>(
          ever, touched, nwords);
      KRS_LAUNCH_CHECK();
    } else {
      KRS_CUDA(cudaMemsetAsync(touched, 0, sizeof(uint32_t) * (size_t)nwords, s));
    }
  }
  return KRS_OK;
}

extern "C" int krs_sgd_adagrad(float* p, float* acc, float* g, uint32_t* touched, int64_t n, int row_len, float lr,
                               float eps, int kind, void* stream) {
  KRS_REQUIRE(p && g, "krs_sgd_adagrad: null argument");
  KRS_REQUIRE(kind == 0 || (kind == 1 && acc != nullptr), "krs_sgd_adagrad: kind must be 0 (SGD) or 1 (Adagrad + acc)");
  KRS_REQUIRE(touched == nullptr || (row_len > 0 && n % row_len == 0), "krs_sgd_adagrad: arena needs n %% row_len == 0");
  if (n == 0) return KRS_OK;
  cudaStream_t s = as_stream(stream);
  if (touched) {
    const int64_t rows = n / row_len;
    const int64_t nwords = ceil_div<int64_t>(rows, 32);
    const int64_t max_ctas = (int64_t)sm_count() * 16;
    int wpw = 32;                                          // words per warp and pass: halve until the grid fills the GPU
    while (wpw > 1 && ceil_div<int64_t>(ceil_div<int64_t>(nwords, wpw), 8) < max_ctas) wpw >>= 1;
    const unsigned grid = (unsigned)krs::imax<int64_t>(1, krs::imin<int64_t>(ceil_div<int64_t>(ceil_div<int64_t>(nwords, wpw), 8), max_ctas));
    sparse_rows_kernel<<<grid, 256, 0, s>>>(p, acc, g, touched, nwords, rows, row_len, lr, eps, kind, wpw);
  } else {
    const unsigned grid = (unsigned)krs::imax<int64_t>(1, krs::imin<int64_t>(ceil_div<int64_t>(n, 256), (int64_t)sm_count() * 32));
    dense_sgd_adagrad_kernel<<<grid, 256, 0, s>>>(p, acc, g, n, lr, eps, kind);
  }
  KRS_LAUNCH_CHECK();
  return KRS_OK;
}

extern "C" int krs_adam_hyper_advance(float* hyper_dev, void* stream) {
  KRS_REQUIRE(hyper_dev != nullptr, "krs_adam_hyper_advance: null argument");
  adam_hyper_advance_kernel<<<1, 1, 0, as_stream(stream)>>>(hyper_dev);
  KRS_LAUNCH_CHECK();
  return KRS_OK;
}

extern "C" int krs_rows_apply(float* p, float* s1, float* s2, float* compact, const int32_t* uniq_rows, const uint32_t* n_unique,
                              int64_t cap_rows, int E, int kind, const float* hyper, uint32_t* touched_to_clear, int64_t nwords,
                              void* stream) {
  KRS_REQUIRE(p && compact && uniq_rows && n_unique && hyper, "krs_rows_apply: null argument");
  KRS_REQUIRE(kind >= KRS_OPT_SGD && kind <= KRS_OPT_FTRL, "krs_rows_apply: unknown optimizer kind %d", kind);
  KRS_REQUIRE(kind == KRS_OPT_SGD || s1 != nullptr, "krs_rows_apply: optimizer kind %d needs its first slot variable", kind);
  KRS_REQUIRE((kind != KRS_OPT_ADAM && kind != KRS_OPT_FTRL) || s2 != nullptr, "krs_rows_apply: optimizer kind %d needs two slot variables", kind);
  KRS_REQUIRE(E >= 4 && E % 4 == 0 && aligned16(p) && aligned16(compact) && (s1 == nullptr || aligned16(s1)) &&
                  (s2 == nullptr || aligned16(s2)),
              "krs_rows_apply: rows must be 16-byte aligned multiples of 4 floats");
  KRS_REQUIRE(cap_rows > 0, "krs_rows_apply: cap_rows must be positive");
  RowsHyper H;
  for (int i = 0; i < 8; ++i) H.h[i] = hyper[i];
  int lpr = 1;
  while (lpr < E / 4 && lpr < 32) lpr <<= 1;
  // the number of distinct rows lives on the device: size the grid for the capacity, capped at a few waves
  const int64_t groups_per_cta = 256 / lpr;
  const unsigned grid = (unsigned)krs::imax<int64_t>(1, krs::imin<int64_t>(ceil_div<int64_t>(cap_rows, groups_per_cta), (int64_t)sm_count() * 16));
  cudaStream_t s = as_stream(stream);
  switch (kind) {
    case KRS_OPT_SGD: rows_apply_kernel<KRS_OPT_SGD><<<grid, 256, 0, s>>>(p, s1, s2, compact, uniq_rows, n_unique, cap_rows, E / 4, lpr, H); break;
    case KRS_OPT_ADAGRAD: rows_apply_kernel<KRS_OPT_ADAGRAD><<<grid, 256, 0, s>>>(p, s1, s2, compact, uniq_rows, n_unique, cap_rows, E / 4, lpr, H); break;
    case KRS_OPT_ADAM: rows_apply_kernel<KRS_OPT_ADAM><<<grid, 256, 0, s>>>(p, s1, s2, compact, uniq_rows, n_unique, cap_rows, E / 4, lpr, H); break;
    default: rows_apply_kernel<KRS_OPT_FTRL><<<grid, 256, 0, s>>>(p, s1, s2, compact, uniq_rows, n_unique, cap_rows, E / 4, lpr, H); break;
  }
  KRS_LAUNCH_CHECK();
  if (touched_to_clear != nullptr && nwords > 0) KRS_CUDA(cudaMemsetAsync(touched_to_clear, 0, sizeof(uint32_t) * (size_t)nwords, s));
  return KRS_OK;
}

extern "C" int krs_adamw_compact(float* p, float* m, float* v, float* compact, uint32_t* touched, const uint32_t* wordprefix,
                                 const uint32_t* blockbase, uint32_t* ever, int64_t n, int row_len, float lr, float b1, float b2,
                                 float eps, float wd, int64_t step, void* stream) {
  KRS_REQUIRE(p && m && v && compact && touched && wordprefix && blockbase, "krs_adamw_compact: null argument");
  KRS_REQUIRE(n >= 0 && step >= 1 && row_len >= 4 && row_len % 4 == 0 && n % row_len == 0, "krs_adamw_compact: bad n / row_len / step");
  KRS_REQUIRE(aligned16(p) && aligned16(m) && aligned16(v) && aligned16(compact), "krs_adamw_compact: buffers must be 16-byte aligned");
  if (n == 0) return KRS_OK;
  AdamArgs a;
  a.lr = lr; a.b1 = b1; a.b2 = b2; a.eps = eps; a.wd = wd;
  a.alpha = (float)((double)lr * sqrt(1.0 - pow((double)b2, (double)step)) / (1.0 - pow((double)b1, (double)step)));
  const int64_t work = n / 4;
  const unsigned grid = (unsigned)krs::imax<int64_t>(1, krs::imin<int64_t>(ceil_div<int64_t>(work, 256), (int64_t)sm_count() * 32));
  cudaStream_t s = as_stream(stream);
  adamw_compact_kernel<<<grid, 256, 0, s>>>(p, m, v, compact, touched, wordprefix, blockbase, ever, work, row_len / 4, a);
  KRS_LAUNCH_CHECK();
  const int64_t nwords = ceil_div<int64_t>(n / row_len, 32);
  if (ever != nullptr) {
    fold_touched_kernel<<<(unsigned)krs::imax<int64_t>(1, krs::imin<int64_t>(ceil_div<int64_t>(nwords, 256), (int64_t)sm_count() * 8)), 256, 0, s>>>(
        ever, touched, nwords);
    KRS_LAUNCH_CHECK();
  } else {
    KRS_CUDA(cudaMemsetAsync(touched, 0, sizeof(uint32_t) * (size_t)nwords, s));
  }
  return KRS_OK;
}


extern "C" int krs_opt_apply(float* p, float* s1, float* s2, float* g, uint32_t* touched, int64_t n, int row_len, int kind,
                             const float* hyper, void* stream) {
  KRS_REQUIRE(p && g && hyper, "krs_opt_apply: null argument");
  KRS_REQUIRE(kind >= KRS_OPT_SGD && kind <= KRS_OPT_FTRL, "krs_opt_apply: unknown optimizer kind %d", kind);
  KRS_REQUIRE(kind == KRS_OPT_SGD || s1 != nullptr, "krs_opt_apply: optimizer kind %d needs its first slot variable", kind);
  KRS_REQUIRE((kind != KRS_OPT_ADAM && kind != KRS_OPT_FTRL) || s2 != nullptr, "krs_opt_apply: optimizer kind %d needs two slot variables", kind);
  KRS_REQUIRE(touched == nullptr || (row_len > 0 && n % row_len == 0), "krs_opt_apply: arena needs n %% row_len == 0");
  if (n == 0) return KRS_OK;
  RowsHyper H;
  for (int i = 0; i < 8; ++i) H.h[i] = hyper[i];
  cudaStream_t s = as_stream(stream);
  if (touched) {
    const int64_t rows = n / row_len, nwords = ceil_div<int64_t>(rows, 32);
    const unsigned grid = (unsigned)krs::imax<int64_t>(1, krs::imin<int64_t>(ceil_div<int64_t>(ceil_div<int64_t>(nwords, 32), 8), (int64_t)sm_count() * 16));
    switch (kind) {
      case KRS_OPT_SGD: rows_apply_arena_kernel<KRS_OPT_SGD><<<grid, 256, 0, s>>>(p, s1, s2, g, touched, nwords, rows, row_len, H); break;
      case KRS_OPT_ADAGRAD: rows_apply_arena_kernel<KRS_OPT_ADAGRAD><<<grid, 256, 0, s>>>(p, s1, s2, g, touched, nwords, rows, row_len, H); break;
      case KRS_OPT_ADAM: rows_apply_arena_kernel<KRS_OPT_ADAM><<<grid, 256, 0, s>>>(p, s1, s2, g, touched, nwords, rows, row_len, H); break;
      default: rows_apply_arena_kernel<KRS_OPT_FTRL><<<grid, 256, 0, s>>>(p, s1, s2, g, touched, nwords, rows, row_len, H); break;
    }
  } else {
    const unsigned grid = (unsigned)krs::imax<int64_t>(1, krs::imin<int64_t>(ceil_div<int64_t>(n, 256), (int64_t)sm_count() * 32));
    switch (kind) {
      case KRS_OPT_SGD: dense_apply_kernel<KRS_OPT_SGD><<<grid, 256, 0, s>>>(p, s1, s2, g, n, H); break;
      case KRS_OPT_ADAGRAD: dense_apply_kernel<KRS_OPT_ADAGRAD><<<grid, 256, 0, s>>>(p, s1, s2, g, n, H); break;
      case KRS_OPT_ADAM: dense_apply_kernel<KRS_OPT_ADAM><<<grid, 256, 0, s>>>(p, s1, s2, g, n, H); break;
      default: dense_apply_kernel<KRS_OPT_FTRL><<<grid, 256, 0, s>>>(p, s1, s2, g, n, H); break;
    }
  }
  KRS_LAUNCH_CHECK();
  return KRS_OK;
}
