// exchange.cu — row-sharded (MOD) embedding exchange over NVLink peer memory (config C5).
//
// The reference shards tables only on TPU SparseCore: row r lives on shard r % S at local row r / S
// (jax/embedding_utils.py:187-197 "MOD"; tensorflow/distributed_embedding.py:316-328), ids are routed to
// their owner, owners look the rows up, activations come back, and the backward applies the optimizer on the
// owner without ever building a dense gradient (jax/embedding_lookup.py:174-273).  This file is that protocol
// for one NVSwitch box, one process per GPU, every buffer in a cudaIpc-exported "region" all peers map:
//
//   requester r                                   owner o
//   1 route:  ids (B,F) -> per-owner request lists
//             (arena row on the owner, position
//             b*F+f), stable counting sort
//   2 barrier ------------------------------------ 2 barrier            (flags in peer memory, no NCCL)
//                                                  3 gather_push: ONE launch over all requesters' lists:
//                                                    list entries read sequentially over NVLink, table rows
//                                                    read locally (random), rows written straight into the
//                                                    requester's activation at increasing positions
//   4 barrier ------------------------------------ 4 barrier
//   5 dense forward / backward (local)
//   6 barrier ------------------------------------ 6 barrier
//                                                  7 grad_pull: ONE launch: gradient rows read from the
//                                                    requesters' dL/dx0 at increasing positions, duplicates
//                                                    combined, accumulated into a COMPACT buffer with one row
//                                                    per distinct touched table row (slot = rank of the row's
//                                                    bit in the touched bitmap, from a prefix popcount)
//                                                  8 optimizer on the compact rows (optim.cu)
// Random accesses never cross NVLink (measured: 58 GB/s random vs 660-730 GB/s monotonic,
// profiles/r1_p2p_probe_2gpu.txt), and no table-sized gradient buffer exists.
#include "common.cuh"

namespace krs {
namespace {

constexpr int MAXS = KRS_XCHG_MAX_SHARDS;
constexpr int ROUTE_CHUNK = 2048;             // positions per CTA of the routing kernels (8 warps x 8 x 32)
constexpr int ROUTE_SUB = ROUTE_CHUNK / 32;   // 32-position sub-blocks per chunk

struct Peers {
  unsigned char* base[MAXS];
};

struct RouteArgs {
  const void* ids;
  int64_t ids_ld;
  int64_t P;            // positions = B*F
  int F, S;
  const int64_t* vocab;           // device [F]
  const int32_t* owner_row_off;   // device [S*F]: first arena row of table f on owner o
  int32_t* counts;                // [nchunks*S] per-chunk per-owner counts -> (after the scan) exclusive offsets
  int32_t* hdr;                   // [S+1] bucket starts (written by the scan)
  int32_t* rows;                  // [P] request lists, bucket o at hdr[o]
  int32_t* pos;
  float* x0;                      // nullable: invalid ids get a NaN row here (jnp.take mode="fill")
  int E;
  int shift;                      // log2(S) or -1
};

template <typename IdT>
__device__ __forceinline__ bool route_one(const RouteArgs& a, int64_t p, int& owner, int32_t& row) {
  const int64_t b = (a.P < 0x7fffffffLL) ? (int64_t)((uint32_t)p / (uint32_t)a.F) : p / a.F;
  const int f = (int)(p - b * a.F);
  int64_t id = (int64_t)reinterpret_cast<const IdT*>(a.ids)[b * a.ids_ld + f];
  const int64_t v = a.vocab[f];
  if (id < 0) id += v;                              // negative ids count from the end (numpy / jnp.take)
  if ((uint64_t)id >= (uint64_t)v) return false;    // still out of range: no owner (forward NaN, backward dropped)
  int64_t local;
  if (a.shift >= 0) { owner = (int)(id & (a.S - 1)); local = id >> a.shift; }
  else if (v < 0x7fffffffLL) { owner = (int)((uint32_t)id % (uint32_t)a.S); local = (uint32_t)id / (uint32_t)a.S; }
  else { owner = (int)(id % a.S); local = id / a.S; }
  row = a.owner_row_off[owner * a.F + f] + (int32_t)local;
  return true;
}

// FILL = false: counts[chunk][o] = ids of this chunk owned by o.
// FILL = true : counts holds the exclusive prefix over chunks; entries are written in position order (stable), so
//               every bucket lists increasing positions and the owner's remote accesses are monotonic.
template <typename IdT, bool FILL>
__global__ void __launch_bounds__(256) route_kernel(const RouteArgs a) {
  __shared__ int sub_cnt[ROUTE_SUB][MAXS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < ROUTE_SUB * MAXS; i += blockDim.x) (&sub_cnt[0][0])[i] = 0;
  __syncthreads();
  const int64_t chunk0 = (int64_t)blockIdx.x * ROUTE_CHUNK;
  int owner[ROUTE_SUB / 8];
  int32_t row[ROUTE_SUB / 8];
  int rank[ROUTE_SUB / 8];
#pragma unroll
  for (int it = 0; it < ROUTE_SUB / 8; ++it) {
    const int sb = it * 8 + warp;
    const int64_t p = chunk0 + (int64_t)sb * 32 + lane;
    owner[it] = -1;
    row[it] = 0;
    if (p < a.P) {
      int o;
      int32_t r;
      if (route_one<IdT>(a, p, o, r)) { owner[it] = o; row[it] = r; }
      else if (FILL && a.x0 != nullptr) {
        const float qnan = __int_as_float(0x7fc00000);
        for (int c = 0; c < a.E; ++c) a.x0[p * a.E + c] = qnan;
      }
    }
    // lanes with the same owner: rank inside the sub-block and the sub-block's count (invalid lanes get unique keys)
    const unsigned same = __match_any_sync(0xffffffffu, owner[it] >= 0 ? owner[it] : -1 - lane);
    rank[it] = __popc(same & ((1u << lane) - 1u));
    if (owner[it] >= 0 && rank[it] == 0) sub_cnt[sb][owner[it]] = __popc(same);
  }
  __syncthreads();
  if (!FILL) {
    if (threadIdx.x < a.S) {
      int t = 0;
      for (int sb = 0; sb < ROUTE_SUB; ++sb) t += sub_cnt[sb][threadIdx.x];
      a.counts[(int64_t)blockIdx.x * a.S + threadIdx.x] = t;
    }
    return;
  }
  if (threadIdx.x < a.S) {   // exclusive prefix over the sub-blocks, based at this chunk's offset inside the bucket
    int run = a.hdr[threadIdx.x] + a.counts[(int64_t)blockIdx.x * a.S + threadIdx.x];
    for (int sb = 0; sb < ROUTE_SUB; ++sb) {
      const int c = sub_cnt[sb][threadIdx.x];
      sub_cnt[sb][threadIdx.x] = run;
      run += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int it = 0; it < ROUTE_SUB / 8; ++it) {
    if (owner[it] < 0) continue;
    const int sb = it * 8 + warp;
    const int dst = sub_cnt[sb][owner[it]] + rank[it];
    a.rows[dst] = row[it];
    a.pos[dst] = (int32_t)(chunk0 + sb * 32 + lane);
  }
}

// One warp per owner: exclusive scan of counts[:, o] over the chunks; bucket totals -> hdr (exclusive over owners).
__global__ void __launch_bounds__(32 * MAXS) route_scan_kernel(int32_t* __restrict__ counts, int nchunks, int S,
                                                               int32_t* __restrict__ hdr) {
  __shared__ int total[MAXS];
  const int lane = threadIdx.x & 31, o = threadIdx.x >> 5;
  if (o < S) {
    int carry = 0;
    int nxt = lane < nchunks ? counts[(int64_t)lane * S + o] : 0;
    for (int c0 = 0; c0 < nchunks; c0 += 32) {
      const int c = c0 + lane;
      const int v = nxt;
      nxt = (c + 32 < nchunks) ? counts[(int64_t)(c + 32) * S + o] : 0;
      int inc = v;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
      }
      if (c < nchunks) counts[(int64_t)c * S + o] = carry + inc - v;
      carry += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) total[o] = carry;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int i = 0; i < S; ++i) { hdr[i] = run; run += total[i]; }
    hdr[S] = run;
  }
}

// ------------------------------------------------------------------ flag barrier in peer memory
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// Thread p signals peer p (writes `epoch` into slot `me` of p's flag array), then waits until peer p's signal for
// this epoch has arrived in the local array.  Everything this rank enqueued before the barrier has completed (stream
// order) before the signal is released at system scope.  A peer that never arrives trips the timeout: the error
// word is set and the kernel returns instead of hanging the GPU.
__global__ void xchg_barrier_kernel(const Peers peers, int64_t off_flags, int S, int me, uint32_t epoch,
                                    unsigned long long timeout_ns) {
  const int p = threadIdx.x;
  if (p >= S) return;
  __threadfence_system();
  uint32_t* remote = reinterpret_cast<uint32_t*>(peers.base[p] + off_flags);
  st_release_sys(remote + me, epoch);
  uint32_t* mine = reinterpret_cast<uint32_t*>(peers.base[me] + off_flags);
  const unsigned long long t0 = globaltimer_ns();
  while ((int32_t)(ld_acquire_sys(mine + p) - epoch) < 0) {
    if (globaltimer_ns() - t0 > timeout_ns) {
      atomicOr(mine + MAXS, 1u << (p & 31));      // error word: which peers were missing
      break;
    }
    __nanosleep(200);
  }
  __threadfence_system();
}

// ------------------------------------------------------------------ owner side, forward
struct OwnerArgs {
  Peers peers;
  int64_t off_hdr, off_rows, off_pos, off_x0, off_grad;   // byte offsets inside a region (lists: current parity)
  int S, me, E;
  const float* arena;       // local shard: all tables, row-major (rows, E)
  uint32_t* touched;        // nullable: bit per local arena row, set for every requested row (training)
  // backward
  const uint32_t* wordprefix;   // per bitmap word: set bits before it inside its 1024-word block
  const uint32_t* blockbase;    // per 1024-word block: set bits before the block
  float* compact;               // (cap_rows, E) one row per distinct touched arena row
  int32_t* uniq_rows;           // (cap_rows) arena row of every slot
  int64_t cap_rows;
  uint32_t* err;                // error word (local flags[MAXS]): bit 31 = compact buffer overflow
};

struct OwnerCtaState {
  int start[MAXS];      // first entry of my bucket inside requester r's lists
  int count[MAXS];
  int chunk0[MAXS + 1]; // first 32-entry chunk of stage k (requester (me + k) % S)
};

__device__ __forceinline__ void owner_prologue(const OwnerArgs& a, OwnerCtaState& st) {
  if (threadIdx.x < a.S) {
    const int32_t* hdr = reinterpret_cast<const int32_t*>(a.peers.base[threadIdx.x] + a.off_hdr);
    const int s0 = hdr[a.me], s1 = hdr[a.me + 1];
    st.start[threadIdx.x] = s0;
    st.count[threadIdx.x] = s1 - s0;
  }
  __syncthreads();
  // Requesters are served in ROTATED order, stage k = requester (me + k) % S: all owners run their stages roughly in
  // lockstep, so with the plain order 0, 1, 2 ... every GPU of the box would push to (pull from) the same peer at the
  // same time and that peer's NVLink ingress (egress) would be the whole system's bandwidth.
  if (threadIdx.x == 0) {
    int run = 0;
    for (int k = 0; k < a.S; ++k) { st.chunk0[k] = run; run += (st.count[(a.me + k) % a.S] + 31) >> 5; }
    st.chunk0[a.S] = run;
  }
  __syncthreads();
}

// LPR lanes move one row (16 bytes each per pass); rows wider than LPR*4 floats take several passes.
template <int LPR>
__global__ void __launch_bounds__(256) gather_push_kernel(const OwnerArgs a) {
  constexpr int RPW = 32 / LPR;
  constexpr int STEPS = LPR;                // sub-steps to cover the 32 rows of a chunk
  constexpr int U = STEPS < 8 ? STEPS : 8;
  __shared__ OwnerCtaState st;
  owner_prologue(a, st);
  const int lane = threadIdx.x & 31;
  const int sub = lane % LPR, rsub = lane / LPR;
  const int E4 = a.E >> 2;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int total = st.chunk0[a.S];
  for (int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; c < total; c += nwarps) {
    int k = 0;
    while (c >= st.chunk0[k + 1]) ++k;
    const int r = (a.me + k) % a.S;
    const int j = ((c - st.chunk0[k]) << 5) + lane;
    const float* src = nullptr;
    float* dst = nullptr;
    if (j < st.count[r]) {
      const unsigned char* rb = a.peers.base[r];
      const int64_t e = (int64_t)st.start[r] + j;
      const int32_t row = reinterpret_cast<const int32_t*>(rb + a.off_rows)[e];     // sequential remote reads
      const int32_t pos = reinterpret_cast<const int32_t*>(rb + a.off_pos)[e];
      src = a.arena + (int64_t)row * a.E;
      dst = reinterpret_cast<float*>(const_cast<unsigned char*>(rb) + a.off_x0) + (int64_t)pos * a.E;
      if (a.touched) atomicOr(a.touched + (row >> 5), 1u << (row & 31));
    }
    for (int c0 = 0; c0 < E4; c0 += LPR) {       // uniform trip count: the shuffles below need the whole warp
      const int c4 = c0 + sub;
      const bool in_row = c4 < E4;
#pragma unroll
      for (int s0 = 0; s0 < STEPS; s0 += U) {
        float4 v[U];
        float* q[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int jj = (s0 + u) * RPW + rsub;
          const float* sp = reinterpret_cast<const float*>(__shfl_sync(0xffffffffu, (unsigned long long)src, jj));
          q[u] = reinterpret_cast<float*>(__shfl_sync(0xffffffffu, (unsigned long long)dst, jj));
          if (sp && in_row) v[u] = ldg_nc_f4(sp + c4 * 4);
          else q[u] = nullptr;
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (q[u]) *reinterpret_cast<float4*>(q[u] + c4 * 4) = v[u];     // remote write, increasing positions
      }
    }
  }
}

// ------------------------------------------------------------------ touched bitmap -> slot numbering
// wordprefix[w] = set bits in words [block start, w) ; blocksum[b] = set bits of block b (1024 words = 32768 rows)
__global__ void __launch_bounds__(256) bitmap_block_scan_kernel(const uint32_t* __restrict__ bits, int64_t nwords,
                                                                uint32_t* __restrict__ wordprefix,
                                                                uint32_t* __restrict__ blocksum) {
  __shared__ uint32_t wsum[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t w0 = (int64_t)blockIdx.x * 1024 + threadIdx.x * 4;
  uint32_t c[4];
  uint32_t t = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    c[i] = (w0 + i < nwords) ? __popc(bits[w0 + i]) : 0u;
    t += c[i];
  }
  uint32_t inc = t;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t x = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += x;
  }
  if (lane == 31) wsum[warp] = inc;
  __syncthreads();
  uint32_t base = 0;
  for (int i = 0; i < warp; ++i) base += wsum[i];
  uint32_t run = base + inc - t;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (w0 + i < nwords) wordprefix[w0 + i] = run;
    run += c[i];
  }
  if (threadIdx.x == 255) blocksum[blockIdx.x] = base + inc;
}
// single CTA: blockbase = exclusive scan of blocksum (in place), total -> *n_unique
__global__ void __launch_bounds__(1024) bitmap_base_scan_kernel(uint32_t* __restrict__ blocksum, int nblocks,
                                                                uint32_t* __restrict__ n_unique) {
  __shared__ uint32_t wsum[32];
  __shared__ uint32_t carry_s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int b0 = 0; b0 < nblocks; b0 += 1024) {
    const int b = b0 + threadIdx.x;
    const uint32_t v = b < nblocks ? blocksum[b] : 0u;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t x = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += x;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    uint32_t base = carry_s;
    for (int i = 0; i < warp; ++i) base += wsum[i];
    if (b < nblocks) blocksum[b] = base + inc - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = base + inc;
    __syncthreads();
  }
  if (threadIdx.x == 0) *n_unique = carry_s;
}

__device__ __forceinline__ uint32_t slot_of(const uint32_t* touched, const uint32_t* wordprefix, const uint32_t* blockbase,
                                            int32_t row) {
  const int w = row >> 5;
  return blockbase[w >> 10] + wordprefix[w] + __popc(touched[w] & ((1u << (row & 31)) - 1u));
}

// ------------------------------------------------------------------ owner side, backward
// Warp = 32 consecutive entries of one requester's bucket.  Duplicate rows inside the warp are combined in registers
// (leaders), rows of U leaders are in flight together (remote reads need the memory-level parallelism), then ONE
// 16-byte reduction per lane into the compact row.
template <int LPR>
__global__ void __launch_bounds__(256) grad_pull_kernel(const OwnerArgs a) {
  constexpr int RPW = 32 / LPR;
  constexpr int U = 4;
  __shared__ OwnerCtaState st;
  __shared__ const float* s_src[8][32];
  __shared__ float* s_dst[8][32];
  __shared__ unsigned s_peers[8][32];
  owner_prologue(a, st);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane % LPR, rsub = lane / LPR;
  const int E4 = a.E >> 2;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int total = st.chunk0[a.S];
  for (int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; c < total; c += nwarps) {
    int k = 0;
    while (c >= st.chunk0[k + 1]) ++k;
    const int r = (a.me + k) % a.S;
    const int j = ((c - st.chunk0[k]) << 5) + lane;
    int32_t row = -1 - lane;
    const float* src = nullptr;
    float* dst = nullptr;
    if (j < st.count[r]) {
      const unsigned char* rb = a.peers.base[r];
      const int64_t e = (int64_t)st.start[r] + j;
      row = reinterpret_cast<const int32_t*>(rb + a.off_rows)[e];
      const int32_t pos = reinterpret_cast<const int32_t*>(rb + a.off_pos)[e];
      src = reinterpret_cast<const float*>(rb + a.off_grad) + (int64_t)pos * a.E;
    }
    const unsigned peers_m = __match_any_sync(0xffffffffu, row);
    const bool leader = row >= 0 && (__ffs(peers_m) - 1) == lane;
    if (leader) {
      const uint32_t slot = slot_of(a.touched, a.wordprefix, a.blockbase, row);
      if ((int64_t)slot < a.cap_rows) {
        dst = a.compact + (int64_t)slot * a.E;
        a.uniq_rows[slot] = row;
      } else {
        atomicOr(a.err, 0x80000000u);
      }
    }
    __syncwarp();
    s_src[warp][lane] = src;
    s_dst[warp][lane] = dst;
    const unsigned leaders = __ballot_sync(0xffffffffu, leader && dst != nullptr);
    s_peers[warp][lane] = peers_m;             // the duplicate set of every leader, for whichever lane group serves it
    __syncwarp();
    const int nlead = __popc(leaders);
    for (int c4 = sub; c4 < E4; c4 += LPR) {
      for (int base = 0; base < nlead; base += RPW * U) {
        float4 acc[U];
        int lead[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int nth = base + u * RPW + rsub;
          lead[u] = nth < nlead ? (int)__fns(leaders, 0, nth + 1) : -1;
          if (lead[u] >= 0) acc[u] = ldg_nc_f4(s_src[warp][lead[u]] + c4 * 4);     // first row of U leaders in flight
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (lead[u] < 0) continue;
          for (unsigned q = s_peers[warp][lead[u]] & ~(1u << lead[u]); q; q &= q - 1) {   // rare: duplicates of that row
            const float4 v = ldg_nc_f4(s_src[warp][__ffs(q) - 1] + c4 * 4);
            acc[u].x += v.x; acc[u].y += v.y; acc[u].z += v.z; acc[u].w += v.w;
          }
          atomicAdd(reinterpret_cast<float4*>(s_dst[warp][lead[u]] + c4 * 4), acc[u]);
        }
      }
    }
    __syncwarp();
  }
}

int check_xchg(const krs_xchg_t* x) {
  KRS_REQUIRE(x != nullptr, "krs_xchg: null descriptor");
  KRS_REQUIRE(x->S >= 1 && x->S <= MAXS && x->me >= 0 && x->me < x->S, "krs_xchg: need 1 <= S <= %d and 0 <= me < S", MAXS);
  KRS_REQUIRE(x->F >= 1 && x->E >= 4 && x->E % 4 == 0 && x->B >= 0, "krs_xchg: need F >= 1, E a multiple of 4, B >= 0");
  KRS_REQUIRE(x->B * (int64_t)x->F < 0x7fffffffLL, "krs_xchg: B*F must fit 31 bits");
  for (int s = 0; s < x->S; ++s) KRS_REQUIRE(x->peer_base[s] != nullptr, "krs_xchg: peer %d has no region", s);
  return KRS_OK;
}

Peers peers_of(const krs_xchg_t* x) {
  Peers p;
  for (int s = 0; s < MAXS; ++s) p.base[s] = s < x->S ? reinterpret_cast<unsigned char*>(x->peer_base[s]) : nullptr;
  return p;
}

int lanes_per_row(int E) {
  int lpr = 1;
  while (lpr < E / 4 && lpr < 32) lpr <<= 1;
  return lpr;
}

OwnerArgs owner_args(const krs_xchg_t* x, int parity) {
  OwnerArgs a{};
  a.peers = peers_of(x);
  const int64_t P = x->B * (int64_t)x->F;
  a.off_hdr = x->off_hdr + (int64_t)parity * (MAXS + 1) * 4;
  a.off_rows = x->off_rows + (int64_t)parity * P * 4;
  a.off_pos = x->off_pos + (int64_t)parity * P * 4;
  a.off_x0 = x->off_x0;
  a.off_grad = x->off_grad;
  a.S = x->S;
  a.me = x->me;
  a.E = x->E;
  a.err = reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned char*>(x->peer_base[x->me]) + x->off_flags) + MAXS;
  return a;
}

}  // namespace
}  // namespace krs

using namespace krs;

extern "C" size_t krs_xchg_route_workspace_bytes(int64_t B, int F, int S) {
  const int64_t nchunks = ceil_div<int64_t>(B * (int64_t)F, ROUTE_CHUNK);
  return (size_t)(nchunks > 0 ? nchunks : 1) * (size_t)S * sizeof(int32_t);
}

extern "C" int krs_xchg_route(const krs_xchg_t* x, int parity, const void* ids, int ids_i64, int64_t ids_ld,
                              const int64_t* vocab_dev, const int32_t* owner_row_off_dev, int32_t* workspace,
                              int fill_invalid_nan, void* stream) {
  int rc = check_xchg(x);
  if (rc) return rc;
  KRS_REQUIRE(ids && vocab_dev && owner_row_off_dev && workspace, "krs_xchg_route: null argument");
  KRS_REQUIRE(parity == 0 || parity == 1, "krs_xchg_route: parity must be 0 or 1");
  KRS_REQUIRE(ids_ld >= x->F, "krs_xchg_route: ids_ld < F");
  cudaStream_t s = as_stream(stream);
  unsigned char* mine = reinterpret_cast<unsigned char*>(x->peer_base[x->me]);
  const int64_t P = x->B * (int64_t)x->F;
  RouteArgs a{};
  a.ids = ids;
  a.ids_ld = ids_ld;
  a.P = P;
  a.F = x->F;
  a.S = x->S;
  a.vocab = vocab_dev;
  a.owner_row_off = owner_row_off_dev;
  a.counts = workspace;
  a.hdr = reinterpret_cast<int32_t*>(mine + x->off_hdr) + parity * (MAXS + 1);
  a.rows = reinterpret_cast<int32_t*>(mine + x->off_rows) + (int64_t)parity * P;
  a.pos = reinterpret_cast<int32_t*>(mine + x->off_pos) + (int64_t)parity * P;
  a.x0 = fill_invalid_nan ? reinterpret_cast<float*>(mine + x->off_x0) : nullptr;
  a.E = x->E;
  a.shift = (x->S & (x->S - 1)) == 0 ? 31 - __builtin_clz((unsigned)x->S) : -1;
  const int nchunks = (int)ceil_div<int64_t>(P, ROUTE_CHUNK);
  if (nchunks > 0) {
    if (ids_i64) route_kernel<int64_t, false><<<nchunks, 256, 0, s>>>(a);
    else route_kernel<int32_t, false><<<nchunks, 256, 0, s>>>(a);
    KRS_LAUNCH_CHECK();
  }
  route_scan_kernel<<<1, 32 * MAXS, 0, s>>>(a.counts, nchunks, x->S, a.hdr);
  KRS_LAUNCH_CHECK();
  if (nchunks > 0) {
    if (ids_i64) route_kernel<int64_t, true><<<nchunks, 256, 0, s>>>(a);
    else route_kernel<int32_t, true><<<nchunks, 256, 0, s>>>(a);
    KRS_LAUNCH_CHECK();
  }
  return KRS_OK;
}

extern "C" int krs_xchg_barrier(const krs_xchg_t* x, uint32_t epoch, double timeout_s, void* stream) {
  int rc = check_xchg(x);
  if (rc) return rc;
  KRS_REQUIRE(timeout_s > 0.0, "krs_xchg_barrier: timeout must be positive");
  xchg_barrier_kernel<<<1, 32, 0, as_stream(stream)>>>(peers_of(x), x->off_flags, x->S, x->me, epoch,
                                                       (unsigned long long)(timeout_s * 1e9));
  KRS_LAUNCH_CHECK();
  return KRS_OK;
}

#define KRS_LPR_SWITCH(LPRV, CALL) \
  switch (LPRV) {                  \
    case 1: { constexpr int L = 1; CALL; } break;   \
    case 2: { constexpr int L = 2; CALL; } break;   \
    case 4: { constexpr int L = 4; CALL; } break;   \
    case 8: { constexpr int L = 8; CALL; } break;   \
    case 16: { constexpr int L = 16; CALL; } break; \
    default: { constexpr int L = 32; CALL; } break; \
  }

extern "C" int krs_xchg_gather_push(const krs_xchg_t* x, int parity, const float* arena, uint32_t* touched, void* stream) {
  int rc = check_xchg(x);
  if (rc) return rc;
  KRS_REQUIRE(arena != nullptr && aligned16(arena), "krs_xchg_gather_push: the local shard must be 16-byte aligned");
  KRS_REQUIRE(parity == 0 || parity == 1, "krs_xchg_gather_push: parity must be 0 or 1");
  OwnerArgs a = owner_args(x, parity);
  a.arena = arena;
  a.touched = touched;
  const unsigned grid = (unsigned)sm_count() * 8;
  KRS_LPR_SWITCH(lanes_per_row(x->E), (gather_push_kernel<L><<<grid, 256, 0, as_stream(stream)>>>(a)));
  KRS_LAUNCH_CHECK();
  return KRS_OK;
}

extern "C" size_t krs_slot_scan_blocks(int64_t nrows) { return (size_t)ceil_div<int64_t>(ceil_div<int64_t>(nrows, 32), 1024); }

extern "C" int krs_slot_scan(const uint32_t* touched, int64_t nwords, uint32_t* wordprefix, uint32_t* blockbase,
                             uint32_t* n_unique, void* stream) {
  KRS_REQUIRE(touched && wordprefix && blockbase && n_unique && nwords > 0, "krs_slot_scan: bad argument");
  const int64_t nblocks = ceil_div<int64_t>(nwords, 1024);
  KRS_REQUIRE(nblocks < 0x7fffffffLL, "krs_slot_scan: bitmap too large");
  bitmap_block_scan_kernel<<<(unsigned)nblocks, 256, 0, as_stream(stream)>>>(touched, nwords, wordprefix, blockbase);
  KRS_LAUNCH_CHECK();
  bitmap_base_scan_kernel<<<1, 1024, 0, as_stream(stream)>>>(blockbase, (int)nblocks, n_unique);
  KRS_LAUNCH_CHECK();
  return KRS_OK;
}

extern "C" int krs_xchg_grad_pull(const krs_xchg_t* x, int parity, const uint32_t* touched, const uint32_t* wordprefix,
                                  const uint32_t* blockbase, float* compact, int32_t* uniq_rows, int64_t cap_rows,
                                  void* stream) {
  int rc = check_xchg(x);
  if (rc) return rc;
  KRS_REQUIRE(touched && wordprefix && blockbase && compact && uniq_rows && cap_rows > 0, "krs_xchg_grad_pull: null argument");
  KRS_REQUIRE(aligned16(compact), "krs_xchg_grad_pull: compact buffer must be 16-byte aligned");
  KRS_REQUIRE(parity == 0 || parity == 1, "krs_xchg_grad_pull: parity must be 0 or 1");
  OwnerArgs a = owner_args(x, parity);
  a.touched = const_cast<uint32_t*>(touched);
  a.wordprefix = wordprefix;
  a.blockbase = blockbase;
  a.compact = compact;
  a.uniq_rows = uniq_rows;
  a.cap_rows = cap_rows;
  const unsigned grid = (unsigned)sm_count() * 8;
  KRS_LPR_SWITCH(lanes_per_row(x->E), (grad_pull_kernel<L><<<grid, 256, 0, as_stream(stream)>>>(a)));
  KRS_LAUNCH_CHECK();
  return KRS_OK;
}
