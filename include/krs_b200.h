/*
 * krs_b200.h — C ABI of libkrs_b200.so (sm_100a).
 *
 * This is the drop-in boundary for the keras-rs hot path
 *   Embedding gather -> FeatureCross | DotInteraction -> Dense stack (+ BruteForceRetrieval top-k)
 * The reference (keras-team/keras-rs) has no FFI of its own: its boundary is the Keras Layer
 * protocol and every FLOP is delegated to keras.ops.  Each entry point below
 * names the reference call site it replaces (file:line under /root/reference).
 *
 * Conventions
 *  - extern "C", plain pointers and sizes; no torch / C++ types cross the boundary.
 *  - every pointer is a DEVICE pointer unless the parameter is documented "host".
 *  - all matrices are dense row-major fp32; `ld*` are leading dimensions in ELEMENTS.
 *  - the caller owns every buffer (including workspaces); the library never allocates or frees
 *    device memory on the hot path (krs_ipc_* setup helpers are the only allocators).
 *  - calls are asynchronous: work is enqueued on `stream` (a cudaStream_t passed as void*).
 *  - return value: 0 = KRS_OK, negative = error; text via krs_last_error() (thread-local).
 *  - functions are re-entrant; no global mutable state except the thread-local error string
 *    and one-time attribute setup (cudaFuncSetAttribute) guarded by std::call_once.
 */
#ifndef KRS_B200_H_
#define KRS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KRS_OK 0
#define KRS_EINVAL (-1)   /* bad argument (shape / alignment / enum) */
#define KRS_ECUDA (-2)    /* CUDA runtime error, see krs_last_error() */
#define KRS_EUNSUPPORTED (-3)
#define KRS_ENCCL (-4)

/* ------------------------------------------------------------------ misc */
int krs_version(void);                 /* 10000*major + 100*minor + patch */
const char* krs_last_error(void);      /* thread-local, never NULL */
int krs_device_sm_count(void);         /* SMs of the current device (148 on B200) */
/* Which GEMM engine the dense contractions use: 0 = fp32 FFMA (exact fp32 products),
 * 1 = tcgen05 3xTF32 split (tensor pipe, fp32-level accuracy).  Process-wide default. */
int krs_set_gemm_engine(int engine);
int krs_get_gemm_engine(void);
/* Number of tcgen05 GEMM launches so far in this process (tests use it to prove the tensor-pipe
 * kernel, not the FFMA fallback, produced a result). */
long long krs_gemm_tc_launch_count(void);
/* Debug: device buffer of >= 16001 uint64 receiving a per-role clock64 timeline of CTA 0 of subsequent
 * tcgen05 GEMM launches ([0] = entry count, then (tag<<32|index, clock) pairs); NULL disables. */
int krs_gemm_tc_set_trace(void* dev_buf);
/* Optional caller-owned device scratch (16-byte aligned) for the tcgen05 engines: when a GEMM's B operand is a small
 * matrix re-read by many row tiles (the weights of FeatureCross / Dense), its low-order TF32 plane is computed once
 * per call into this buffer and streamed by TMA instead of being re-derived per tile.  NULL / 0 unregisters.  The
 * buffer must outlive every GEMM call, and calls that use it must be issued on one stream at a time. */
int krs_gemm_set_workspace(void* dev_buf, size_t bytes);
/* Number of split_lo_kernel launches (the precomputed low-order plane above) so far in this process. */
long long krs_gemm_split_launch_count(void);

/* ------------------------------------------------------------------ activations
 * keras.activations used as FeatureCross.pre_activation / Dense.activation
 * (feature_cross.py:115, examples/dcn.py:445, examples/ml_perf/model.py:214-266). */
enum { KRS_ACT_LINEAR = 0, KRS_ACT_RELU = 1, KRS_ACT_SIGMOID = 2, KRS_ACT_TANH = 3, KRS_ACT_SWISH = 4 };

/* ------------------------------------------------------------------ embedding gather
 * Replaces keras.layers.Embedding.call == ops.take(table, ids, axis=0) at examples/dcn.py:430-435,
 * EmbedReduce.call (embed_reduce.py:162-274: weights, sum over axis -2, mean/sum/sqrtn divisors),
 * the per-feature loop of DistributedEmbedding._default_device_call
 * (base_distributed_embedding.py:910-928) and the following ops.concatenate
 * (examples/dcn.py:437, examples/ml_perf/model.py:204-207): ONE launch for all features, output
 * written directly in the concatenated (B, out_ld) layout. */
enum { KRS_COMBINER_SUM = 0, KRS_COMBINER_MEAN = 1, KRS_COMBINER_SQRTN = 2 };

typedef struct krs_feature {
  const float* table;     /* (vocab, dim) fp32 row-major; 16-byte aligned when dim % 4 == 0     */
  const void* ids;        /* int32 or int64 ids; element (b,h) at ids[b*ids_stride + h]          */
  const float* weights;   /* NULL or per-id weights, same indexing as ids                        */
  float* grad;            /* bwd only: dense (vocab, dim) gradient arena, accumulated into       */
  uint32_t* touched;      /* bwd only: NULL or bitmap of ceil(vocab/32) words, bit r set when    */
                          /*           row r received gradient                                    */
  int64_t vocab;          /* rows in table.  ids < 0 count from the end; ids still outside        */
                          /* [0, vocab) address no row: NaN row forward, no gradient backward    */
                          /* (jnp.take mode="fill", the JAX backend of keras.ops.take)           */
  int64_t ids_stride;     /* elements between consecutive samples                                */
  int32_t hotness;        /* H: ids per sample (1 for rank-1 ids)                                */
  int32_t dim;            /* E                                                                   */
  int32_t out_offset;     /* first output column of this feature                                 */
  int32_t combiner;       /* KRS_COMBINER_*                                                      */
  int32_t ids_i64;        /* 1 = int64 ids, 0 = int32                                            */
  int32_t reduce;         /* 1 = ids were rank 2 (reduce over H, apply divisor);                 */
                          /* 0 = rank 1: no reduction, weights only honoured for `sum`           */
                          /*     (embed_reduce.py:224)                                           */
  /* row-sharded (MOD) tables, config C5 / SURVEY F6: when num_shards > 1, `shard_tables[s]`     */
  /* (device array of num_shards pointers, possibly peer-mapped) holds rows r with r%S==s at     */
  /* local row r/S (tensorflow/distributed_embedding.py:316-328).  `table` is ignored.           */
  const float* const* shard_tables;
  float* const* shard_grads;
  uint32_t* const* shard_touched; /* bwd only, nullable: per-shard bitmaps indexed by LOCAL row     */
  int32_t num_shards;
  int32_t shard_mode;     /* 0 KRS_SHARD_DIRECT: row = shard_tables[id%S] + (id/S)*E (peer-mapped tables read  */
                          /* in place; random remote rows are slow — the training path is krs_xchg_* below)   */
} krs_feature_t;
#define KRS_SHARD_DIRECT 0

/* features: HOST array of F descriptors (copied into kernel parameters; F <= 96 per call).
 * out: (B, out_ld) fp32. */
int krs_gather_fwd(const krs_feature_t* features, int F, int64_t B, float* out, int64_t out_ld,
                   int variant, void* stream);
/* Backward of the above (a10; oracle jax/test_utils.py:395-417 `grad[col] += w * g[row]`):
 * scatter-add of gout (B, gout_ld) into each feature's dense `grad` arena, duplicate ids inside a
 * warp combined with shuffles before one vector atomic per row; sets `touched` bits. */
int krs_gather_bwd(const krs_feature_t* features, int F, int64_t B, const float* gout,
                   int64_t gout_ld, void* stream);

/* ------------------------------------------------------------------ FeatureCross (DCN-v2)
 * Replaces FeatureCross.call, feature_cross.py:155-194:
 *     h = x            (projection_dim None)      | x @ U      (D,P)
 *     z = h @ V + b ;  a = pre_activation(z) ;  h2 = a + diag_scale * x ;  y = x0 * h2 + x
 * x0, x, y, h2: (B, D) row-major, ld = D.  U: (D,P) or NULL.  V: (P or D, D).  b: (D) or NULL.
 * h2_out (nullable) is saved for the backward; hproj (B,P) workspace/saved, required when U given.
 * z_out (nullable) receives the pre-activation, required for non-linear activations in training. */
int krs_cross_fwd(const float* x0, const float* x, const float* U, const float* V, const float* b,
                  float diag_scale, int act, float* y, float* h2_out, float* z_out, float* hproj,
                  int64_t B, int D, int P, void* stream);
/* Backward (analytic).  Outputs: dx0, dx (B,D) [dx0 may alias nothing; if same_input
 * the caller adds them], dU (D,P) nullable, dV, db nullable.  dz (B,D) and dh (B,P) are caller
 * workspaces.  dV/dU/db are OVERWRITTEN. */
#define KRS_CROSS_ACC_DX0 1     /* dx0 += gy*h2 instead of dx0 = gy*h2 (stacked layers share x0)   */
#define KRS_CROSS_SAME_INPUT 2  /* x is x0 (call(x0) with x=None): dx receives the TOTAL gradient    */
                                /* w.r.t. x0 (incl. whatever dx0 already holds when ACC_DX0 is set);  */
                                /* dx0 is left holding scratch                                        */
int krs_cross_bwd(const float* gy, const float* x0, const float* x, const float* U, const float* V,
                  const float* h2, const float* z, const float* hproj, float diag_scale, int act,
                  float* dx0, float* dx, float* dU, float* dV, float* db, float* dz, float* dh,
                  int64_t B, int D, int P, int flags, void* stream);
/* Elementwise tail only (used when pre_activation is a user callable evaluated by the caller):
 * y = x0 * (a + diag*x) + x ; and its backward. */
int krs_cross_combine_fwd(const float* x0, const float* x, const float* a, float diag_scale,
                          float* y, int64_t n, void* stream);
int krs_cross_combine_bwd(const float* gy, const float* x0, const float* x, const float* a,
                          float diag_scale, float* dx0, float* dx, float* da, int64_t n,
                          void* stream);

/* ------------------------------------------------------------------ Dense
 * Replaces keras.layers.Dense: y = act(x @ W + b), W is (in, out) (examples/dcn.py:444-447,
 * examples/ml_perf/model.py:214-266). */
int krs_dense_fwd(const float* x, const float* W, const float* b, int act, float* y, int64_t B,
                  int K, int N, void* stream);
/* dz workspace (B,N).  dx nullable (first layer).  dW, db overwritten. */
int krs_dense_bwd(const float* gy, const float* x, const float* W, const float* y, int act,
                  float* dx, float* dW, float* db, float* dz, int64_t B, int K, int N,
                  void* stream);

/* Plain GEMM, exposed for tests:  C(M,N) = op(A) @ op(B) [+ C if accumulate].
 * transA: A stored (K,M); transB: B stored (N,K). */
int krs_sgemm(const float* A, const float* Bm, float* C, int64_t M, int64_t N, int64_t K, int transA,
              int transB, int accumulate, void* stream);

/* ------------------------------------------------------------------ DotInteraction (DLRM)
 * Replaces DotInteraction.call, dot_interaction.py:134-205 (stack + bmm + tril take / mask).
 * feats: HOST array of N device pointers, feature i element (b,e) at feats[i][b*strides[i] + e]
 * (so slices of a concatenated buffer need no stack copy).  out: (B, out_dim) with
 * out_dim = N(N-1)/2, N(N+1)/2 (self_interaction) or N*N (skip_gather; upper part exact 0). */
int krs_dot_fwd(const float* const* feats, const int64_t* strides, int N, int E, int64_t B,
                int self_interaction, int skip_gather, float* out, void* stream);
/* dfeats: HOST array of N device pointers with dstrides; dF = (G + G^T) F on the selected set. */
int krs_dot_bwd(const float* const* feats, const int64_t* strides, const float* gout,
                float* const* dfeats, const int64_t* dstrides, int N, int E, int64_t B,
                int self_interaction, int skip_gather, void* stream);

/* ------------------------------------------------------------------ BruteForceRetrieval
 * Replaces Retrieval.compute_score (retrieval.py:101-117) + keras.ops.top_k + ops.take
 * (brute_force_retrieval.py:139-143): scores = Q @ C^T are streamed tile by tile and never
 * materialised; exact top-k, sorted descending, ties -> lowest candidate index first.
 * Q (nq,d), C (nc,d), cand_ids nullable int32 (nc).  top_scores (nq,k) fp32, top_ids (nq,k) int32.
 * workspace: krs_topk_workspace_bytes(). */
size_t krs_topk_workspace_bytes(int64_t nq, int64_t nc, int d, int k);
/* Score engine of krs_topk: 0 = auto (tensor pipe when the problem fills the machine and the shape is eligible:
 * d % 4 == 0, d <= 64, k <= 128, nc >= 96), 1 = exact-fp32 FMA score tiles only, 2 = tcgen05 (3xTF32, fp32-level
 * accuracy) whenever eligible.  krs_topk_tc_launch_count() lets tests prove which kernel produced a result. */
int krs_set_topk_engine(int engine);
long long krs_topk_tc_launch_count(void);
int krs_topk(const float* Q, const float* C, const int32_t* cand_ids, float* top_scores,
             int32_t* top_ids, int64_t nq, int64_t nc, int d, int k, void* workspace,
             size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------ row-wise selection / masking (retrieval training)
 * krs_row_topk: exact top-k of every row of x (rows, n), ld elements between rows, n <= 16384; values descending, ties ->
 * lowest index first.  boost (nullable): the ordering key is x + boost * boost_scale while out_vals stay the unboosted
 * x — HardNegativeMining (hard_negative_mining.py:43-94: top-(num_hard_negatives + 1) of logits + labels * MAX_FLOAT,
 * then take_along_axis of logits and of labels = gather2 -> out_gather2).  Also merges the per-shard result lists of the
 * candidate-sharded BruteForceRetrieval (multi-GPU caller: examples/data_parallel_retrieval.py:145-165). */
int krs_row_topk(const float* x, int64_t rows, int n, int64_t ld, const float* boost, int64_t boost_ld, float boost_scale,
                 int k, float* out_vals, int32_t* out_idx, const float* gather2, int64_t gather2_ld, float* out_gather2,
                 const int32_t* gather_i32 /* nullable int32 payload, e.g. candidate ids */, int64_t gather_i32_ld,
                 int32_t* out_gather_i32, void* stream);
/* Backward of a row selection: dst (rows, n) = 0, dst[r, idx[r, j]] = g[r, j]. */
int krs_row_scatter(const float* g, const int32_t* idx, int64_t rows, int k, int n, float* dst, void* stream);
/* RemoveAccidentalHits.call (remove_accidental_hits.py:84-97), literally: positive index = argmax(labels[r]), positive
 * id = take(FLATTENED candidate_ids, positive index), out = logits + ((ids == positive id) - labels) * smallest.
 * candidate_ids: int32 / int64, one row of n ids shared by every row (ids_per_row = 0) or (rows, n) (ids_per_row = 1). */
int krs_remove_accidental_hits(const float* logits, const float* labels, const void* candidate_ids, int ids_i64,
                               int ids_per_row, int64_t rows, int n, float smallest, float* out, void* stream);
/* SamplingProbabilityCorrection.call (sampling_probability_correction.py:39-58): out = logits - log(clip(p, eps, 1)),
 * probs broadcast with period probs_period over the flattened logits. */
int krs_sampling_prob_correction(const float* logits, const float* probs, int64_t total, int64_t probs_period, float eps,
                                 float* out, void* stream);

/* ------------------------------------------------------------------ losses
 * keras.losses.MeanSquaredError / BinaryCrossentropy(from_logits=False) on (B,1) predictions
 * (examples/dcn.py:128, examples/ml_perf/main.py:201-210): loss scalar (mean over B) and
 * dpred = dloss/dpred.  kind: 0 = MSE, 1 = BCE on probabilities, 2 = BCE with logits. */
/* denom: normalisation count (<= 0 -> B).  Data-parallel ranks pass the GLOBAL batch so that the
 * all-reduced loss / gradients are the global-batch mean. */
int krs_loss_fwd_bwd(const float* pred, const float* label, float* loss, float* dpred, int64_t B,
                     int kind, int64_t denom, void* stream);

/* ------------------------------------------------------------------ optimizers (Keras 3 formulas)
 * AdamW (examples/dcn.py:127):   p -= lr*wd*p ; m += (1-b1)(g-m) ; v += (1-b2)(g*g-v) ;
 *                                p -= lr*sqrt(1-b2^t)/(1-b1^t) * m / (sqrt(v)+eps)
 * If `touched` is non-NULL, g is a gradient ARENA: rows whose bit is clear are treated as g = 0
 * without reading them, set rows are read, zeroed and their bit cleared (row_len = floats per row).
 * If touched is NULL, g is read densely and left untouched. */
int krs_adamw(float* p, float* m, float* v, float* g, uint32_t* touched, int64_t n, int row_len,
              float lr, float b1, float b2, float eps, float wd, int64_t step,
              const float* hyper_dev /* nullable device [lr,b1,b2,eps,wd,alpha,step]: overrides the scalars */,
              void* stream);
/* Same update with an extra persistent bitmap `ever` (one bit per row, zero-initialised by the caller, nullable): rows that
 * have never received a gradient still hold m = v = 0 exactly, so for them the rule reduces bit for bit to the decoupled
 * decay of p alone and the sweep moves 8 instead of 24 bytes per parameter; `ever |= touched` is folded in after the sweep. */
int krs_adamw_cold(float* p, float* m, float* v, float* g, uint32_t* touched, uint32_t* ever, int64_t n, int row_len,
                   float lr, float b1, float b2, float eps, float wd, int64_t step, const float* hyper_dev, void* stream);
/* Advances hyper_dev[6] (step) and refreshes hyper_dev[5] (alpha) on the device, so a captured CUDA graph of
 * the training step can be replayed without per-step host parameters. */
int krs_adam_hyper_advance(float* hyper_dev, void* stream);
/* Adagrad (examples/ml_perf/main.py:203): acc += g*g ; p -= lr * g / sqrt(acc + eps).
 * SGD: p -= lr * g.  kind: 0 = SGD, 1 = Adagrad.  Same `touched` arena semantics; with an arena only
 * touched rows are visited (row-sparse update, identical to the dense update for these rules —
 * the precedent is jax/embedding_lookup.py:174-273, oracle jax/test_utils.py:474-497). */
int krs_sgd_adagrad(float* p, float* acc, float* g, uint32_t* touched, int64_t n, int row_len,
                    float lr, float eps, int kind, void* stream);

/* ------------------------------------------------------------------ MOD shard routing (C5)
 * owner[i] = ids[i] % S, local[i] = ids[i] / S  (jax/embedding_utils.py:187-197 "MOD";
 * tensorflow/distributed_embedding.py:316-328) and per-owner counts (S ints, pre-zeroed). */
int krs_mod_route(const void* ids, int ids_i64, int64_t n, int S, int32_t* owner, int64_t* local,
                  int32_t* counts, void* stream);

/* ------------------------------------------------------------------ row-sharded exchange over peer memory (C5)
 * The B200 form of the reference's SparseCore protocol (ids routed to the owner, owner-side lookup, activations
 * returned, optimizer applied on the owner without a dense gradient: jax/embedding_utils.py:144-217,
 * jax/embedding_lookup.py:174-273).  Every rank owns one cudaIpc-exported REGION that all peers map; the caller
 * chooses the layout (byte offsets below, identical on every rank):
 *   flags  : uint32[KRS_XCHG_MAX_SHARDS + 1]   barrier epochs written by the peers + one error word
 *   hdr    : int32[2][KRS_XCHG_MAX_SHARDS + 1] bucket starts of the request lists, double buffered by step parity
 *   rows   : int32[2][B*F]                     request lists: arena row ON THE OWNER, bucket o at hdr[o]
 *   pos    : int32[2][B*F]                     position b*F+f of that request in the requester's activation
 *   x0     : float[B*F*E]                      the requester's concatenated activation (owners write into it)
 *   grad   : float[B*F*E]                      dL/dx0 (owners read from it)
 * ids < 0 count from the end of the table, ids still outside [0, vocab) have no owner: their activation row is NaN
 * and they receive no gradient (jnp.take mode="fill", the JAX backend of keras.ops.take). */
#define KRS_XCHG_MAX_SHARDS 16
typedef struct krs_xchg {
  int32_t S, me, F, E;
  int64_t B;                                  /* samples per rank */
  void* peer_base[KRS_XCHG_MAX_SHARDS];       /* region of rank s as mapped in THIS process ([me] = local)      */
  int64_t off_flags, off_hdr, off_rows, off_pos, off_x0, off_grad;
} krs_xchg_t;
size_t krs_xchg_route_workspace_bytes(int64_t B, int F, int S);
/* Requester: stable counting sort of the local ids (B, ids_ld) by owner into rows/pos/hdr[parity] of my region.
 * vocab_dev: device int64[F] GLOBAL vocabulary sizes; owner_row_off_dev: device int32[S*F], first arena row of
 * table f on owner o.  fill_invalid_nan: write NaN rows into my x0 for ids without an owner. */
int krs_xchg_route(const krs_xchg_t* x, int parity, const void* ids, int ids_i64, int64_t ids_ld,
                   const int64_t* vocab_dev, const int32_t* owner_row_off_dev, int32_t* workspace,
                   int fill_invalid_nan, void* stream);
/* All ranks: stream-ordered barrier through the flag arrays (epoch must increase by one per barrier, same sequence
 * on every rank).  A peer that does not arrive within timeout_s sets bit p of the error word instead of hanging. */
int krs_xchg_barrier(const krs_xchg_t* x, uint32_t epoch, double timeout_s, void* stream);
/* Owner: ONE launch serving every requester's bucket for this rank: rows of the local shard `arena` (rows, E) are
 * written into the requesters' x0.  touched (nullable): bit per local arena row, set for every requested row. */
int krs_xchg_gather_push(const krs_xchg_t* x, int parity, const float* arena, uint32_t* touched, void* stream);
/* Slot numbering of the touched rows: slot(row) = blockbase[row >> 15] + wordprefix[row >> 5] + popc(bits below).
 * wordprefix: uint32[nwords]; blockbase: uint32[krs_slot_scan_blocks(rows)]; n_unique: device scalar. */
size_t krs_slot_scan_blocks(int64_t nrows);
int krs_slot_scan(const uint32_t* touched, int64_t nwords, uint32_t* wordprefix, uint32_t* blockbase,
                  uint32_t* n_unique, void* stream);
/* Owner: ONE launch pulling the gradient rows of every requester's bucket from the peers' `grad` buffers and
 * accumulating them (duplicates combined) into compact (cap_rows, E), which must be zero on entry; uniq_rows[slot]
 * receives the arena row.  More distinct rows than cap_rows sets bit 31 of the error word. */
int krs_xchg_grad_pull(const krs_xchg_t* x, int parity, const uint32_t* touched, const uint32_t* wordprefix,
                       const uint32_t* blockbase, float* compact, int32_t* uniq_rows, int64_t cap_rows, void* stream);
/* Row-sparse optimizers fused onto the compact gradient rows (TableConfig.optimizer, base_distributed_embedding.py:
 * 176-186; the supported set is the reference's, jax/config_conversion.py:211-288; oracle jax/test_utils.py:474-497):
 * only rows that received gradient are visited — for SGD / Adagrad that IS the dense rule (g = 0 changes nothing);
 * Adam and FTRL are the per-row ("lazy") forms the SparseCore path applies.  For slot < *n_unique: row = uniq_rows[slot],
 * g = compact[slot]; p (and the slot variables s1, s2) are updated in place, compact[slot] is re-zeroed.
 * hyper (host, 8 floats): SGD [lr] | Adagrad [lr, eps] | Adam [lr, b1, b2, eps, alpha] (alpha = lr*sqrt(1-b2^t)/
 * (1-b1^t)) | FTRL [lr, lr_power, l1, l2, beta] (keras Ftrl with l2_shrinkage = 0). */
enum { KRS_OPT_SGD = 0, KRS_OPT_ADAGRAD = 1, KRS_OPT_ADAM = 2, KRS_OPT_FTRL = 3 };
int krs_rows_apply(float* p, float* s1, float* s2, float* compact, const int32_t* uniq_rows, const uint32_t* n_unique,
                   int64_t cap_rows, int E, int kind, const float* hyper, uint32_t* touched_to_clear /*nullable*/,
                   int64_t nwords, void* stream);
/* The same four rules from a gradient ARENA + touched bitmap (krs_gather_bwd's output; only touched rows are visited, the
 * arena rows are re-zeroed, the bitmap cleared) or, with touched == NULL, from a dense gradient over all n elements. */
int krs_opt_apply(float* p, float* s1, float* s2, float* g, uint32_t* touched /*nullable*/, int64_t n, int row_len, int kind,
                  const float* hyper, void* stream);
/* Dense-semantics AdamW (every row decays, examples/dcn.py:127) reading its gradient from the compact rows:
 * rows whose touched bit is clear have g = 0; compact rows are re-zeroed and the bitmap is cleared afterwards.
 * ever (nullable): the ever-touched bitmap of krs_adamw_cold, same meaning (ever |= touched is folded in). */
int krs_adamw_compact(float* p, float* m, float* v, float* compact, uint32_t* touched, const uint32_t* wordprefix,
                      const uint32_t* blockbase, uint32_t* ever, int64_t n, int row_len, float lr, float b1, float b2,
                      float eps, float wd, int64_t step, void* stream);

/* ------------------------------------------------------------------ peer memory (setup only)
 * Row-sharded tables live in cudaMalloc'd arenas exported with cudaIpc so that the fused gather
 * can read (and the backward can atomically add to) peer shards directly over NVLink. */
int krs_ipc_alloc(void** dptr, size_t bytes, void* handle_out_64B /*host*/);
int krs_ipc_open(const void* handle_64B /*host*/, void** dptr);
int krs_ipc_close(void* dptr);
int krs_ipc_free(void* dptr);
int krs_enable_peer_access(int peer_device);

/* (NCCL is used where a collective is really needed — the all-reduce of the dense gradients — through the caller's
 * torch.distributed communicator; the embedding exchange itself is the krs_xchg_* kernels above.  The NCCL all-to-all
 * of the same bytes is measured as the baseline by bench.py's phase breakdown.) */

#ifdef __cplusplus
}
#endif
#endif /* KRS_B200_H_ */
