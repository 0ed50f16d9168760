"""CPU ORACLE — test infrastructure only, never the product path.

A numpy (float32) restatement of the keras-rs hot path
    Embedding gather -> FeatureCross | DotInteraction -> Dense stack (+ BruteForceRetrieval)
following the reference sources cited per function (paths relative to /root/reference).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this package.  The product (keras_rs_b200) never does; it fails loudly when the CUDA
extension is missing.

Pinning: the reference itself cannot be imported here (every module does `import keras`; keras /
jax / tensorflow are not installed and there are no wheels — SURVEY.md F3), so this restatement is
pinned against the closed-form golden vectors the reference's OWN tests hold for this path
(SURVEY.md §8c), transcribed into tests/golden/*.json by tests/golden/make_golden.py:
  feature_cross_test.py:15-79, dot_interaction_test.py:17-91, embed_reduce_test.py:45-119,
  brute_force_retrieval_test.py:13-64, retrieval_test.py:21-48,
  embedding/test_utils.py:245-267, embedding/jax/test_utils.py:395-417,474-497.
Forward parity is pinned at the keras_rs layer boundary; backward formulas are derived analytically
(the reference has no gradient tests for these layers) and cross-checked against torch-CPU autograd
in tests/test_oracle_grad.py.

The arithmetic of keras.ops.{matmul,take,top_k,...} lives in the third-party `keras` package
(unpinned: pyproject.toml:29-32).  Semantics assumed: Dense y = act(x @ kernel + bias) with kernel
(in, out); Embedding = take(table, ids, axis=0); top_k -> (values sorted descending, int32 indices),
ties broken lowest-index-first (jax.lax.top_k behaviour).
"""
from __future__ import annotations

import math
from typing import Sequence

import numpy as np

F32 = np.float32


# --------------------------------------------------------------------------- activations
def activation(name_or_fn, z: np.ndarray) -> np.ndarray:
    """keras.activations.get(...) for the names the hot path uses (feature_cross.py:115)."""
    if name_or_fn is None or name_or_fn == "linear":
        return z
    if callable(name_or_fn):
        return np.asarray(name_or_fn(z), dtype=z.dtype)
    if name_or_fn == "relu":
        return np.maximum(z, F32(0))
    if name_or_fn == "sigmoid":
        return (F32(1) / (F32(1) + np.exp(-z))).astype(z.dtype)
    if name_or_fn == "tanh":
        return np.tanh(z)
    if name_or_fn in ("swish", "silu"):
        return (z / (F32(1) + np.exp(-z))).astype(z.dtype)
    raise ValueError(f"unknown activation {name_or_fn!r}")


def activation_grad(name, z: np.ndarray, a: np.ndarray) -> np.ndarray:
    """d act(z) / dz given z and a = act(z)."""
    if name is None or name == "linear":
        return np.ones_like(z)
    if name == "relu":
        return (z > 0).astype(z.dtype)
    if name == "sigmoid":
        return a * (F32(1) - a)
    if name == "tanh":
        return F32(1) - a * a
    if name in ("swish", "silu"):
        s = F32(1) / (F32(1) + np.exp(-z))
        return (s * (F32(1) + z * (F32(1) - s))).astype(z.dtype)
    raise ValueError(f"unknown activation {name!r}")


# --------------------------------------------------------------------------- embedding
def embedding_lookup(table: np.ndarray, ids: np.ndarray) -> np.ndarray:
    """keras.layers.Embedding.call == ops.take(table, ids, axis=0) (examples/dcn.py:430-435;
    embed_reduce.py:178).

    Out-of-range ids follow the backend the north star names (KERAS_BACKEND=jax on CPU): keras' jax `take`
    is `jnp.take(x, indices, axis=axis)` with the default mode, which jax documents as "fill": negative
    indices count from the end (idx + n), indices still outside [0, n) return NaN for floating tables, and the
    transpose (the gradient scatter-add) drops them.  keras and jax are third-party and unpinned
    (pyproject.toml:29-32: keras >= 3.10, jax unpinned); the rule is restated from the published jnp.take
    contract.  The product follows this oracle (gather.cu resolve_id), not the other way round."""
    ids, valid = resolve_ids(ids, table.shape[0])
    out = table[np.where(valid, ids, 0)]
    if not valid.all():
        out = out.copy()
        out[~valid] = np.nan
    return out


def resolve_ids(ids, vocab: int):
    """jnp.take(mode="fill") index rule: (ids wrapped once from the end, mask of ids that address a row)."""
    ids = np.asarray(ids).astype(np.int64)
    ids = np.where(ids < 0, ids + vocab, ids)
    return ids, (ids >= 0) & (ids < vocab)


def _divide_no_nan(x: np.ndarray, d: np.ndarray) -> np.ndarray:
    out = np.zeros_like(x)
    np.divide(x, d, out=out, where=(d != 0))
    return out


def embed_reduce(table: np.ndarray, ids: np.ndarray, weights: np.ndarray | None = None,
                 combiner: str = "mean") -> np.ndarray:
    """EmbedReduce.call, embed_reduce.py:162-274 (dense inputs)."""
    if combiner not in ("mean", "sum", "sqrtn"):
        raise ValueError(f"Invalid `combiner`: '{combiner}', use one of mean, sum, sqrtn.")
    ids = np.asarray(ids)
    x = embedding_lookup(table, ids)                       # :178
    unreduced_rank = x.ndim
    if weights is not None:
        weights = np.asarray(weights)
        if weights.ndim > unreduced_rank or tuple(x.shape[: weights.ndim]) != tuple(weights.shape):
            raise ValueError(                                # :184-190
                f"The shape of `weights`: {weights.shape} is not compatible"
                f" with the shape of `inputs` after embedding: {x.shape}.")
    if weights is None or (unreduced_rank <= 2 and combiner != "sum"):   # :224
        w = np.ones(ids.shape, dtype=x.dtype)
    else:
        w = weights.astype(x.dtype)
    wx = w.reshape(w.shape + (1,) * (unreduced_rank - w.ndim))           # :244-248
    x = x * wx                                                            # :253
    if unreduced_rank <= 2:                                               # :255-257
        return x
    # sequential sum over axis -2, h = 0..H-1 (matches the kernel's accumulation order)
    acc = np.zeros(x.shape[:-2] + x.shape[-1:], dtype=x.dtype)
    for h in range(x.shape[-2]):
        acc = acc + x[..., h, :]
    wsum_axis = wx
    if combiner == "mean":                                                # :267-268
        d = np.zeros(acc.shape[:-1] + (1,), dtype=x.dtype)
        for h in range(x.shape[-2]):
            d = d + wsum_axis[..., h, :]
        return _divide_no_nan(acc, np.broadcast_to(d, acc.shape))
    if combiner == "sum":
        return acc
    d = np.zeros(acc.shape[:-1] + (1,), dtype=x.dtype)                    # sqrtn :271-274
    for h in range(x.shape[-2]):
        d = d + wsum_axis[..., h, :] * wsum_axis[..., h, :]
    return _divide_no_nan(acc, np.broadcast_to(np.sqrt(d), acc.shape))


def expected_lookup_np(sample_ids, sample_weights, table, combiner):
    """The reference's own NumPy oracle, embedding/test_utils.py:245-267 (float64 weights)."""
    batch = len(sample_ids)
    out = np.zeros((batch, table.shape[1]), dtype=table.dtype)
    for i in range(batch):
        w = np.asarray(sample_weights[i], dtype=float)
        if combiner == "mean":
            w = w / np.sum(w)
        elif combiner == "sqrtn":
            w = w / np.sqrt(np.sum(np.square(w)))
        out[i, :] += w @ table[np.asarray(sample_ids[i]), :]
    return out


def multi_table_gather(tables: Sequence[np.ndarray], feature_table: Sequence[int],
                       ids: Sequence[np.ndarray], weights: Sequence[np.ndarray | None] | None,
                       combiners: Sequence[str]) -> np.ndarray:
    """DistributedEmbedding._default_device_call (base_distributed_embedding.py:910-928): one
    EmbedReduce per table, one lookup per feature — followed by ops.concatenate(axis=1)
    (examples/dcn.py:437, examples/ml_perf/model.py:204-207)."""
    outs = []
    for f, t in enumerate(feature_table):
        w = None if weights is None else weights[f]
        outs.append(embed_reduce(tables[t], ids[f], w, combiners[t]))
    return np.concatenate(outs, axis=1)


def embedding_grad(ids: np.ndarray, weights: np.ndarray | None, vocab: int, gout: np.ndarray,
                   combiner: str = "sum", reduce: bool | None = None) -> np.ndarray:
    """Dense (vocab, dim) gradient of embed_reduce w.r.t. the table:  grad[col] += w * g[row]
    (jax/test_utils.py:395-417), with the combiner divisor folded into w."""
    ids = np.asarray(ids)
    if reduce is None:
        reduce = ids.ndim == 2
    B = ids.shape[0]
    ids2, valid = resolve_ids(ids.reshape(B, -1), vocab)       # invalid ids receive no gradient (jnp.take "fill")
    H = ids2.shape[1]
    if weights is None or ((not reduce) and combiner != "sum"):
        w = np.ones((B, H), dtype=gout.dtype)
    else:
        w = np.asarray(weights, dtype=gout.dtype).reshape(B, H)
    scale = np.ones((B, 1), dtype=gout.dtype)
    if reduce and combiner == "mean":
        d = w.sum(axis=1, keepdims=True)
        scale = _divide_no_nan(np.ones_like(d), d)
    elif reduce and combiner == "sqrtn":
        d = np.sqrt((w * w).sum(axis=1, keepdims=True))
        scale = _divide_no_nan(np.ones_like(d), d)
    grad = np.zeros((vocab, gout.shape[1]), dtype=gout.dtype)
    for h in range(H):
        ok = valid[:, h]
        np.add.at(grad, ids2[ok, h], ((w[:, h:h + 1] * scale) * gout)[ok])
    return grad


# --------------------------------------------------------------------------- FeatureCross
def feature_cross(x0: np.ndarray, x: np.ndarray | None, V: np.ndarray, b: np.ndarray | None = None,
                  U: np.ndarray | None = None, diag_scale: float | None = 0.0,
                  pre_activation=None) -> np.ndarray:
    """FeatureCross.call, feature_cross.py:155-194.  V is dense.kernel (P or D, D), U is
    down_proj_dense.kernel (D, P); weight order pinned by feature_cross_test.py:28-32,41-47."""
    if x is None:                                  # :172-173
        x = x0
    if x0.shape != x.shape:                        # :175-179
        raise ValueError("`x0` and `x` should have the same shape. Received: "
                         f"`x.shape` = {x.shape}, `x0.shape` = {x0.shape}")
    h = x if U is None else x @ U                  # :182-185
    z = h @ V                                      # :187 Dense: act(x @ kernel + bias)
    if b is not None:
        z = z + b
    a = activation(pre_activation, z)
    if diag_scale:                                 # :191-192
        a = a + F32(diag_scale) * x
    return x0 * a + x                              # :194


def feature_cross_bwd(gy, x0, x, V, b=None, U=None, diag_scale=0.0, pre_activation=None, z_for_grad=None):
    """Analytic backward of feature_cross.  Returns dict dx0, dx, dV, db, dU.
    x0 and x are treated as independent inputs (caller sums when they are the same tensor).
    z_for_grad: optional pre-activation at which the activation DERIVATIVE is evaluated.  relu'(z) is
    discontinuous at 0, so a checker comparing against an implementation whose z differs in the last
    ulp must take the derivative mask from that implementation's own z (otherwise one flipped element
    near z = 0 changes whole rows of the gradients)."""
    h = x if U is None else x @ U
    z = h @ V + (0 if b is None else b)
    a = activation(pre_activation, z)
    h2 = a + (F32(diag_scale) * x if diag_scale else 0)
    dh2 = gy * x0
    dx0 = gy * h2
    if z_for_grad is not None:
        zg = np.asarray(z_for_grad, dtype=z.dtype)
        dz = dh2 * activation_grad(pre_activation, zg, activation(pre_activation, zg))
    else:
        dz = dh2 * activation_grad(pre_activation, z, a)
    dV = h.T @ dz
    db = dz.sum(axis=0)
    dh = dz @ V.T
    dx = gy + (F32(diag_scale) * dh2 if diag_scale else 0)
    dU = None
    if U is None:
        dx = dx + dh
    else:
        dU = x.T @ dh
        dx = dx + dh @ U.T
    return dict(dx0=dx0, dx=dx, dV=dV, db=db, dU=dU)


def dcn_block(x0: np.ndarray, layers: Sequence[dict]) -> np.ndarray:
    """DCNBlock.call, examples/ml_perf/model.py:332-336; README.md:54-55:  xl = layer(x0, xl)."""
    xl = x0
    for p in layers:
        xl = feature_cross(x0, xl, **p)
    return xl


# --------------------------------------------------------------------------- DotInteraction
def tril_indices(num_features: int, self_interaction: bool) -> list[int]:
    """DotInteraction._get_lower_triangular_indices, dot_interaction.py:118-132."""
    out = []
    for i in range(num_features):
        k = i + 1 if self_interaction else i
        for j in range(k):
            out.append(i * num_features + j)
    return out


def dot_interaction(inputs: Sequence[np.ndarray], self_interaction: bool = False,
                    skip_gather: bool = False) -> np.ndarray:
    """DotInteraction.call, dot_interaction.py:134-205."""
    shape = inputs[0].shape
    for idx, t in enumerate(inputs):
        if len(shape) != 2:                          # :155-160 (only inputs[0]'s rank is tested)
            raise ValueError("All feature tensors inside `inputs` should have rank 2. "
                             f"Received rank {len(shape)} at index {idx}.")
        if tuple(t.shape) != tuple(shape):           # :162-167
            raise ValueError("All feature tensors in `inputs` should have the same shape. "
                             f"Found at least one conflict: shape = {shape} at index 0 and "
                             f"shape = {t.shape} at index {idx}.")
    feats = np.stack(inputs, axis=1)                 # :170  (B, N, E)
    B, N, _ = feats.shape
    pair = feats @ np.transpose(feats, (0, 2, 1))    # :176-178  (B, N, N)
    if skip_gather:                                  # :182-192
        mask = np.tril(np.ones((N, N), dtype=bool), k=0 if self_interaction else -1)
        return (pair * mask.astype(pair.dtype)).reshape(B, N * N)
    idx = tril_indices(N, self_interaction)          # :194-203
    return pair.reshape(B, N * N)[:, idx]


def dot_interaction_bwd(gout: np.ndarray, inputs: Sequence[np.ndarray],
                        self_interaction: bool = False, skip_gather: bool = False):
    """dF = (G + G^T) F with G the (B,N,N) gradient scattered onto the selected lower triangle."""
    feats = np.stack(inputs, axis=1)
    B, N, _ = feats.shape
    G = np.zeros((B, N * N), dtype=gout.dtype)
    if skip_gather:
        mask = np.tril(np.ones((N, N), dtype=bool), k=0 if self_interaction else -1).reshape(-1)
        G = gout * mask.astype(gout.dtype)
    else:
        G[:, tril_indices(N, self_interaction)] = gout
    G = G.reshape(B, N, N)
    dF = (G + np.transpose(G, (0, 2, 1))) @ feats
    return [dF[:, i, :] for i in range(N)]


# --------------------------------------------------------------------------- Dense stack
def dense(x: np.ndarray, W: np.ndarray, b: np.ndarray | None = None, act=None) -> np.ndarray:
    """keras.layers.Dense: act(x @ kernel + bias); kernel (in, out) (examples/dcn.py:444-447)."""
    z = x @ W
    if b is not None:
        z = z + b
    return activation(act, z)


def dense_bwd(gy, x, W, b, act, y):
    if act in (None, "linear"):
        dz = gy
    elif act == "relu":
        dz = gy * (y > 0)
    elif act == "sigmoid":
        dz = gy * y * (F32(1) - y)
    elif act == "tanh":
        dz = gy * (F32(1) - y * y)
    else:
        z = x @ W + (0 if b is None else b)
        dz = gy * activation_grad(act, z, y)
    dz = dz.astype(x.dtype)
    return dict(dx=dz @ W.T, dW=x.T @ dz, db=dz.sum(axis=0))


# --------------------------------------------------------------------------- Retrieval
def validate_candidates(candidate_embeddings, candidate_ids, k: int) -> None:
    """Retrieval._validate_candidate_embeddings_and_ids, retrieval.py:35-68 (messages pinned by
    retrieval_test.py:21-40)."""
    if candidate_embeddings is None:
        raise ValueError("`candidate_embeddings` is required.")
    if len(candidate_embeddings.shape) != 2:
        raise ValueError("`candidate_embeddings` must be a tensor of rank 2 "
                         "(num_candidates, embedding_size), received "
                         f"`candidate_embeddings` with shape {tuple(candidate_embeddings.shape)}")
    if candidate_embeddings.shape[0] < k:
        raise ValueError(f"The number of candidates provided ({candidate_embeddings.shape[0]}) is "
                         f"less than the number of candidates to retrieve (k={k}).")
    if candidate_ids is not None and candidate_ids.shape[0] != candidate_embeddings.shape[0]:
        raise ValueError("The `candidate_embeddings` and `candidate_is` tensors must have the same "
                         f"number of rows, got tensors of shape {tuple(candidate_embeddings.shape)} "
                         f"and {tuple(candidate_ids.shape)}.")


def compute_score(q: np.ndarray, c: np.ndarray) -> np.ndarray:
    """Retrieval.compute_score, retrieval.py:101-117: matmul(q, transpose(c))."""
    return q @ c.T


def top_k(scores: np.ndarray, k: int):
    """keras.ops.top_k: values sorted descending + int32 indices; ties -> lowest index first."""
    idx = np.argsort(-scores, axis=1, kind="stable")[:, :k]
    return np.take_along_axis(scores, idx, axis=1), idx.astype(np.int32)


def brute_force_retrieval(q, c, candidate_ids=None, k: int = 10, return_scores: bool = True):
    """BruteForceRetrieval.call, brute_force_retrieval.py:126-148."""
    vals, idx = top_k(compute_score(q, c), k)          # :139-140
    if candidate_ids is not None:                      # :142-143 (ids are int32, :118-123)
        idx = np.asarray(candidate_ids, dtype=np.int32)[idx]
    return (vals, idx) if return_scores else idx


# --------------------------------------------------------------------------- losses / optimizers
# ---------------------------------------------------------------------------------------------------------
# two-tower training helpers (restated for the host-side mirrors in layers/retrieval_helpers.py)
MAX_FLOAT = float(np.finfo(np.float32).max) / 100.0          # hard_negative_mining.py:9
SMALLEST_FLOAT = float(np.finfo(np.float32).tiny) / 100.0     # remove_accidental_hits.py:9 (a subnormal)


def hard_negative_mining(logits: np.ndarray, labels: np.ndarray, num_hard_negatives: int):
    """hard_negative_mining.py:70-94: top-(k+1) columns of logits + labels*MAX_FLOAT per row (order unspecified by the
    reference: sorted=False; here descending), then take_along_axis on logits and labels."""
    n = logits.shape[-1]
    num_sampled = min(num_hard_negatives + 1, n)
    boosted = logits + labels * np.float32(MAX_FLOAT)
    idx = np.argsort(-boosted, axis=-1, kind="stable")[..., :num_sampled]
    return np.take_along_axis(logits, idx, axis=-1), np.take_along_axis(labels, idx, axis=-1)


def remove_accidental_hits(logits: np.ndarray, labels: np.ndarray, candidate_ids: np.ndarray) -> np.ndarray:
    """remove_accidental_hits.py:60-97, literally (np.take without an axis = flattened ids, as keras.ops.take)."""
    if labels.shape != logits.shape:
        raise ValueError("`labels` and `logits` should have the same shape.")
    r = candidate_ids.ndim
    if tuple(labels.shape[labels.ndim - r:]) != tuple(candidate_ids.shape):
        raise ValueError("`candidate_ids` should have the same shape as the last dimensions of `labels`.")
    ids = candidate_ids.reshape((1,) * (labels.ndim - r) + candidate_ids.shape)
    pos = np.expand_dims(np.argmax(labels, axis=-1), -1)
    pos_ids = np.take(candidate_ids, pos)
    dup = (pos_ids == ids).astype(labels.dtype) - labels
    return (logits + dup.astype(logits.dtype) * np.float32(SMALLEST_FLOAT)).astype(logits.dtype)


def sampling_probability_correction(logits: np.ndarray, probs: np.ndarray, epsilon: float = 1e-6) -> np.ndarray:
    """sampling_probability_correction.py:39-58."""
    return logits - np.log(np.clip(probs.astype(logits.dtype), np.float32(epsilon), np.float32(1.0)))


def mse_loss(pred: np.ndarray, label: np.ndarray):
    """keras.losses.MeanSquaredError on (B,1) (examples/dcn.py:128): mean over batch; + dpred."""
    d = pred.reshape(-1) - label.reshape(-1)
    return F32(np.mean(d.astype(np.float64) ** 2)), (F32(2) * d / F32(d.size)).astype(pred.dtype)


def bce_loss(prob: np.ndarray, label: np.ndarray, eps: float = 1e-7):
    """keras.losses.BinaryCrossentropy(from_logits=False) (examples/ml_perf/main.py:201-210):
    probabilities clipped to [eps, 1-eps]."""
    p = np.clip(prob.reshape(-1), F32(eps), F32(1 - eps))
    y = label.reshape(-1)
    loss = -(y * np.log(p) + (F32(1) - y) * np.log(F32(1) - p))
    inside = (prob.reshape(-1) > eps) & (prob.reshape(-1) < 1 - eps)
    dp = (-(y / p) + (F32(1) - y) / (F32(1) - p)) / F32(p.size) * inside
    return F32(np.mean(loss.astype(np.float64))), dp.astype(prob.dtype)


def adamw_step(p, m, v, g, step: int, lr=0.001, b1=0.9, b2=0.999, eps=1e-7, wd=0.004):
    """Keras 3 AdamW (examples/dcn.py:127): decoupled decay then Adam with folded bias
    correction.  `step` is 1-based.  Returns new (p, m, v)."""
    p = p - p * F32(wd) * F32(lr)
    m = m + (g - m) * F32(1 - b1)
    v = v + (g * g - v) * F32(1 - b2)
    alpha = F32(lr * math.sqrt(1 - b2 ** step) / (1 - b1 ** step))
    p = p - (m * alpha) / (np.sqrt(v) + F32(eps))
    return p.astype(F32), m.astype(F32), v.astype(F32)


def adagrad_step(p, acc, g, lr=0.001, eps=1e-7):
    """Keras Adagrad (examples/ml_perf/main.py:203; same rule as jax/test_utils.py:474-497 up to
    eps): acc += g^2 ; p -= lr * g / sqrt(acc + eps)."""
    acc = acc + g * g
    p = p - F32(lr) * g / np.sqrt(acc + F32(eps))
    return p.astype(F32), acc.astype(F32)


def sgd_step(p, g, lr=0.01):
    """SGD: table - lr * grad (jax/test_utils.py:493-497)."""
    return (p - F32(lr) * g).astype(F32)


def lazy_adam_step(p, m, v, g, step: int, lr=0.001, b1=0.9, b2=0.999, eps=1e-7, rows=None):
    """keras Adam applied only to the rows that received gradient (the per-row form of the reference's SparseCore
    path: jax/config_conversion.py:259-268 maps keras Adam onto embedding_spec.AdamOptimizerSpec, whose update visits
    the looked-up rows only).  `rows`: boolean mask of the looked-up rows (default: rows with a non-zero gradient);
    all other rows keep p, m and v."""
    rows = np.any(g != 0, axis=-1) if rows is None else np.asarray(rows, bool)
    p2, m2, v2 = adamw_step(p[rows], m[rows], v[rows], g[rows], step, lr=lr, b1=b1, b2=b2, eps=eps, wd=0.0)
    p, m, v = p.copy(), m.copy(), v.copy()
    p[rows], m[rows], v[rows] = p2, m2, v2
    return p, m, v


def ftrl_step(p, accum, linear, g, lr=0.001, lr_power=-0.5, l1=0.0, l2=0.0, beta=0.0, rows=None):
    """keras Ftrl.update_step (third-party keras/src/optimizers/ftrl.py, restated; l2_shrinkage = 0 is the only form
    jax/config_conversion.py:269-285 accepts), applied to the looked-up rows (`rows` mask; default g != 0)."""
    rows = np.any(g != 0, axis=-1) if rows is None else np.asarray(rows, bool)
    pr, ar, lr_, gr = p[rows], accum[rows], linear[rows], g[rows]
    l2r = F32(l2) + F32(beta) / (F32(2.0) * F32(lr))
    na = ar + gr * gr
    pa, pn = np.power(ar, F32(-lr_power)), np.power(na, F32(-lr_power))
    lin = lr_ + (gr - (pn - pa) / F32(lr) * pr)
    quad = pn / F32(lr) + F32(2.0) * l2r
    lc = np.clip(lin, F32(-l1), F32(l1))
    p, accum, linear = p.copy(), accum.copy(), linear.copy()
    p[rows], accum[rows], linear[rows] = ((lc - lin) / quad).astype(F32), na.astype(F32), lin.astype(F32)
    return p, accum, linear


# --------------------------------------------------------------------------- MOD sharding (C5)
def route_requests(ids: np.ndarray, vocab: Sequence[int], num_shards: int, owner_row_off):
    """Request lists of one rank's (B, F) id matrix (csrc/exchange.cu route): for every owner o, in position order, the
    pairs (arena row on o, position b*F+f) of the ids o owns; ids that address no row (embedding_lookup) are left out.
    owner_row_off[o][f] = first arena row of table f on owner o.  MOD layout as mod_route."""
    ids = np.asarray(ids).astype(np.int64)
    B, F = ids.shape
    rows = [[] for _ in range(num_shards)]
    pos = [[] for _ in range(num_shards)]
    for b in range(B):
        for f in range(F):
            i = int(ids[b, f])
            if i < 0:
                i += int(vocab[f])
            if not 0 <= i < int(vocab[f]):
                continue
            o = i % num_shards
            rows[o].append(int(owner_row_off[o][f]) + i // num_shards)
            pos[o].append(b * F + f)
    return [np.asarray(r, np.int32) for r in rows], [np.asarray(q, np.int32) for q in pos]


def mod_route(ids: np.ndarray, num_shards: int):
    """Row-wise MOD sharding: row r lives on shard r % S at local row r // S
    (jax/embedding_utils.py:187-197; tensorflow/distributed_embedding.py:316-328)."""
    ids = np.asarray(ids).astype(np.int64)
    return (ids % num_shards).astype(np.int32), ids // num_shards


def mod_shard_table(table: np.ndarray, num_shards: int) -> list[np.ndarray]:
    return [np.ascontiguousarray(table[s::num_shards]) for s in range(num_shards)]


def mod_unshard_table(shards: Sequence[np.ndarray]) -> np.ndarray:
    S = len(shards)
    V = sum(s.shape[0] for s in shards)
    out = np.empty((V, shards[0].shape[1]), dtype=shards[0].dtype)
    for s in range(S):
        out[s::S] = shards[s]
    return out


# --------------------------------------------------------------------------- DCN model (C1/C2)
def dcn_forward(params: dict, ids: np.ndarray, cache: dict | None = None) -> np.ndarray:
    """examples/dcn.py:418-449 wiring with the stacked cross of README.md:54-55:
    per-feature Embedding -> concatenate -> FeatureCross x L -> Dense(relu)... -> Dense(1).
    params: tables [F x (V,E)], cross [L x dict(V,b,U?)], mlp [(W,b,act)...]. ids (B,F)."""
    embs = [embedding_lookup(t, ids[:, f]) for f, t in enumerate(params["tables"])]
    x0 = np.concatenate(embs, axis=1)
    xs = [x0]
    xl = x0
    for p in params["cross"]:
        xl = feature_cross(x0, xl, p["V"], p.get("b"), p.get("U"), p.get("diag_scale", 0.0),
                           p.get("pre_activation"))
        xs.append(xl)
    hs = [xl]
    h = xl
    for (W, b, act) in params["mlp"]:
        h = dense(h, W, b, act)
        hs.append(h)
    if cache is not None:
        cache["xs"], cache["hs"] = xs, hs
    return h


def dcn_backward(params: dict, ids: np.ndarray, dpred: np.ndarray, cache: dict) -> dict:
    """Reverse of dcn_forward; returns grads with the same structure as params
    (tables -> dense (V,E) grads as the non-TPU reference produces)."""
    xs, hs = cache["xs"], cache["hs"]
    g = dpred.reshape(hs[-1].shape)
    mlp_g = []
    for li in range(len(params["mlp"]) - 1, -1, -1):
        W, b, act = params["mlp"][li]
        r = dense_bwd(g, hs[li], W, b, act, hs[li + 1])
        mlp_g.append((r["dW"], r["db"]))
        g = r["dx"]
    mlp_g.reverse()
    x0 = xs[0]
    gx0 = np.zeros_like(x0)
    cross_g = []
    for li in range(len(params["cross"]) - 1, -1, -1):
        p = params["cross"][li]
        r = feature_cross_bwd(g, x0, xs[li], p["V"], p.get("b"), p.get("U"),
                              p.get("diag_scale", 0.0), p.get("pre_activation"))
        cross_g.append(dict(V=r["dV"], b=r["db"], U=r["dU"]))
        gx0 = gx0 + r["dx0"]
        g = r["dx"]
    cross_g.reverse()
    gx0 = gx0 + g
    tg = []
    E = params["tables"][0].shape[1]
    for f, t in enumerate(params["tables"]):
        tg.append(embedding_grad(ids[:, f], None, t.shape[0], gx0[:, f * E:(f + 1) * E], "sum",
                                 reduce=False))
    return dict(tables=tg, cross=cross_g, mlp=mlp_g)


# --------------------------------------------------------------------------- DLRM (examples/ml_perf/model.py:175-212)
def dlrm_forward(params: dict, dense_in: np.ndarray, ids: np.ndarray, interaction: str = "dot", cache: dict | None = None):
    """params: tables [(V,E)], bottom [(W,b)], top [(W,b)], cross [dict(V,b,U)] ; bottom layers relu, top relu ... sigmoid.
    interaction "dot": DotInteraction over [bottom, e_1..e_F] behind the bottom output (classic DLRM, BASELINE C3);
    "cross": DCNBlock over concat([bottom, e_1..e_F]) (model.py:204-208,332-336)."""
    c = {} if cache is None else cache
    h = dense_in
    c["bottom_in"], c["bottom_out"] = [], []
    for W, b in params["bottom"]:
        c["bottom_in"].append(h)
        h = dense(h, W, b, "relu")
        c["bottom_out"].append(h)
    embs = [embedding_lookup(t, ids[:, f]) for f, t in enumerate(params["tables"])]
    c["embs"], c["bottom"] = embs, h
    if interaction == "dot":
        feats = [h] + embs
        z = dot_interaction(feats)
        x = np.concatenate([h, z], axis=-1)
    else:
        x0 = np.concatenate([h] + embs, axis=-1)
        x = x0
        c["x0"], c["cross_in"] = x0, []
        for layer in params["cross"]:
            c["cross_in"].append(x)
            x = feature_cross(x0, x, layer["V"], layer.get("b"), layer.get("U"))
    c["top_in"], c["top_out"] = [], []
    n = len(params["top"])
    for i, (W, b) in enumerate(params["top"]):
        c["top_in"].append(x)
        x = dense(x, W, b, "sigmoid" if i == n - 1 else "relu")
        c["top_out"].append(x)
    return x


def dlrm_backward(params: dict, ids: np.ndarray, dpred: np.ndarray, cache: dict, interaction: str = "dot") -> dict:
    """Analytic gradients of dlrm_forward w.r.t. every parameter (dense (V,E) table gradients)."""
    g = dpred
    n = len(params["top"])
    gtop = [None] * n
    for i in range(n - 1, -1, -1):
        W, b = params["top"][i]
        r = dense_bwd(g, cache["top_in"][i], W, b, "sigmoid" if i == n - 1 else "relu", cache["top_out"][i])
        gtop[i] = (r["dW"], r["db"])
        g = r["dx"]
    E = cache["bottom"].shape[1]
    F = len(params["tables"])
    out = dict(top=gtop, cross=[])
    if interaction == "dot":
        gh = g[:, :E].copy()
        dfeats = dot_interaction_bwd(g[:, E:], [cache["bottom"]] + cache["embs"])
        gh = gh + dfeats[0]
        gembs = dfeats[1:]
    else:
        gx0 = np.zeros_like(cache["x0"])
        gcross = [None] * len(params["cross"])
        for i in range(len(params["cross"]) - 1, -1, -1):
            layer = params["cross"][i]
            r = feature_cross_bwd(g, cache["x0"], cache["cross_in"][i], layer["V"], layer.get("b"), layer.get("U"))
            gcross[i] = {k: r[k] for k in ("dV", "db", "dU") if k in r and r[k] is not None}
            gx0 = gx0 + r["dx0"]
            g = r["dx"]
        gx0 = gx0 + g                      # the first layer's x is x0
        out["cross"] = gcross
        gh = gx0[:, :E]
        gembs = [gx0[:, E * (f + 1):E * (f + 2)] for f in range(F)]
    out["tables"] = [embedding_grad(ids[:, f], None, params["tables"][f].shape[0], gembs[f], reduce=False) for f in range(F)]
    nb = len(params["bottom"])
    gbot = [None] * nb
    g = gh
    for i in range(nb - 1, -1, -1):
        W, b = params["bottom"][i]
        r = dense_bwd(g, cache["bottom_in"][i], W, b, "relu", cache["bottom_out"][i])
        gbot[i] = (r["dW"], r["db"])
        g = r["dx"]
    out["bottom"] = gbot
    return out


def glorot_uniform(rng: np.random.Generator, fan_in: int, fan_out: int) -> np.ndarray:
    limit = math.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-limit, limit, size=(fan_in, fan_out)).astype(F32)
