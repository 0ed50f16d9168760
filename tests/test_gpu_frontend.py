"""DistributedEmbedding front end: ragged / CSR inputs densified like base_distributed_embedding.py:31-92
and checked against the reference's own ragged NumPy oracle (embedding/test_utils.py:245-267), and per-table optimizers
(TableConfig.optimizer, base_distributed_embedding.py:172-186; supported set jax/config_conversion.py:211-288) applied
row-sparsely from the gradient arenas, checked against the oracle's update rules (jax/test_utils.py:474-497 for SGD /
Adagrad; keras Adam / Ftrl restated in np_oracle)."""
import numpy as np
import pytest
import torch

from oracle import np_oracle as O
from util import assert_close, dev, npy

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def K():
    import keras_rs_b200 as K
    return K


def _ragged(rng, B, V, max_len):
    lens = rng.integers(0, max_len + 1, size=B)
    lens[0], lens[1] = max_len, 0                       # a full row and an empty row
    ids = [rng.integers(0, V, size=n) for n in lens]
    w = [rng.uniform(0.5, 2.0, size=n).astype(np.float32) for n in lens]
    return ids, w


@pytest.mark.parametrize("combiner", ["sum", "mean", "sqrtn"])
@pytest.mark.parametrize("form", ["lists", "object_array", "csr"])
@pytest.mark.parametrize("use_w", [False, True])
def test_ragged_inputs_match_the_reference_numpy_oracle(K, combiner, form, use_w):
    rng = np.random.default_rng(11)
    B, V, E, L = 37, 50, 8, 6
    t = K.layers.TableConfig("t", V, E, combiner=combiner)
    layer = K.layers.DistributedEmbedding({"f": K.layers.FeatureConfig("f", t, (B, L), (B, E))})
    layer.build()
    table = npy(layer.get_embedding_tables()["t"])
    ids, w = _ragged(rng, B, V, L)
    if form == "lists":
        x, xw = [list(r) for r in ids], [list(r) for r in w]
    elif form == "object_array":
        x = np.empty((B,), dtype=object)
        xw = np.empty((B,), dtype=object)
        for i in range(B):
            x[i], xw[i] = ids[i], w[i]
    else:
        splits = np.concatenate([[0], np.cumsum([len(r) for r in ids])])
        x, xw = (np.concatenate(ids), splits), (np.concatenate(w), splits)
    out = layer({"f": x}, {"f": xw} if use_w else None)["f"]
    ones = [np.ones(len(r), np.float32) for r in ids]
    exp = np.zeros((B, E), np.float32)
    keep = [i for i in range(B) if len(ids[i]) > 0]       # the oracle divides by sum(w): empty rows are 0 by divide_no_nan
    sub = O.expected_lookup_np([ids[i] for i in keep], [(w if use_w else ones)[i] for i in keep], table, combiner)
    exp[keep] = sub
    assert_close(npy(out), exp, rel=1e-5, what=f"ragged {combiner} {form}")
    assert float(npy(out)[1].__abs__().max()) == 0.0      # the empty row


def test_ragged_row_longer_than_valence_is_an_error(K):
    t = K.layers.TableConfig("t", 10, 4, combiner="sum")
    layer = K.layers.DistributedEmbedding({"f": K.layers.FeatureConfig("f", t, (2, 2), (2, 4))})
    with pytest.raises(ValueError, match="exceeds the dense row length"):
        layer({"f": [[1, 2, 3], [4]]})


def test_per_table_optimizers_are_applied_row_sparsely(K):
    """Four tables, four optimizers (the reference's supported set); one fused backward fills the arenas, then
    apply_table_gradients() runs each table's own rule on the rows that received gradient."""
    rng = np.random.default_rng(5)
    B, E = 64, 8
    specs = [("sgd_t", 40, K.optimizers.SGD(0.5)), ("ada_t", 30, K.optimizers.Adagrad(0.5)), ("adam_t", 50, "adam"),
             ("ftrl_t", 20, K.optimizers.Ftrl(0.5, l1_regularization_strength=0.001, l2_regularization_strength=0.01))]
    tables = {n: K.layers.TableConfig(n, v, E, optimizer=o, combiner="sum") for n, v, o in specs}
    feats = {n: K.layers.FeatureConfig(n, tables[n], (B,), (B, E)) for n, _, _ in specs}
    layer = K.layers.DistributedEmbedding(feats, sparse_grad_arena=True)
    layer.build()
    P = {n: npy(p).copy() for n, p in layer.get_embedding_tables().items()}
    state = {n: dict(m=np.zeros_like(P[n]), v=np.zeros_like(P[n]), acc=np.full_like(P[n], 0.1), lin=np.zeros_like(P[n])) for n in P}
    for step in range(1, 4):
        ids = {n: rng.integers(0, v, size=B).astype(np.int32) for n, v, _ in specs}
        gout = rng.normal(size=(B, 4 * E)).astype(np.float32)
        out = layer({n: dev(i) for n, i in ids.items()}, concat=True)
        out.backward(dev(gout))
        layer.apply_table_gradients()
        order = sorted(ids)                                  # _flatten orders dict features by key
        for j, n in enumerate(order):
            g = O.embedding_grad(ids[n], None, P[n].shape[0], gout[:, j * E:(j + 1) * E])
            rows = np.zeros((P[n].shape[0],), bool)
            rows[ids[n]] = True
            st = state[n]
            if n == "sgd_t":
                P[n] = O.sgd_step(P[n], g, lr=0.5)
            elif n == "ada_t":
                P[n], st["acc"] = O.adagrad_step(P[n], st["acc"], g, lr=0.5)
            elif n == "adam_t":
                P[n], st["m"], st["v"] = O.lazy_adam_step(P[n], st["m"], st["v"], g, step, lr=0.001, rows=rows)
            else:
                P[n], st["acc"], st["lin"] = O.ftrl_step(P[n], st["acc"], st["lin"], g, lr=0.5, l1=0.001, l2=0.01, rows=rows)
        for n, p in layer.get_embedding_tables().items():
            assert_close(npy(p), P[n], rel=2e-5, what=f"{n} after step {step}")
            assert float(p._krs_arena.abs().max()) == 0.0 and int(p._krs_touched.abs().max()) == 0


def test_unsupported_table_optimizer_is_rejected(K):
    t = K.layers.TableConfig("t", 10, 4, optimizer="rmsprop")
    layer = K.layers.DistributedEmbedding({"f": K.layers.FeatureConfig("f", t, (2,), (2, 4))}, sparse_grad_arena=True)
    with pytest.raises(ValueError, match="Optimizer must be one of"):
        layer.table_optimizers()
