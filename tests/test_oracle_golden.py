"""Oracle vs the reference's own known-answer vectors (tests/golden/*.json)."""
import json
import os
import re

import numpy as np
import pytest

from oracle import np_oracle as O

G = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    with open(os.path.join(G, name + ".json")) as f:
        return json.load(f)


def test_feature_cross_golden():
    g = load("feature_cross")
    for c in g["cases"]:
        x0 = np.array(c["x0"], np.float32)
        x = None if c["x"] is None else np.array(c["x"], np.float32)
        D = x0.shape[-1]
        P = c["projection_dim"]
        U = None if P is None else np.ones((D, P), np.float32)
        V = np.ones((D if P is None else P, D), np.float32)
        b = np.zeros((D,), np.float32)
        act = (lambda z: np.zeros_like(z)) if c.get("pre_activation") == "zeros_like" else None
        y = O.feature_cross(x0, x, V, b, U, c["diag_scale"], act)
        np.testing.assert_allclose(y, np.array(c["expected"], np.float32), atol=g["atol"], rtol=g["rtol"])
        shapes = ([list(U.shape)] if U is not None else []) + [list(V.shape), list(b.shape)]
        assert shapes == c["weight_shapes"]


def test_feature_cross_errors():
    with pytest.raises(ValueError):
        O.feature_cross(np.ones((12, 5), np.float32), np.ones((12, 7), np.float32), np.ones((5, 5), np.float32))


def test_dot_interaction_golden():
    g = load("dot_interaction")
    inputs = [np.array(a, np.float32) for a in g["inputs"]]
    for c in g["cases"]:
        out = O.dot_interaction(inputs, c["self_interaction"], c["skip_gather"])
        np.testing.assert_allclose(out, np.array(c["expected"], np.float32), atol=1e-5, rtol=1e-6)


def test_dot_interaction_errors():
    with pytest.raises(ValueError):
        O.dot_interaction([np.ones((3,), np.float32), np.ones((3,), np.float32)])
    with pytest.raises(ValueError):
        O.dot_interaction([np.ones((1, 3), np.float32), np.ones((1, 4), np.float32)])


def test_tril_indices_golden():
    for c in load("tril")["cases"]:
        assert O.tril_indices(c["n"], c["self_interaction"]) == c["idx"]


def test_embed_reduce_golden():
    g = load("embed_reduce")
    rng = np.random.default_rng(0)
    table = rng.uniform(-0.05, 0.05, size=(g["vocab"], g["dim"])).astype(np.float32)
    for c in g["cases"]:
        ids = np.array(c["inputs"], np.int32)
        w = np.array(c["weights"], np.float32) if c["use_weights"] else None
        out = O.embed_reduce(table, ids, w, c["combiner"])
        assert out.shape == (2, g["dim"])
        exp = np.zeros((2, g["dim"]), np.float64)
        for r, terms in enumerate(c["expected_terms"]):
            for row, coeff in terms:
                exp[r] += coeff * table[row].astype(np.float64)
            exp[r] /= c["divisors"][r]
        np.testing.assert_allclose(out, exp, atol=g["atol"], rtol=g["rtol"])


def test_embed_reduce_matches_reference_numpy_oracle():
    # embedding/test_utils.py:245-267 restated as O.expected_lookup_np
    rng = np.random.default_rng(1)
    table = rng.normal(size=(50, 8)).astype(np.float32)
    ids = rng.integers(0, 50, size=(7, 5))
    w = rng.uniform(0.5, 2.0, size=(7, 5)).astype(np.float32)
    for comb in ("sum", "mean", "sqrtn"):
        a = O.embed_reduce(table, ids, w, comb)
        b = O.expected_lookup_np(list(ids), list(w), table, comb)
        np.testing.assert_allclose(a, b, atol=1e-5, rtol=1e-5)


def test_embed_reduce_errors():
    t = np.zeros((10, 4), np.float32)
    with pytest.raises(ValueError):
        O.embed_reduce(t, np.array([1, 2]), None, "max")
    with pytest.raises(ValueError):
        O.embed_reduce(t, np.array([1, 2]), np.ones((3,), np.float32), "sum")


def test_distributed_embedding_golden():
    g = load("distributed_embedding")
    rng = np.random.default_rng(2)
    tab = rng.normal(size=(10, 4)).astype(np.float32)
    ids = np.array(g["ids"] * 4, np.int32)
    out = O.multi_table_gather([tab], [0, 0], [ids, ids], None, ["mean"])
    assert out.shape == (8, 8)
    np.testing.assert_array_equal(out[0, :4], tab[2])
    np.testing.assert_array_equal(out[1, 4:], tab[3])
    w = np.array([1.0, 2.0] * 4, np.float32)
    out = O.multi_table_gather([tab], [0], [ids], [w], ["sum"])
    np.testing.assert_allclose(out[1], 2.0 * tab[3], rtol=1e-6)
    out = O.multi_table_gather([tab], [0], [ids], [w], ["mean"])   # 1-D: weights ignored unless sum
    np.testing.assert_array_equal(out[1], tab[3])


def test_retrieval_validation_golden():
    g = load("retrieval")
    for e in g["errors"]:
        emb = None if e["emb_shape"] is None else np.zeros(e["emb_shape"], np.float32)
        ids = None if e["ids_shape"] is None else np.zeros(e["ids_shape"], np.int32)
        with pytest.raises(ValueError, match=e["regex"]):
            O.validate_candidates(emb, ids, e["k"])


def test_brute_force_golden():
    b = load("retrieval")["brute_force"]
    rng = np.random.default_rng(42)
    cand = rng.normal(size=(b["num_candidates"], b["dim"])).astype(np.float32)
    q = rng.normal(size=(b["num_queries"], b["dim"])).astype(np.float32)
    scores = q @ cand.T
    exp_idx = np.argsort(-scores, axis=1)[:, : b["k"]]
    exp_scores = np.take_along_axis(scores, exp_idx, 1)
    for has_ids in (True, False):
        ids = np.arange(3, b["num_candidates"] + 3, dtype=np.int32) if has_ids else None
        s, i = O.brute_force_retrieval(q, cand, ids, b["k"], True)
        np.testing.assert_allclose(s, exp_scores, atol=b["score_atol"])
        np.testing.assert_array_equal(i, exp_idx + (b["id_offset"] if has_ids else 0))
        assert i.dtype == np.int32
        only = O.brute_force_retrieval(q, cand, ids, b["k"], False)
        np.testing.assert_array_equal(only, i)


def test_topk_tie_break_lowest_index():
    s = np.array([[1.0, 3.0, 3.0, 2.0, 3.0]], np.float32)
    v, i = O.top_k(s, 3)
    assert i.tolist() == [[1, 2, 4]]


def test_embedding_grad_matches_reference_scatter_add():
    # jax/test_utils.py:395-417: grad[cols] += vals * g[rows]
    rng = np.random.default_rng(3)
    ids = rng.integers(0, 9, size=(6, 3))
    w = rng.uniform(0.5, 2, size=(6, 3)).astype(np.float32)
    g = rng.normal(size=(6, 4)).astype(np.float32)
    got = O.embedding_grad(ids, w, 9, g, "sum")
    exp = np.zeros((9, 4), np.float32)
    rows = np.repeat(np.arange(6), 3)
    np.add.at(exp, ids.reshape(-1), w.reshape(-1, 1) * g[rows])
    np.testing.assert_allclose(got, exp, rtol=1e-6, atol=1e-6)


def test_optimizer_rules():
    # jax/test_utils.py:474-497 (SGD, Adagrad)
    rng = np.random.default_rng(4)
    p = rng.normal(size=(5, 3)).astype(np.float32)
    g = rng.normal(size=(5, 3)).astype(np.float32)
    np.testing.assert_allclose(O.sgd_step(p, g, 0.1), p - 0.1 * g, rtol=1e-6)
    acc = np.full_like(p, 0.1)
    p2, acc2 = O.adagrad_step(p, acc, g, lr=0.05, eps=0.0)
    np.testing.assert_allclose(acc2, acc + g * g, rtol=1e-6)
    np.testing.assert_allclose(p2, p - 0.05 / np.sqrt(acc + g * g) * g, rtol=1e-5)


def test_mod_route_roundtrip():
    rng = np.random.default_rng(5)
    t = rng.normal(size=(37, 4)).astype(np.float32)
    sh = O.mod_shard_table(t, 8)
    np.testing.assert_array_equal(O.mod_unshard_table(sh), t)
    ids = rng.integers(0, 37, size=100)
    owner, local = O.mod_route(ids, 8)
    for i, o, l in zip(ids, owner, local):
        np.testing.assert_array_equal(sh[o][l], t[i])


def test_embedding_lookup_out_of_range_ids_jnp_take_fill():
    """jnp.take default mode "fill" (keras.ops.take on the JAX backend): negative ids wrap once, the rest is NaN forward
    and dropped backward."""
    tab = np.arange(12, dtype=np.float32).reshape(4, 3)
    out = O.embedding_lookup(tab, np.array([0, -1, -4, -5, 4, 3]))
    np.testing.assert_array_equal(out[[0, 1, 2, 5]], tab[[0, 3, 0, 3]])
    assert np.isnan(out[[3, 4]]).all()
    g = np.ones((6, 3), np.float32)
    grad = O.embedding_grad(np.array([0, -1, -4, -5, 4, 3]), None, 4, g)
    np.testing.assert_array_equal(grad[:, 0], [2, 0, 0, 2])
