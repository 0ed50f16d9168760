"""GPU parity tests: every CUDA kernel, called through the C ABI / public layers, against the CPU
oracle (oracle/np_oracle.py) and the reference's golden vectors (tests/golden)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import np_oracle as O
from util import assert_close, dev, npy

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def golden(name):
    with open(os.path.join(G, name + ".json")) as f:
        return json.load(f)


@pytest.fixture(scope="module")
def K():
    import keras_rs_b200 as k
    return k


# ------------------------------------------------------------------------------------ sgemm
@pytest.mark.parametrize("tA", [False, True])
@pytest.mark.parametrize("tB", [False, True])
@pytest.mark.parametrize("M,N,K_", [(128, 128, 64), (257, 131, 77), (5, 3, 2), (1000, 1, 192), (192, 1, 1000), (64, 832, 4096)])
def test_sgemm(K, tA, tB, M, N, K_):
    rng = np.random.default_rng(M * 7 + N)
    A = rng.normal(size=(K_, M) if tA else (M, K_)).astype(np.float32)
    B = rng.normal(size=(N, K_) if tB else (K_, N)).astype(np.float32)
    ref = (A.T if tA else A).astype(np.float64) @ (B.T if tB else B).astype(np.float64)
    got = K.ops.sgemm(dev(A), dev(B), tA, tB)
    assert_close(npy(got), ref, what=f"sgemm {tA}{tB}")


def test_sgemm_accumulate(K):
    rng = np.random.default_rng(0)
    A = rng.normal(size=(70, 33)).astype(np.float32)
    B = rng.normal(size=(33, 20)).astype(np.float32)
    C0 = rng.normal(size=(70, 20)).astype(np.float32)
    out = dev(C0.copy())
    K.ops.sgemm(dev(A), dev(B), out=out, accumulate=True)
    assert_close(npy(out), C0 + A @ B, what="accumulate")


# ------------------------------------------------------------------------------------ gather fwd
def _tables(rng, F, V, E):
    return [rng.uniform(-0.05, 0.05, size=(V, E)).astype(np.float32) for _ in range(F)]


@pytest.mark.parametrize("E", [4, 8, 32, 64, 128, 6, 20])
@pytest.mark.parametrize("idt", [torch.int32, torch.int64])
@pytest.mark.parametrize("variant", [0, 2, 3])
def test_gather_onehot_bit_exact(K, E, idt, variant):
    if variant == 2 and (E % 4 != 0 or (E // 4) & (E // 4 - 1)):
        pytest.skip("bulk variant needs E/4 power of two")
    rng = np.random.default_rng(E)
    F, V, B = 5, 1000, 777
    tabs = _tables(rng, F, V, E)
    ids = rng.integers(0, V, size=(B, F))
    ids[0, 0], ids[1, 1] = 0, V - 1
    dt = [dev(t) for t in tabs]
    dids = dev(ids, idt)
    feats = [dict(table=dt[f], ids=dids[:, f], combiner="sum") for f in range(F)]
    plan = K.ops.GatherPlan(feats)
    out = plan.forward(variant=variant)
    ref = O.multi_table_gather(tabs, list(range(F)), [ids[:, f] for f in range(F)], None, ["sum"] * F)
    np.testing.assert_array_equal(npy(out), ref)      # pure row copy: bit exact


def test_gather_separate_id_tensors_and_shared_table(K):
    rng = np.random.default_rng(1)
    V, E, B = 50, 32, 64
    tab = _tables(rng, 1, V, E)[0]
    a, b = rng.integers(0, V, size=B), rng.integers(0, V, size=B)
    dt = dev(tab)
    out = K.ops.gather_concat([dict(table=dt, ids=dev(a, torch.int32), combiner="mean"),
                               dict(table=dt, ids=dev(b, torch.int32), combiner="mean")])
    np.testing.assert_array_equal(npy(out), np.concatenate([tab[a], tab[b]], axis=1))


def test_gather_empty_batch(K):
    dt = dev(np.zeros((10, 8), np.float32))
    out = K.ops.gather_concat([dict(table=dt, ids=torch.zeros((0,), dtype=torch.int32, device="cuda"), combiner="sum")])
    assert tuple(out.shape) == (0, 8)


@pytest.mark.parametrize("variant,E", [(0, 4), (0, 32), (2, 32), (3, 4), (3, 6)])
@pytest.mark.parametrize("idt", [torch.int32, torch.int64])
def test_gather_out_of_range_ids_follow_jnp_take_fill(K, variant, E, idt):
    """ids < 0 count from the end; ids still outside [0, vocab) give a NaN row forward and no gradient backward
    (jnp.take mode="fill" = keras.ops.take on the JAX backend; oracle embedding_lookup / embedding_grad)."""
    rng = np.random.default_rng(5)
    V, B = 10, 70
    tab = rng.normal(size=(V, E)).astype(np.float32)
    ids = rng.integers(0, V, size=B)
    ids[[0, 1, 2, 3, 40, 69]] = [-3, 99, -V, -V - 1, V, -1]
    dt = dev(tab)
    plan = K.ops.GatherPlan([dict(table=dt, ids=dev(ids, idt), combiner="sum")])
    out = npy(plan.forward(variant=variant))
    ref = O.embedding_lookup(tab, ids)
    assert np.isnan(ref[[1, 3, 40]]).all() and not np.isnan(ref[[0, 2, 69]]).any()
    np.testing.assert_array_equal(out, ref)                    # NaN == NaN under assert_array_equal
    np.testing.assert_array_equal(out[0], tab[V - 3])
    gout = rng.normal(size=(B, E)).astype(np.float32)
    grad = torch.zeros_like(dt)
    touched = torch.zeros(((V + 31) // 32,), dtype=torch.int32, device="cuda")
    plan.backward(dev(gout), [grad], [touched])
    assert_close(npy(grad), O.embedding_grad(ids, None, V, gout), rel=1e-6, what="grad with dropped ids")
    exp_bits = 0
    for r in set(int(i) % V if -V <= int(i) < V else -1 for i in ids) - {-1}:
        exp_bits |= 1 << r
    assert int(touched[0].item()) & 0xffffffff == exp_bits


def test_gather_multihot_invalid_id_poisons_only_its_sample(K):
    rng = np.random.default_rng(6)
    V, E, B, H = 20, 8, 9, 3
    tab = rng.normal(size=(V, E)).astype(np.float32)
    ids = rng.integers(0, V, size=(B, H))
    ids[4, 1] = V + 7
    w = rng.uniform(0.5, 1.5, size=(B, H)).astype(np.float32)
    for comb in ("sum", "mean", "sqrtn"):
        out = npy(K.ops.gather_concat([dict(table=dev(tab), ids=dev(ids, torch.int32), weights=dev(w), combiner=comb)]))
        ref = O.embed_reduce(tab, ids, w, comb)
        assert np.isnan(out[4]).all() and np.isnan(ref[4]).all()
        keep = np.arange(B) != 4
        assert_close(out[keep], ref[keep], rel=1e-6, what=comb)


@pytest.mark.parametrize("combiner", ["sum", "mean", "sqrtn"])
@pytest.mark.parametrize("use_w", [False, True])
@pytest.mark.parametrize("E", [32, 6])
def test_gather_multihot_combiners(K, combiner, use_w, E):
    rng = np.random.default_rng(3)
    V, B = 200, 129
    hots = [1, 3, 7]
    tabs = _tables(rng, 3, V, E)
    ids = [rng.integers(0, V, size=(B, h)) for h in hots]
    ws = [rng.uniform(0.0, 2.0, size=(B, h)).astype(np.float32) for h in hots]
    ws[1][0, :] = 0.0      # all-zero weights row -> divide_no_nan gives 0
    feats = [dict(table=dev(tabs[f]), ids=dev(ids[f], torch.int64), weights=dev(ws[f]) if use_w else None,
                  combiner=combiner) for f in range(3)]
    out = K.ops.gather_concat(feats)
    ref = O.multi_table_gather(tabs, [0, 1, 2], ids, ws if use_w else None, [combiner] * 3)
    assert_close(npy(out), ref, rel=1e-6, what="combiner")


def test_embed_reduce_golden(K):
    g = golden("embed_reduce")
    for c in g["cases"]:
        layer = K.layers.EmbedReduce(g["vocab"], g["dim"], combiner=c["combiner"])
        table = npy(layer.embeddings)
        ids = dev(np.array(c["inputs"]), torch.int32)
        w = dev(np.array(c["weights"], np.float32)) if c["use_weights"] else None
        out = npy(layer(ids, w))
        assert out.shape == (2, g["dim"])
        exp = np.zeros((2, g["dim"]))
        for r, terms in enumerate(c["expected_terms"]):
            for row, coeff in terms:
                exp[r] += coeff * table[row].astype(np.float64)
            exp[r] /= c["divisors"][r]
        np.testing.assert_allclose(out, exp, atol=g["atol"], rtol=g["rtol"])


def test_embed_reduce_errors(K):
    with pytest.raises(ValueError, match="Invalid `combiner`"):
        K.layers.EmbedReduce(10, 4, combiner="max")
    layer = K.layers.EmbedReduce(10, 4, combiner="sum")
    with pytest.raises(ValueError, match="not compatible"):
        layer(dev(np.array([1, 2]), torch.int32), dev(np.ones((3,), np.float32)))


def test_embedding_layer_rank2_ids(K):
    layer = K.layers.Embedding(20, 8)
    ids = np.array([[1, 2, 3], [4, 5, 6]])
    out = npy(layer(dev(ids, torch.int32)))
    np.testing.assert_array_equal(out, npy(layer.embeddings)[ids])


def test_distributed_embedding_golden(K):
    from keras_rs_b200.layers import DistributedEmbedding, FeatureConfig, TableConfig
    g = golden("distributed_embedding")
    B = 8
    t1 = TableConfig("t1", 10, 4, combiner="sum")
    t2 = TableConfig("t2", 12, 8, combiner="mean")
    fcs = {"a": FeatureConfig("a", t1, (B,), (B, 4)), "n": {"b": FeatureConfig("b", t2, (B,), (B, 8)),
                                                              "c": FeatureConfig("c", t1, (B,), (B, 4))}}
    layer = DistributedEmbedding(fcs)
    ids = dev(np.array(g["ids"] * (B // 2)), torch.int32)
    w = dev(np.array([1.0, 2.0] * (B // 2), np.float32))
    out = layer({"a": ids, "n": {"b": ids, "c": ids}}, {"a": w, "n": {"b": w, "c": w}})
    tabs = {k: npy(v) for k, v in layer.get_embedding_tables().items()}
    assert set(out.keys()) == {"a", "n"} and set(out["n"].keys()) == {"b", "c"}
    assert len(layer.weights) == 2                                   # shared table -> one variable
    np.testing.assert_allclose(npy(out["a"])[0], tabs["t1"][2] * 1.0, rtol=1e-6)
    np.testing.assert_allclose(npy(out["a"])[1], tabs["t1"][3] * 2.0, rtol=1e-6)   # sum + weights
    np.testing.assert_array_equal(npy(out["n"]["b"])[1], tabs["t2"][3])            # mean, 1-D: weights ignored
    pre = layer.preprocess({"a": ids, "n": {"b": ids, "c": ids}})
    out2 = layer(pre)
    np.testing.assert_array_equal(npy(out2["n"]["c"])[0], tabs["t1"][2])
    cat = layer({"a": ids, "n": {"b": ids, "c": ids}}, concat=True)
    assert tuple(cat.shape) == (B, 16)
    with pytest.raises(ValueError, match="incompatible"):
        layer({"a": ids.reshape(4, 2), "n": {"b": ids, "c": ids}})


# ------------------------------------------------------------------------------------ gather bwd
@pytest.mark.parametrize("E", [32, 128, 6])
@pytest.mark.parametrize("idt", [torch.int32, torch.int64])
def test_scatter_onehot_with_duplicates(K, E, idt):
    rng = np.random.default_rng(5)
    F, V, B = 3, 40, 300          # tiny vocab: many duplicates inside every warp
    tabs = _tables(rng, F, V, E)
    ids = rng.integers(0, V, size=(B, F))
    ids[:64, 0] = 7                # two full warps of identical ids
    gout = rng.normal(size=(B, F * E)).astype(np.float32)
    dt = [dev(t) for t in tabs]
    dids = dev(ids, idt)
    plan = K.ops.GatherPlan([dict(table=dt[f], ids=dids[:, f], combiner="sum") for f in range(F)])
    grads = [torch.zeros_like(t) for t in dt]
    touched = [torch.zeros(((V + 31) // 32,), dtype=torch.int32, device="cuda") for _ in range(F)]
    plan.backward(dev(gout), grads, touched)
    for f in range(F):
        ref = O.embedding_grad(ids[:, f], None, V, gout[:, f * E:(f + 1) * E], "sum", reduce=False)
        assert_close(npy(grads[f]), ref, what=f"scatter f={f}")
        bits = npy(touched[f]).view(np.uint32)
        exp = np.zeros(((V + 31) // 32,), np.uint32)
        for r in np.unique(ids[:, f]):
            exp[r >> 5] |= np.uint32(1 << (r & 31))
        np.testing.assert_array_equal(bits, exp)                    # integer path: exact


@pytest.mark.parametrize("combiner", ["sum", "mean", "sqrtn"])
def test_scatter_multihot_weighted(K, combiner):
    rng = np.random.default_rng(6)
    V, B, E, H = 60, 100, 8, 4
    tab = _tables(rng, 1, V, E)[0]
    ids = rng.integers(0, V, size=(B, H))
    w = rng.uniform(0.5, 2.0, size=(B, H)).astype(np.float32)
    gout = rng.normal(size=(B, E)).astype(np.float32)
    dt = dev(tab).requires_grad_(True)
    out = K.ops.gather_concat([dict(table=dt, ids=dev(ids, torch.int32), weights=dev(w), combiner=combiner)])
    out.backward(dev(gout))
    ref = O.embedding_grad(ids, w, V, gout, combiner)
    assert_close(npy(dt.grad), ref, what="multihot grad")


def test_shared_table_grad_accumulates_across_features(K):
    # jax/test_utils.py:450-471: per-table gradient = sum over the features that use the table
    rng = np.random.default_rng(7)
    V, E, B = 30, 32, 50
    tab = _tables(rng, 1, V, E)[0]
    a, b = rng.integers(0, V, size=B), rng.integers(0, V, size=B)
    g = rng.normal(size=(B, 2 * E)).astype(np.float32)
    dt = dev(tab).requires_grad_(True)
    out = K.ops.gather_concat([dict(table=dt, ids=dev(a, torch.int32), combiner="sum"),
                               dict(table=dt, ids=dev(b, torch.int32), combiner="sum")])
    out.backward(dev(g))
    ref = (O.embedding_grad(a, None, V, g[:, :E], "sum", False) + O.embedding_grad(b, None, V, g[:, E:], "sum", False))
    assert_close(npy(dt.grad), ref, what="shared table grad")


# ------------------------------------------------------------------------------------ FeatureCross
def test_feature_cross_golden(K):
    g = golden("feature_cross")
    for c in g["cases"]:
        kw = dict(projection_dim=c["projection_dim"], diag_scale=c["diag_scale"], kernel_initializer="ones")
        if c.get("pre_activation") == "zeros_like":
            kw["pre_activation"] = torch.zeros_like
        layer = K.layers.FeatureCross(**kw)
        x0 = dev(np.array(c["x0"], np.float32))
        out = layer(x0) if c["x"] is None else layer(x0, dev(np.array(c["x"], np.float32)))
        np.testing.assert_allclose(npy(out), np.array(c["expected"]), atol=g["atol"], rtol=g["rtol"])
        assert [list(w.shape) for w in layer.weights] == c["weight_shapes"]
    layer = K.layers.FeatureCross()
    with pytest.raises(ValueError, match="same shape"):
        layer(torch.ones(12, 5, device="cuda"), torch.ones(12, 7, device="cuda"))


ACTS = [None, "relu", "sigmoid", "tanh", "swish"]


@pytest.mark.parametrize("P", [None, 20])
@pytest.mark.parametrize("act", ACTS)
@pytest.mark.parametrize("diag", [0.0, 0.5])
@pytest.mark.parametrize("same", [False, True])
def test_feature_cross_fwd_bwd_vs_oracle(K, P, act, diag, same):
    rng = np.random.default_rng(11)
    B, D = 300, 96
    x0 = rng.normal(size=(B, D)).astype(np.float32)
    x = x0 if same else rng.normal(size=(B, D)).astype(np.float32)
    layer = K.layers.FeatureCross(projection_dim=P, diag_scale=diag, pre_activation=act,
                                  bias_initializer=K.initializers.RandomUniform(-0.5, 0.5, seed=3))
    tx0 = dev(x0).requires_grad_(True)
    tx = None if same else dev(x).requires_grad_(True)
    y = layer(tx0) if same else layer(tx0, tx)
    U = npy(layer.down_proj_kernel) if P is not None else None
    V, b = npy(layer.kernel), npy(layer.bias)
    ref = O.feature_cross(x0, None if same else x, V, b, U, diag, act)
    assert_close(npy(y), ref, what="cross fwd")
    gy = rng.normal(size=(B, D)).astype(np.float32)
    y.backward(dev(gy))
    # derivative mask from the kernel's own pre-activation (relu' is discontinuous at 0)
    hz = dev(x) if P is None else K.ops.linear_no_bias(dev(x), layer.down_proj_kernel.detach())
    z_gpu = npy(K.ops.dense(hz.contiguous(), layer.kernel.detach(), layer.bias.detach(), 0))
    r = O.feature_cross_bwd(gy, x0, x, V, b, U, diag, act, z_for_grad=z_gpu)
    if same:
        assert_close(npy(tx0.grad), r["dx0"] + r["dx"], what="dx total")
    else:
        assert_close(npy(tx0.grad), r["dx0"], what="dx0")
        assert_close(npy(tx.grad), r["dx"], what="dx")
    assert_close(npy(layer.kernel.grad), r["dV"], what="dV")
    assert_close(npy(layer.bias.grad), r["db"], what="db")
    if P is not None:
        assert_close(npy(layer.down_proj_kernel.grad), r["dU"], what="dU")


def test_feature_cross_rank3_and_no_bias(K):
    rng = np.random.default_rng(12)
    x0 = rng.normal(size=(4, 5, 16)).astype(np.float32)
    x = rng.normal(size=(4, 5, 16)).astype(np.float32)
    layer = K.layers.FeatureCross(use_bias=False)
    y = layer(dev(x0), dev(x))
    assert len(layer.weights) == 1
    ref = O.feature_cross(x0.reshape(-1, 16), x.reshape(-1, 16), npy(layer.kernel), None).reshape(4, 5, 16)
    assert_close(npy(y), ref, what="rank3")


def test_dcn_block_stack(K):
    # README.md:54-55 / ml_perf/model.py:332-336: xl = layer(x0, xl)
    rng = np.random.default_rng(13)
    x0 = rng.normal(size=(64, 48)).astype(np.float32)
    layers = [K.layers.FeatureCross() for _ in range(3)]
    t0 = dev(x0)
    xl = t0
    for l in layers:
        xl = l(t0, xl)
    ref = O.dcn_block(x0, [dict(V=npy(l.kernel), b=npy(l.bias)) for l in layers])
    assert_close(npy(xl), ref, what="dcn block")


# ------------------------------------------------------------------------------------ Dense
@pytest.mark.parametrize("act", [None, "relu", "sigmoid", "tanh"])
@pytest.mark.parametrize("Kd,N", [(96, 192), (192, 1), (13, 512)])
def test_dense_fwd_bwd(K, act, Kd, N):
    rng = np.random.default_rng(21)
    B = 333
    x = rng.normal(size=(B, Kd)).astype(np.float32)
    layer = K.layers.Dense(N, activation=act, bias_initializer=K.initializers.RandomUniform(-0.5, 0.5, seed=1))
    tx = dev(x).requires_grad_(True)
    y = layer(tx)
    W, b = npy(layer.kernel), npy(layer.bias)
    ref = O.dense(x, W, b, act)
    assert_close(npy(y), ref, what="dense fwd")
    gy = rng.normal(size=(B, N)).astype(np.float32)
    y.backward(dev(gy))
    r = O.dense_bwd(gy, x, W, b, act, ref)
    assert_close(npy(tx.grad), r["dx"], what="dense dx")
    assert_close(npy(layer.kernel.grad), r["dW"], what="dense dW")
    # a column sum with cancellation: bound the error by the magnitude of the summed terms
    dz_abs = np.abs(O.dense_bwd(gy, x, W, b, act, ref)["dx"]).max() * 0 + np.abs(gy).sum(axis=0).max()
    assert_close(npy(layer.bias.grad), r["db"], what="dense db", scale=dz_abs)


# ------------------------------------------------------------------------------------ DotInteraction
def test_dot_interaction_golden(K):
    g = golden("dot_interaction")
    inputs = [dev(np.array(a, np.float32)) for a in g["inputs"]]
    for c in g["cases"]:
        layer = K.layers.DotInteraction(self_interaction=c["self_interaction"], skip_gather=c["skip_gather"])
        out = layer(inputs)
        np.testing.assert_allclose(npy(out), np.array(c["expected"]), atol=1e-5, rtol=1e-6)
        assert tuple(out.shape) == layer.compute_output_shape([(1, 5)] * 3)


@pytest.mark.parametrize("self_i", [False, True])
@pytest.mark.parametrize("skip", [False, True])
@pytest.mark.parametrize("N,E", [(27, 128), (8, 32), (3, 16), (5, 7), (32, 64)])
def test_dot_interaction_vs_oracle(K, self_i, skip, N, E):
    rng = np.random.default_rng(31)
    B = 130
    buf = rng.normal(size=(B, N * E)).astype(np.float32)
    tb = dev(buf).requires_grad_(True)
    feats = [tb[:, i * E:(i + 1) * E] for i in range(N)]        # strided views of a concat buffer
    layer = K.layers.DotInteraction(self_interaction=self_i, skip_gather=skip)
    out = layer(feats)
    inputs = [buf[:, i * E:(i + 1) * E] for i in range(N)]
    ref = O.dot_interaction(inputs, self_i, skip)
    assert_close(npy(out), ref, what="dot fwd")
    if skip:     # masked entries are exact zeros (dot_interaction.py:182-192)
        mask = np.tril(np.ones((N, N), bool), 0 if self_i else -1).reshape(-1)
        assert (npy(out)[:, ~mask] == 0).all()
    g = rng.normal(size=ref.shape).astype(np.float32)
    out.backward(dev(g))
    dref = np.concatenate(O.dot_interaction_bwd(g, inputs, self_i, skip), axis=1)
    assert_close(npy(tb.grad), dref, what="dot bwd")


# ------------------------------------------------------------------------------------ retrieval
@pytest.mark.parametrize("has_ids", [True, False])
@pytest.mark.parametrize("return_scores", [True, False])
def test_brute_force_golden(K, has_ids, return_scores):
    b = golden("retrieval")["brute_force"]      # brute_force_retrieval_test.py:13-64
    rng = np.random.default_rng(42)
    cand = rng.normal(size=(b["num_candidates"], b["dim"])).astype(np.float32)
    q = rng.normal(size=(b["num_queries"], b["dim"])).astype(np.float32)
    ids = np.arange(3, b["num_candidates"] + 3, dtype=np.int32) if has_ids else None
    layer = K.layers.BruteForceRetrieval(candidate_embeddings=dev(cand), candidate_ids=None if ids is None else dev(ids),
                                         k=b["k"], return_scores=return_scores)
    exp_s, exp_i = O.brute_force_retrieval(q, cand, ids, b["k"], True)
    for rep in range(2):
        if rep:
            layer.update_candidates(dev(cand), None if ids is None else dev(ids))
        out = layer(dev(q))
        if return_scores:
            s, i = out
            np.testing.assert_allclose(npy(s), exp_s, atol=b["score_atol"])
        else:
            i = out
        assert i.dtype == torch.int32 and tuple(i.shape) == exp_i.shape
        np.testing.assert_array_equal(npy(i), exp_i)


def _check_topk(scores_ref, got_s, got_i, k, tol=2e-5):
    """indices exact where the score gap to the neighbours exceeds tol, set-consistent otherwise."""
    nq = scores_ref.shape[0]
    order = np.argsort(-scores_ref, axis=1, kind="stable")
    for r in range(nq):
        ref_scores = scores_ref[r, order[r, :k + 1]]
        np.testing.assert_allclose(got_s[r], ref_scores[:k], atol=1e-4, rtol=1e-5)
        # the returned indices must carry (near-)the returned scores
        np.testing.assert_allclose(scores_ref[r, got_i[r]], got_s[r], atol=1e-4, rtol=1e-5)
        gaps = np.abs(np.diff(ref_scores))
        for j in range(k):
            left = gaps[j - 1] if j > 0 else np.inf
            right = gaps[j]
            if left > tol and right > tol:
                assert got_i[r, j] == order[r, j], f"row {r} rank {j}"
        assert len(set(got_i[r].tolist())) == k


@pytest.fixture(params=["ffma", "tcgen05"])
def topk_engine(request, K):
    K.ops.set_topk_engine(request.param)
    yield request.param
    K.ops.set_topk_engine("auto")


@pytest.mark.parametrize("nq,nc,d,k", [(33, 5000, 64, 100), (200, 20000, 32, 10), (7, 300, 20, 128), (64, 129, 128, 5),
                                       (300, 40000, 64, 100), (129, 3000, 48, 1), (5, 97, 8, 96)])
def test_topk_vs_oracle(K, topk_engine, nq, nc, d, k):
    rng = np.random.default_rng(nq + nc)
    q = rng.normal(size=(nq, d)).astype(np.float32)
    c = rng.normal(size=(nc, d)).astype(np.float32)
    before = K._lib.lib.krs_topk_tc_launch_count()
    s, i = K.ops.top_k_scores(dev(q), dev(c), None, k)
    ran_tc = K._lib.lib.krs_topk_tc_launch_count() - before
    eligible = d % 4 == 0 and d <= 64 and nc >= 96
    assert ran_tc == (1 if (topk_engine == "tcgen05" and eligible) else 0)
    ref = (q.astype(np.float64) @ c.astype(np.float64).T)
    _check_topk(ref, npy(s), npy(i), k)
    assert (np.diff(npy(s), axis=1) <= 0).all()                     # sorted descending


def test_topk_ties_lowest_index_first(K, topk_engine):
    c = np.zeros((300, 8), np.float32)
    c[:, 0] = 1.0
    c[10, 0] = 2.0
    q = np.ones((3, 8), np.float32)
    s, i = K.ops.top_k_scores(dev(q), dev(c), None, 6)
    assert npy(i).tolist() == [[10, 0, 1, 2, 3, 4]] * 3           # jax.lax.top_k tie rule


def test_topk_tc_candidate_ids_and_large_slice(K):
    """tensor-pipe path with candidate ids, several slices per query tile and a ragged last tile."""
    K.ops.set_topk_engine("tcgen05")
    try:
        rng = np.random.default_rng(9)
        nq, nc, d, k = 260, 100_003, 64, 50
        q = rng.normal(size=(nq, d)).astype(np.float32)
        c = rng.normal(size=(nc, d)).astype(np.float32)
        ids = (np.arange(nc, dtype=np.int64) * 7 + 11).astype(np.int32)
        s, i = K.ops.top_k_scores(dev(q), dev(c), dev(ids), k)
        ref = q.astype(np.float64) @ c.astype(np.float64).T
        raw = (npy(i).astype(np.int64) - 11) // 7
        _check_topk(ref, npy(s), raw, k)
    finally:
        K.ops.set_topk_engine("auto")


def test_retrieval_shared_variable_pattern(K):
    # examples/basic_retrieval.py:249-257: assign another layer's embedding variable, then call
    emb = K.layers.Embedding(50, 16)
    r = K.layers.BruteForceRetrieval(k=5, return_scores=False)
    r.candidate_embeddings = emb.embeddings
    q = npy(emb.embeddings)[[3, 7]]
    out = r(dev(q))
    exp = O.brute_force_retrieval(q, npy(emb.embeddings), None, 5, False)
    np.testing.assert_array_equal(npy(out), exp)


# ------------------------------------------------------------------------------------ loss + optimizers
@pytest.mark.parametrize("kind", ["mse", "bce"])
def test_loss(K, kind):
    rng = np.random.default_rng(41)
    B = 1000
    y = rng.uniform(size=B).astype(np.float32)
    p = rng.uniform(0.01, 0.99, size=(B, 1)).astype(np.float32)
    tp = dev(p).requires_grad_(True)
    loss = K.ops.loss_fn(tp, dev(y), kind)
    loss.backward()
    rl, rd = (O.mse_loss if kind == "mse" else O.bce_loss)(p, y)
    np.testing.assert_allclose(float(loss), float(rl), rtol=1e-5)
    assert_close(npy(tp.grad).reshape(-1), rd, what="dloss")


def test_adamw_dense_and_arena_match_oracle(K):
    rng = np.random.default_rng(51)
    V, E = 100, 8
    p0 = rng.normal(size=(V, E)).astype(np.float32)
    opt_d = K.optimizers.AdamW(learning_rate=0.01)
    opt_a = K.optimizers.AdamW(learning_rate=0.01)
    pd_, pa = dev(p0.copy()), dev(p0.copy())
    arena = torch.zeros_like(pa)
    touched = torch.zeros(((V + 31) // 32,), dtype=torch.int32, device="cuda")
    pa._krs_arena, pa._krs_touched = arena, touched
    p, m, v = p0.copy(), np.zeros_like(p0), np.zeros_like(p0)
    for step in range(1, 4):
        rows = rng.choice(V, size=10, replace=False)
        g = np.zeros_like(p0)
        g[rows] = rng.normal(size=(10, E)).astype(np.float32)
        pd_.grad = dev(g)
        opt_d.apply([pd_])
        arena.copy_(dev(g))
        bits = np.zeros(((V + 31) // 32,), np.uint32)
        for r in rows:
            bits[r >> 5] |= np.uint32(1 << (r & 31))
        touched.copy_(dev(bits.view(np.int32)))
        opt_a.apply([pa])
        p, m, v = O.adamw_step(p, m, v, g, step, lr=0.01)
        assert_close(npy(pd_), p, rel=2e-6, what="adamw dense")
        np.testing.assert_array_equal(npy(pa), npy(pd_))            # arena path == dense path, bitwise
        assert float(arena.abs().max()) == 0.0 and int(touched.abs().max()) == 0   # arena re-zeroed


def test_adamw_cold_rows_are_bitwise_equal_to_the_dense_update(K):
    """krs_adamw_cold: rows that never received a gradient get the decay-only update; parameters AND moments must stay
    bit-identical to the dense rule over several steps, and the ever-touched bitmap must accumulate the touched rows."""
    rng = np.random.default_rng(53)
    V, E = 4096, 32
    p0 = rng.normal(size=(V, E)).astype(np.float32)
    opt_d, opt_c = K.optimizers.AdamW(learning_rate=0.01), K.optimizers.AdamW(learning_rate=0.01)
    pd_, pc = dev(p0.copy()), dev(p0.copy())
    arena = torch.zeros_like(pc)
    touched = torch.zeros((V // 32,), dtype=torch.int32, device="cuda")
    ever = torch.zeros_like(touched)
    pc._krs_arena, pc._krs_touched, pc._krs_ever = arena, touched, ever
    seen = np.zeros((V,), bool)
    for step in range(6):
        rows = rng.choice(V, size=200, replace=False)
        g = np.zeros_like(p0)
        g[rows] = rng.normal(size=(200, E)).astype(np.float32)
        pd_.grad = dev(g)
        opt_d.apply([pd_])
        arena.copy_(dev(g))
        bits = np.zeros((V // 32,), np.uint32)
        for r in rows:
            bits[r >> 5] |= np.uint32(1 << (r & 31))
        touched.copy_(dev(bits.view(np.int32)))
        opt_c.apply([pc])
        seen[rows] = True
        np.testing.assert_array_equal(npy(pc), npy(pd_))
        for slot in ("m", "v"):
            np.testing.assert_array_equal(npy(opt_c._state[id(pc)][slot]), npy(opt_d._state[id(pd_)][slot]))
        got = npy(ever).view(np.uint32)
        exp = np.zeros((V // 32,), np.uint32)
        for r in np.nonzero(seen)[0]:
            exp[r >> 5] |= np.uint32(1 << (r & 31))
        np.testing.assert_array_equal(got, exp)
        assert float(arena.abs().max()) == 0.0 and int(touched.abs().max()) == 0


@pytest.mark.parametrize("V,E,nrows", [(1000, 32, 50), (1_300_003, 8, 5000), (5_000_000, 4, 20000)])   # 1 / 2 / 8 bitmap words per warp
@pytest.mark.parametrize("name", ["sgd", "adagrad"])
def test_sparse_rows_optimizers(K, name, V, E, nrows):
    rng = np.random.default_rng(52)
    p0 = rng.normal(size=(V, E)).astype(np.float32)
    mk = (lambda: K.optimizers.SGD(0.1)) if name == "sgd" else (lambda: K.optimizers.Adagrad(0.05))
    opt_d, opt_a = mk(), mk()
    pd_, pa = dev(p0.copy()), dev(p0.copy())
    arena = torch.zeros_like(pa)
    touched = torch.zeros(((V + 31) // 32,), dtype=torch.int32, device="cuda")
    pa._krs_arena, pa._krs_touched = arena, touched
    p, acc = p0.copy(), np.full_like(p0, 0.1)
    for step in range(3):
        rows = rng.choice(V, size=nrows, replace=False)
        g = np.zeros_like(p0)
        g[rows] = rng.normal(size=(nrows, E)).astype(np.float32)
        pd_.grad = dev(g)
        opt_d.apply([pd_])
        arena.copy_(dev(g))
        bits = np.zeros(((V + 31) // 32,), np.uint32)
        np.bitwise_or.at(bits, rows >> 5, (np.uint32(1) << (rows & 31).astype(np.uint32)))
        touched.copy_(dev(bits.view(np.int32)))
        opt_a.apply([pa])
        if name == "sgd":
            p = O.sgd_step(p, g, 0.1)
        else:
            p, acc = O.adagrad_step(p, acc, g, lr=0.05)
        assert_close(npy(pd_), p, rel=2e-6, what=name)
        np.testing.assert_array_equal(npy(pa), npy(pd_))
        assert float(arena.abs().max()) == 0.0 and int(touched.abs().max()) == 0


def test_mod_route_bit_exact(K):
    from keras_rs_b200._lib import check, lib, ptr, stream
    rng = np.random.default_rng(61)
    ids = rng.integers(0, 10**9, size=5000)
    for dt in (torch.int32, torch.int64):
        t = dev(ids, dt)
        owner = torch.empty(5000, dtype=torch.int32, device="cuda")
        local = torch.empty(5000, dtype=torch.int64, device="cuda")
        counts = torch.zeros(8, dtype=torch.int32, device="cuda")
        check(lib.krs_mod_route(ptr(t), int(dt == torch.int64), 5000, 8, ptr(owner), ptr(local), ptr(counts), stream()))
        o, l = O.mod_route(ids, 8)
        np.testing.assert_array_equal(npy(owner), o)
        np.testing.assert_array_equal(npy(local), l)
        np.testing.assert_array_equal(npy(counts), np.bincount(o, minlength=8))


@pytest.mark.parametrize("engine", ["ffma", "tcgen05"])
@pytest.mark.parametrize("S,nc,nq,d,k", [(2, 10_007, 70, 64, 100), (8, 200_003, 300, 64, 100), (3, 3_001, 33, 32, 10)])
def test_candidate_sharded_retrieval_equals_unsharded(K, S, nc, nq, d, k, engine):
    """Candidates split row-wise into S contiguous shards, local exact top-k per shard with global ids, lists concatenated
    in shard order and merged by krs_row_topk == the unsharded search, scores and ids exactly (same per-pair arithmetic,
    ties -> lowest id)."""
    K.ops.set_topk_engine(engine)          # the same score engine on both sides: identical per-pair arithmetic
    try:
        rng = np.random.default_rng(S)
        q = rng.normal(size=(nq, d)).astype(np.float32)
        c = rng.normal(size=(nc, d)).astype(np.float32)
        c[nc // 2] = c[5]                                      # exact score ties across shards
        tq, tc_ = dev(q), dev(c)
        ref_s, ref_i = K.ops.top_k_scores(tq, tc_, None, k)
        bounds = [nc * s // S for s in range(S + 1)]
        parts = []
        for s in range(S):
            shard = K.layers.CandidateShardedRetrieval(tc_[bounds[s]:bounds[s + 1]], bounds[s], k=k)
            parts.append(shard.local_top_k(tq))
        ms, mi = K.layers.merge_top_k(torch.cat([p[0] for p in parts], dim=1), torch.cat([p[1] for p in parts], dim=1), k)
        np.testing.assert_array_equal(npy(mi), npy(ref_i))
        np.testing.assert_array_equal(npy(ms), npy(ref_s))
        ref = q.astype(np.float64) @ c.astype(np.float64).T
        _check_topk(ref, npy(ms), npy(mi), k)
    finally:
        K.ops.set_topk_engine("auto")


def test_dense_with_a_width_that_is_not_a_multiple_of_4_runs_on_the_tensor_pipe(K):
    """Dense over K = 70 inputs (DLRM's top MLP sees 128 + 351 = 479): zero-padded to 72 inside the autograd function so
    that the tcgen05 engine takes it; shapes and gradients the caller sees are the unpadded ones."""
    rng = np.random.default_rng(8)
    B, Kd, Nn = 256, 70, 32
    x = rng.normal(size=(B, Kd)).astype(np.float32)
    K.set_gemm_engine("tcgen05_ts")
    try:
        layer = K.layers.Dense(Nn, activation="relu")
        tx = dev(x).requires_grad_(True)
        before = K._lib.lib.krs_gemm_tc_launch_count()
        y = layer(tx)
        assert K._lib.lib.krs_gemm_tc_launch_count() > before, "the padded GEMM did not reach the tensor-pipe kernel"
        W, b = npy(layer.kernel), npy(layer.bias)
        z = x.astype(np.float64) @ W.astype(np.float64) + b
        assert_close(npy(y), np.maximum(z, 0), rel=2e-5, what="padded dense fwd")
        g = rng.normal(size=(B, Nn)).astype(np.float32)
        y.backward(dev(g))
        dz = g * (z > 0)
        assert tuple(tx.grad.shape) == (B, Kd) and tuple(layer.kernel.grad.shape) == (Kd, Nn)
        assert_close(npy(tx.grad), dz @ W.T.astype(np.float64), rel=2e-5, what="padded dense dx")
        assert_close(npy(layer.kernel.grad), x.T.astype(np.float64) @ dz, rel=2e-5, what="padded dense dW")
    finally:
        K.set_gemm_engine("ffma")


@pytest.mark.parametrize("self_i,skip", [(False, False), (True, True)])
def test_dot_interaction_packed_equals_the_list_form(K, self_i, skip):
    rng = np.random.default_rng(4)
    B, n, E = 77, 5, 32
    buf = rng.normal(size=(B, n * E)).astype(np.float32)
    t1 = dev(buf).requires_grad_(True)
    feats = [t1[:, j * E:(j + 1) * E] for j in range(n)]
    o1 = K.layers.DotInteraction(self_interaction=self_i, skip_gather=skip)(feats)
    t2 = dev(buf).requires_grad_(True)
    o2 = K.ops.dot_interaction_packed(t2, n, E, self_i, skip)
    np.testing.assert_array_equal(npy(o1), npy(o2))
    g = dev(rng.normal(size=tuple(o1.shape)).astype(np.float32))
    o1.backward(g)
    o2.backward(g)
    np.testing.assert_array_equal(npy(t1.grad), npy(t2.grad))
