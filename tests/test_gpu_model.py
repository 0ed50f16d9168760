"""GPU parity of the whole DCN-v2 step (gather -> cross stack -> MLP -> loss -> backward -> optimizer)
against the CPU oracle, for the fused C-ABI training path and the public layer/autograd path, plus
size-independent properties at BASELINE.json's full C2 sizes."""
import numpy as np
import pytest
import torch

from oracle import np_oracle as O
from util import assert_close, dev, npy

pytestmark = pytest.mark.gpu


def _oracle_params(model):
    tabs = [npy(t) for t in model.tables()]
    cross = []
    for c in model.cross:
        p = dict(V=npy(c.kernel), b=npy(c.bias), diag_scale=c.diag_scale or 0.0, pre_activation=c._act_name if c._act_id else None)
        if c.down_proj_kernel is not None:
            p["U"] = npy(c.down_proj_kernel)
        cross.append(p)
    mlp = [(npy(d.kernel), npy(d.bias), d._act_name if d._act_id else None) for d in model.mlp]
    return dict(tables=tabs, cross=cross, mlp=mlp)


def _mk(vocab, E, L, P, units, seed=0, **kw):
    from keras_rs_b200.dcn import DCN
    return DCN(vocab, embedding_dim=E, num_cross_layers=L, projection_dim=P, dense_units=units, seed=seed, **kw)


@pytest.mark.parametrize("P", [None, 8])
@pytest.mark.parametrize("L", [1, 3])
def test_fused_step_grads_vs_oracle(P, L):
    rng = np.random.default_rng(0)
    vocab, E, B = [50, 33, 64, 7], 8, 96
    m = _mk(vocab, E, L, P, (16, 16))
    ids = np.stack([rng.integers(0, v, size=B) for v in vocab], axis=1).astype(np.int32)
    ids[1] = ids[0]
    y = rng.uniform(size=B).astype(np.float32)
    params = _oracle_params(m)
    cache = {}
    pred = O.dcn_forward(params, ids, cache)
    loss_ref, dpred = O.mse_loss(pred, y)
    gref = O.dcn_backward(params, ids, dpred, cache)

    loss = m.forward_backward(dev(ids), dev(y))
    np.testing.assert_allclose(float(loss), float(loss_ref), rtol=1e-5)
    assert_close(npy(m.predict(dev(ids))), pred, what="pred")
    for f, v in enumerate(vocab):
        got = npy(m.emb_grad[m.row_off[f]:m.row_off[f] + v])
        assert_close(got, gref["tables"][f], what=f"table grad {f}")
    for c, g in zip(m.cross, gref["cross"]):
        assert_close(npy(m._g(c.kernel)), g["V"], what="dV")
        assert_close(npy(m._g(c.bias)), g["b"], what="db")
        if P is not None:
            assert_close(npy(m._g(c.down_proj_kernel)), g["U"], what="dU")
    for d, (dW, db) in zip(m.mlp, gref["mlp"]):
        assert_close(npy(m._g(d.kernel)), dW, what="mlp dW")
        assert_close(npy(m._g(d.bias)), db, what="mlp db")


def test_layer_autograd_path_matches_fused_path():
    rng = np.random.default_rng(1)
    vocab, E, B = [40, 40, 40], 32, 64
    m = _mk(vocab, E, 2, None, (32,))
    ids = np.stack([rng.integers(0, v, size=B) for v in vocab], axis=1).astype(np.int32)
    y = rng.uniform(size=B).astype(np.float32)
    m.forward_backward(dev(ids), dev(y))
    fused_emb = m.emb_grad.clone()
    fused_dense = m.dense_grad_flat.clone()
    m.emb_grad.zero_(); m.emb_touched.zero_()
    import keras_rs_b200 as K
    pred = m.forward(dev(ids), sparse_arena=True)
    loss = K.ops.loss_fn(pred, dev(y), "mse")
    loss.backward()
    assert_close(npy(m.emb_grad), npy(fused_emb), rel=2e-6, what="emb grad (autograd vs fused)")
    for p in m.dense_params():
        assert_close(npy(p.grad), npy(m._g(p)), rel=2e-6, what="dense grad (autograd vs fused)")
    assert fused_dense.numel() == m.dense_flat.numel()


@pytest.mark.parametrize("opt_name", ["adamw", "adagrad", "sgd"])
def test_training_steps_vs_oracle(opt_name):
    import keras_rs_b200 as K
    rng = np.random.default_rng(2)
    vocab, E, B = [30, 20], 8, 64
    m = _mk(vocab, E, 2, None, (16,))
    opt = {"adamw": K.optimizers.AdamW(0.01), "adagrad": K.optimizers.Adagrad(0.05), "sgd": K.optimizers.SGD(0.05)}[opt_name]
    params = _oracle_params(m)
    flat = lambda P: ([t for t in P["tables"]] + [a for c in P["cross"] for a in (c["V"], c["b"])] +
                      [a for W, b, _ in P["mlp"] for a in (W, b)])
    state = [dict(m=np.zeros_like(a), v=np.zeros_like(a), acc=np.full_like(a, 0.1)) for a in flat(params)]
    for step in range(1, 4):
        ids = np.stack([rng.integers(0, v, size=B) for v in vocab], axis=1).astype(np.int32)
        y = rng.uniform(size=B).astype(np.float32)
        cache = {}
        pred = O.dcn_forward(params, ids, cache)
        loss_ref, dpred = O.mse_loss(pred, y)
        g = O.dcn_backward(params, ids, dpred, cache)
        gl = [t for t in g["tables"]] + [a for c in g["cross"] for a in (c["V"], c["b"])] + [a for dW, db in g["mlp"] for a in (dW, db)]
        new = []
        for a, ga, st in zip(flat(params), gl, state):
            if opt_name == "adamw":
                p2, st["m"], st["v"] = O.adamw_step(a, st["m"], st["v"], ga, step, lr=0.01)
            elif opt_name == "adagrad":
                p2, st["acc"] = O.adagrad_step(a, st["acc"], ga, lr=0.05)
            else:
                p2 = O.sgd_step(a, ga, 0.05)
            new.append(p2)
        nt = len(params["tables"])
        params["tables"] = new[:nt]
        k = nt
        for c in params["cross"]:
            c["V"], c["b"] = new[k], new[k + 1]
            k += 2
        params["mlp"] = [(new[k + 2 * i], new[k + 2 * i + 1], params["mlp"][i][2]) for i in range(len(params["mlp"]))]
        loss = m.train_on_batch(dev(ids), dev(y), opt)
        np.testing.assert_allclose(float(loss), float(loss_ref), rtol=2e-4)
    got = _oracle_params(m)
    for a, b in zip(flat(got), flat(params)):
        assert_close(a, b, rel=2e-4, what=f"params after 3 {opt_name} steps")
    assert float(m.emb_grad.abs().max()) == 0.0 and int(m.emb_touched.abs().max()) == 0


def test_c1_readme_toy_config():
    """BASELINE configs[0]: vocab=32 embed_dim=6 batch=2 (README.md:43-75)."""
    rng = np.random.default_rng(0)
    m = _mk([32], 6, 2, None, ())
    ids = rng.integers(0, 32, size=(2, 1)).astype(np.int32)
    y = rng.uniform(size=2).astype(np.float32)
    params = _oracle_params(m)
    cache = {}
    pred = O.dcn_forward(params, ids, cache)
    loss_ref, dpred = O.mse_loss(pred, y)
    g = O.dcn_backward(params, ids, dpred, cache)
    loss = m.forward_backward(dev(ids), dev(y))
    np.testing.assert_allclose(float(loss), float(loss_ref), rtol=1e-5)
    assert_close(npy(m.emb_grad[:32]), g["tables"][0], what="C1 table grad")
    assert_close(npy(m._g(m.cross[0].kernel)), g["cross"][0]["V"], what="C1 dV")


# ------------------------------------------------------------------ full-size properties (C2 shapes)
def test_c2_full_size_gather_properties():
    """B=65536 x 26 features, V=1e6, E=32: the oracle cannot finish in seconds here, so check
    size-independent properties: every output row equals the addressed table row (bit exact, checked
    by an independent torch index on the same device) and scatter-add conserves the gradient sum."""
    import keras_rs_b200 as K
    F, V, E, B = 26, 1_000_000, 32, 65536
    g = torch.Generator(device="cuda").manual_seed(1234)
    arena = torch.rand((F * V, E), device="cuda", generator=g) * 0.1 - 0.05
    ids = torch.randint(0, V, (B, F), device="cuda", generator=g, dtype=torch.int32)
    tabs = [arena[f * V:(f + 1) * V] for f in range(F)]
    plan = K.ops.GatherPlan([dict(table=tabs[f], ids=ids[:, f], combiner="sum") for f in range(F)])
    for variant in (0, 2):
        out = plan.forward(variant=variant)
        flat_rows = (ids.long() + torch.arange(F, device="cuda") * V).reshape(-1)
        ref = arena[flat_rows].reshape(B, F * E)
        assert torch.equal(out, ref), f"variant {variant}"
    gout = torch.randn((B, F * E), device="cuda", generator=g)
    grad = torch.zeros_like(arena)
    touched = torch.zeros((F * V // 32,), dtype=torch.int32, device="cuda")
    plan.backward(gout, [grad[f * V:] for f in range(F)], [touched[f * V // 32:] for f in range(F)])
    # conservation: column sums of the dense gradient == column sums of gout per feature
    got = grad.reshape(F, V, E).double().sum(dim=1)
    exp = gout.reshape(B, F, E).double().sum(dim=0)
    assert float((got - exp).abs().max()) < 1e-3 * float(exp.abs().max() + 1)
    # popcount(touched) == number of unique (feature,row) pairs
    uniq = torch.unique(flat_rows).numel()
    bits = touched.view(torch.int32)
    pop = 0
    x = bits.clone().to(torch.int64) & 0xFFFFFFFF
    while int(x.max()) > 0:
        pop += int((x & 1).sum())
        x >>= 1
    assert pop == uniq


@pytest.mark.parametrize("opt_name", ["adamw", "adagrad"])
def test_cuda_graph_step_matches_eager_step(opt_name):
    """train_on_batch_graph (one CUDA-graph replay per step, device-resident Adam step/alpha) must produce the
    same parameters as the eager launch sequence."""
    import keras_rs_b200 as K
    rng = np.random.default_rng(9)
    vocab, E, B = [64, 40, 33], 32, 128
    mk_opt = (lambda: K.optimizers.AdamW(0.01)) if opt_name == "adamw" else (lambda: K.optimizers.Adagrad(0.05))
    m1, m2 = _mk(vocab, E, 2, None, (16,), seed=4), _mk(vocab, E, 2, None, (16,), seed=4)
    o1, o2 = mk_opt(), mk_opt()
    for step in range(5):
        ids = np.stack([rng.integers(0, v, size=B) for v in vocab], axis=1).astype(np.int32)
        y = rng.uniform(size=B).astype(np.float32)
        l1 = float(m1.train_on_batch(dev(ids), dev(y), o1))
        l2 = float(m2.train_on_batch_graph(dev(ids), dev(y), o2))
        np.testing.assert_allclose(l1, l2, rtol=1e-6)
    assert o1.iterations == o2.iterations == 5
    assert_close(npy(m2.emb), npy(m1.emb), rel=1e-6, what="emb after 5 graph steps")
    assert_close(npy(m2.dense_flat), npy(m1.dense_flat), rel=1e-6, what="dense params after 5 graph steps")


@pytest.mark.parametrize("opt_name", ["adamw", "adagrad", "sgd"])
@pytest.mark.parametrize("sparse_arena", [False, True])
def test_layer_api_forward_backward_apply_updates_the_tables(opt_name, sparse_arena):
    """The drop-in path a user writes — model(ids) -> loss.backward() -> optimizer.apply(model.parameters()) — must update the
    embedding tables from whichever gradient the backward produced (dense .grad by default, the arena with
    sparse_arena=True), for every optimizer (round-1 advisor finding: the arena was preferred even when it was empty)."""
    import keras_rs_b200 as K
    from oracle import parity as PAR
    rng = np.random.default_rng(9)
    vocab, E, B = [40, 23, 64], 8, 64
    m = _mk(vocab, E, 2, None, (16,), seed=4, dense_activation="tanh")
    init = [npy(t).copy() for t in m.tables()]
    tr = PAR.OracleTrainer(PAR.params_of(init, m.cross, m.mlp), opt_name, lr=0.05)
    opt = {"adamw": K.optimizers.AdamW, "adagrad": K.optimizers.Adagrad, "sgd": K.optimizers.SGD}[opt_name](0.05)
    for step in range(3):
        ids = np.stack([rng.integers(0, v, size=B) for v in vocab], axis=1).astype(np.int32)
        y = rng.uniform(size=B).astype(np.float32)
        ref_loss = tr.train(ids.astype(np.int64), y)
        params = list(m.parameters())
        opt.zero_grad(params)
        pred = m(dev(ids), sparse_arena=sparse_arena)
        loss = K.ops.loss_fn(pred, dev(y), "mse")
        loss.backward()
        opt.apply(params)
        assert abs(float(loss) - ref_loss) <= 1e-5 * max(abs(ref_loss), 1e-6)
    rel = 5e-5 if opt_name == "adamw" else 1e-5
    for f, t in enumerate(m.tables()):
        assert_close(npy(t), tr.P["tables"][f], rel=rel, what=f"table {f} ({opt_name}, arena={sparse_arena})")
        assert float(np.abs(npy(t) - init[f]).max()) > 1e-4            # the tables really moved
    for c, pc in zip(m.cross, tr.P["cross"]):
        assert_close(npy(c.kernel), pc["V"], rel=rel, what="cross V")


def test_dense_swish_trains():
    """Dense(activation='swish') forward and backward (round-1 advisor finding: the backward raised)."""
    import keras_rs_b200 as K
    rng = np.random.default_rng(2)
    x = rng.normal(size=(33, 12)).astype(np.float32)
    layer = K.layers.Dense(7, activation="swish")
    tx = dev(x).requires_grad_(True)
    y = layer(tx)
    W, b = npy(layer.kernel), npy(layer.bias)
    z = x @ W + b
    assert_close(npy(y), z / (1.0 + np.exp(-z)), rel=1e-5, what="swish forward")
    y.sum().backward()
    s = 1.0 / (1.0 + np.exp(-z))
    dz = s * (1.0 + z * (1.0 - s))
    assert_close(npy(tx.grad), dz @ W.T, rel=1e-5, what="swish dx")
    assert_close(npy(layer.kernel.grad), x.T @ dz, rel=1e-5, what="swish dW")


def test_sharded_model_rejects_unaligned_embedding_dim():
    from keras_rs_b200.sharded import ShardedDCN
    with pytest.raises(ValueError, match="multiple of 4"):
        ShardedDCN([10, 10], rank=0, world=1, sim=True, embedding_dim=6)
