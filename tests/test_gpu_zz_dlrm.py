"""DLRM (keras_rs_b200/dlrm.py: examples/ml_perf/model.py wiring) vs the numpy oracle: forward, every gradient and one
Adagrad step, for the dot-interaction (BASELINE C3) and the DCN-block (ml_perf) variants."""
import numpy as np
import pytest
import torch

from oracle import np_oracle as O
from util import assert_close, dev, npy

pytestmark = pytest.mark.gpu


def _copy(param, value):
    with torch.no_grad():
        param.copy_(dev(np.asarray(value, dtype=np.float32)))


@pytest.mark.parametrize("interaction", ["dot", "cross"])
def test_dlrm_matches_oracle(interaction):
    import keras_rs_b200 as K
    from keras_rs_b200.dlrm import DLRM
    rng = np.random.default_rng(17)
    vocab, E, nd, B = [50, 33, 64, 7], 32, 13, 96
    m = DLRM(vocab, embedding_dim=E, num_dense=nd, bottom_mlp_dims=(16, E), top_mlp_dims=(24, 1), interaction=interaction,
             num_dcn_layers=2, dcn_projection_dim=8, seed=3)
    f32 = lambda a: a.astype(np.float32)
    p = dict(tables=[f32(rng.normal(size=(v, E)) * 0.2) for v in vocab], bottom=[], top=[], cross=[])
    for d in m.bottom_mlp:
        p["bottom"].append((f32(rng.normal(size=tuple(d.kernel.shape)) * 0.3), f32(rng.normal(size=(d.units,)) * 0.1)))
    for d in m.top_mlp:
        p["top"].append((f32(rng.normal(size=tuple(d.kernel.shape)) * 0.2), f32(rng.normal(size=(d.units,)) * 0.1)))
    for c in m.cross:
        p["cross"].append(dict(U=f32(rng.normal(size=tuple(c.down_proj_kernel.shape)) * 0.1),
                               V=f32(rng.normal(size=tuple(c.kernel.shape)) * 0.1), b=f32(rng.normal(size=tuple(c.bias.shape)) * 0.1)))
    for t, v in zip(m.tables, p["tables"]):
        _copy(t, v)
    for layers, vals in ((m.bottom_mlp, p["bottom"]), (m.top_mlp, p["top"])):
        for d, (W, b) in zip(layers, vals):
            _copy(d.kernel, W); _copy(d.bias, b)
    for c, v in zip(m.cross, p["cross"]):
        _copy(c.down_proj_kernel, v["U"]); _copy(c.kernel, v["V"]); _copy(c.bias, v["b"])
    dense = f32(rng.uniform(0, 0.9, size=(B, nd)))
    ids = np.stack([rng.integers(0, v, size=B) for v in vocab], axis=1).astype(np.int32)
    y = f32(rng.integers(0, 2, size=B))

    cache = {}
    pred_ref = O.dlrm_forward(p, dense, ids, interaction, cache)
    loss_ref, dpred = O.bce_loss(pred_ref, y)
    g = O.dlrm_backward(p, ids, dpred.reshape(B, 1).astype(np.float32), cache, interaction)

    pred = m(dev(dense), dev(ids))
    assert tuple(pred.shape) == (B, 1)
    assert_close(npy(pred), pred_ref, what="dlrm forward")
    loss = K.ops.loss_fn(pred, dev(y), "bce")
    np.testing.assert_allclose(float(loss), float(loss_ref), rtol=1e-5)
    loss.backward()
    for f in range(len(vocab)):
        assert_close(npy(m.tables[f].grad), g["tables"][f], what=f"table {f} grad", scale=max(np.abs(g["tables"][f]).max(), 1e-6))
    for name, layers in (("bottom", m.bottom_mlp), ("top", m.top_mlp)):
        for i, d in enumerate(layers):
            assert_close(npy(d.kernel.grad), g[name][i][0], what=f"{name} {i} dW")
            assert_close(npy(d.bias.grad), g[name][i][1], what=f"{name} {i} db", scale=max(np.abs(g[name][i][1]).max(), 1e-6))
    for i, c in enumerate(m.cross):
        assert_close(npy(c.kernel.grad), g["cross"][i]["dV"], what=f"cross {i} dV")
        assert_close(npy(c.down_proj_kernel.grad), g["cross"][i]["dU"], what=f"cross {i} dU")
        assert_close(npy(c.bias.grad), g["cross"][i]["db"], what=f"cross {i} db", scale=max(np.abs(g["cross"][i]["db"]).max(), 1e-6))


def test_dlrm_train_step_moves_every_parameter():
    import keras_rs_b200 as K
    from keras_rs_b200.dlrm import DLRM
    rng = np.random.default_rng(2)
    vocab, E, B = [40, 20, 30], 32, 64
    m = DLRM(vocab, embedding_dim=E, bottom_mlp_dims=(16, E), top_mlp_dims=(16, 1), interaction="dot", seed=1)
    before = [q.detach().clone() for q in m.parameters_list()]
    opt = K.optimizers.Adagrad(0.05)
    dense = dev(rng.uniform(0, 0.9, size=(B, 13)).astype(np.float32))
    ids = dev(np.stack([rng.integers(0, v, size=B) for v in vocab], axis=1).astype(np.int32))
    y = dev(rng.integers(0, 2, size=B).astype(np.float32))
    l0 = float(m.train_on_batch(dense, ids, y, opt))
    for _ in range(20):
        l1 = float(m.train_on_batch(dense, ids, y, opt))
    assert np.isfinite(l0) and l1 < l0                                   # fitting one batch reduces its loss
    assert all(float((a - b.detach()).abs().max()) > 0 for a, b in zip(before, m.parameters_list()))


@pytest.mark.parametrize("opt_name", ["adagrad", "adamw", "sgd"])
@pytest.mark.parametrize("interaction", ["dot", "cross"])
def test_dlrm_cuda_graph_step_matches_eager_step(interaction, opt_name):
    """train_on_batch_graph (one graph replay per step) must leave the same parameters as the eager launch sequence."""
    import keras_rs_b200 as K
    from keras_rs_b200.dlrm import DLRM
    rng = np.random.default_rng(5)
    vocab, E, B = [40, 20, 30], 32, 64
    mk = lambda: DLRM(vocab, embedding_dim=E, bottom_mlp_dims=(16, E), top_mlp_dims=(16, 1), interaction=interaction,
                      num_dcn_layers=2, dcn_projection_dim=8, seed=1)
    mk_opt = {"adagrad": lambda: K.optimizers.Adagrad(0.05), "adamw": lambda: K.optimizers.AdamW(0.01),
              "sgd": lambda: K.optimizers.SGD(0.05)}[opt_name]
    m1, m2, o1, o2 = mk(), mk(), mk_opt(), mk_opt()
    for step in range(5):
        dense = dev(rng.uniform(0, 0.9, size=(B, 13)).astype(np.float32))
        ids = dev(np.stack([rng.integers(0, v, size=B) for v in vocab], axis=1).astype(np.int32))
        y = dev(rng.integers(0, 2, size=B).astype(np.float32))
        l1 = float(m1.train_on_batch(dense, ids, y, o1))
        l2 = float(m2.train_on_batch_graph(dense, ids, y, o2))
        np.testing.assert_allclose(l1, l2, rtol=1e-5)
    assert o1.iterations == o2.iterations == 5
    for a, b in zip(m1.parameters_list(), m2.parameters_list()):
        assert_close(npy(b), npy(a), rel=1e-5, what=f"{interaction}/{opt_name} parameter after 5 graph steps")


def test_dlrm_cuda_graph_step_refuses_lazy_adam():
    import keras_rs_b200 as K
    from keras_rs_b200.dlrm import DLRM
    m = DLRM([8, 8], embedding_dim=32, bottom_mlp_dims=(16, 32), top_mlp_dims=(16, 1), seed=1)
    z = lambda *s: torch.zeros(s, device="cuda")
    with pytest.raises(ValueError, match="lazy Adam"):
        m.train_on_batch_graph(z(4, 13), torch.zeros((4, 2), dtype=torch.int32, device="cuda"), z(4), K.optimizers.Adam(sparse_rows=True))
