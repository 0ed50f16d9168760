"""Backward formulas of the oracle (analytic, unpinned by the reference: SURVEY §8c) cross-checked
against torch-CPU autograd in float64-free fp32."""
import numpy as np
import pytest
import torch

from oracle import np_oracle as O

ACT = {None: lambda z: z, "relu": torch.relu, "sigmoid": torch.sigmoid, "tanh": torch.tanh,
       "swish": torch.nn.functional.silu}


@pytest.mark.parametrize("P", [None, 3])
@pytest.mark.parametrize("act", [None, "relu", "sigmoid", "tanh", "swish"])
@pytest.mark.parametrize("diag", [0.0, 0.5])
def test_feature_cross_bwd(P, act, diag):
    rng = np.random.default_rng(0)
    B, D = 5, 7
    x0 = rng.normal(size=(B, D)).astype(np.float32)
    x = rng.normal(size=(B, D)).astype(np.float32)
    U = None if P is None else rng.normal(size=(D, P)).astype(np.float32)
    V = rng.normal(size=(D if P is None else P, D)).astype(np.float32)
    b = rng.normal(size=(D,)).astype(np.float32)
    gy = rng.normal(size=(B, D)).astype(np.float32)
    r = O.feature_cross_bwd(gy, x0, x, V, b, U, diag, act)
    t = {k: torch.tensor(v, requires_grad=True) for k, v in dict(x0=x0, x=x, V=V, b=b).items()}
    tU = None if U is None else torch.tensor(U, requires_grad=True)
    h = t["x"] if tU is None else t["x"] @ tU
    a = ACT[act](h @ t["V"] + t["b"])
    y = t["x0"] * (a + diag * t["x"]) + t["x"]
    np.testing.assert_allclose(y.detach().numpy(), O.feature_cross(x0, x, V, b, U, diag, act), rtol=1e-5, atol=1e-5)
    y.backward(torch.tensor(gy))
    for k, tk in (("dx0", t["x0"]), ("dx", t["x"]), ("dV", t["V"]), ("db", t["b"])):
        np.testing.assert_allclose(r[k], tk.grad.numpy(), rtol=2e-4, atol=2e-5)
    if tU is not None:
        np.testing.assert_allclose(r["dU"], tU.grad.numpy(), rtol=2e-4, atol=2e-5)


@pytest.mark.parametrize("self_i", [False, True])
@pytest.mark.parametrize("skip", [False, True])
def test_dot_bwd(self_i, skip):
    rng = np.random.default_rng(1)
    B, N, E = 3, 5, 4
    inp = [rng.normal(size=(B, E)).astype(np.float32) for _ in range(N)]
    out = O.dot_interaction(inp, self_i, skip)
    g = rng.normal(size=out.shape).astype(np.float32)
    d = O.dot_interaction_bwd(g, inp, self_i, skip)
    ts = [torch.tensor(a, requires_grad=True) for a in inp]
    F = torch.stack(ts, 1)
    Pm = F @ F.transpose(1, 2)
    if skip:
        mask = torch.tril(torch.ones(N, N), diagonal=0 if self_i else -1)
        o = (Pm * mask).reshape(B, N * N)
    else:
        o = Pm.reshape(B, N * N)[:, O.tril_indices(N, self_i)]
    np.testing.assert_allclose(o.detach().numpy(), out, rtol=1e-5, atol=1e-5)
    o.backward(torch.tensor(g))
    for a, t in zip(d, ts):
        np.testing.assert_allclose(a, t.grad.numpy(), rtol=1e-4, atol=1e-5)


def _mk_params(rng, F, V, E, L, P, units):
    D = F * E
    tables = [rng.uniform(-0.05, 0.05, size=(V, E)).astype(np.float32) for _ in range(F)]
    cross = []
    for _ in range(L):
        p = dict(V=O.glorot_uniform(rng, D if P is None else P, D), b=rng.normal(size=(D,)).astype(np.float32) * 0.1)
        if P is not None:
            p["U"] = O.glorot_uniform(rng, D, P)
        cross.append(p)
    mlp = []
    k = D
    for u in units:
        mlp.append((O.glorot_uniform(rng, k, u), np.zeros((u,), np.float32), "relu"))
        k = u
    mlp.append((O.glorot_uniform(rng, k, 1), np.zeros((1,), np.float32), None))
    return dict(tables=tables, cross=cross, mlp=mlp)


@pytest.mark.parametrize("P", [None, 4])
def test_dcn_model_bwd_vs_autograd(P):
    rng = np.random.default_rng(2)
    F, V, E, L, B = 3, 11, 4, 2, 6
    params = _mk_params(rng, F, V, E, L, P, [8, 8])
    ids = rng.integers(0, V, size=(B, F)).astype(np.int32)
    ids[1] = ids[0]   # duplicates
    y = rng.uniform(size=(B,)).astype(np.float32)
    cache = {}
    pred = O.dcn_forward(params, ids, cache)
    loss, dpred = O.mse_loss(pred, y)
    g = O.dcn_backward(params, ids, dpred, cache)
    # torch autograd mirror
    tt = [torch.tensor(t, requires_grad=True) for t in params["tables"]]
    tc = [{k: torch.tensor(v, requires_grad=True) for k, v in p.items()} for p in params["cross"]]
    tm = [(torch.tensor(W, requires_grad=True), torch.tensor(b, requires_grad=True), a) for W, b, a in params["mlp"]]
    x0 = torch.cat([torch.nn.functional.embedding(torch.tensor(ids[:, f]).long(), tt[f]) for f in range(F)], 1)
    xl = x0
    for p in tc:
        h = xl if "U" not in p else xl @ p["U"]
        xl = x0 * (h @ p["V"] + p["b"]) + xl
    h = xl
    for W, b, a in tm:
        h = h @ W + b
        if a == "relu":
            h = torch.relu(h)
    tl = torch.mean((h.reshape(-1) - torch.tensor(y)) ** 2)
    np.testing.assert_allclose(float(tl.detach()), float(loss), rtol=1e-5)
    tl.backward()
    for a, t in zip(g["tables"], tt):
        np.testing.assert_allclose(a, t.grad.numpy(), rtol=1e-3, atol=1e-6)
    for a, t in zip(g["cross"], tc):
        np.testing.assert_allclose(a["V"], t["V"].grad.numpy(), rtol=1e-3, atol=1e-6)
        np.testing.assert_allclose(a["b"], t["b"].grad.numpy(), rtol=1e-3, atol=1e-6)
        if "U" in t:
            np.testing.assert_allclose(a["U"], t["U"].grad.numpy(), rtol=1e-3, atol=1e-6)
    for (dW, db), (W, b, _) in zip(g["mlp"], tm):
        np.testing.assert_allclose(dW, W.grad.numpy(), rtol=1e-3, atol=1e-6)
        np.testing.assert_allclose(db, b.grad.numpy(), rtol=1e-3, atol=1e-6)


def test_adamw_matches_keras_formula_scalar():
    # hand evaluation of the Keras-3 AdamW rule for one scalar, two steps
    p, m, v = np.float32(1.0), np.float32(0.0), np.float32(0.0)
    lr, b1, b2, eps, wd = 0.01, 0.9, 0.999, 1e-7, 0.004
    pp, mm, vv = np.array([p]), np.array([m]), np.array([v])
    ref_p, ref_m, ref_v = 1.0, 0.0, 0.0
    for t, g in enumerate([0.5, -0.25], start=1):
        pp, mm, vv = O.adamw_step(pp, mm, vv, np.array([g], np.float32), t, lr, b1, b2, eps, wd)
        ref_p -= ref_p * wd * lr
        ref_m += (g - ref_m) * (1 - b1)
        ref_v += (g * g - ref_v) * (1 - b2)
        alpha = lr * np.sqrt(1 - b2 ** t) / (1 - b1 ** t)
        ref_p -= ref_m * alpha / (np.sqrt(ref_v) + eps)
    np.testing.assert_allclose(pp[0], ref_p, rtol=1e-5)
