"""The numpy DLRM restatement (oracle.np_oracle.dlrm_forward / dlrm_backward, examples/ml_perf/model.py:175-212) against
torch-CPU autograd in float64, for the dot-interaction (BASELINE C3) and the DCN-block (ml_perf) variants."""
import numpy as np
import pytest
import torch

from oracle import np_oracle as O


def _params(rng, vocab, E, nd, bottom, top, interaction, L=2, P=4):
    def lin(i, o):
        return (rng.normal(size=(i, o)) * 0.3, rng.normal(size=(o,)) * 0.1)
    p = dict(tables=[rng.normal(size=(v, E)) * 0.2 for v in vocab], bottom=[], top=[], cross=[])
    k = nd
    for u in bottom:
        p["bottom"].append(lin(k, u)); k = u
    F = len(vocab)
    k = E + (F + 1) * F // 2 if interaction == "dot" else E * (F + 1)
    if interaction == "cross":
        D = k
        p["cross"] = [dict(U=rng.normal(size=(D, P)) * 0.2, V=rng.normal(size=(P, D)) * 0.2, b=rng.normal(size=(D,)) * 0.1) for _ in range(L)]
    for u in top:
        p["top"].append(lin(k, u)); k = u
    return p


def _torch_forward(tp, dense, ids, interaction):
    h = dense
    for W, b in tp["bottom"]:
        h = torch.relu(h @ W + b)
    embs = [t[ids[:, f]] for f, t in enumerate(tp["tables"])]
    if interaction == "dot":
        Fm = torch.stack([h] + embs, dim=1)
        P = Fm @ Fm.transpose(1, 2)
        N = Fm.shape[1]
        ii, jj = torch.tril_indices(N, N, offset=-1)
        x = torch.cat([h, P[:, ii, jj]], dim=-1)
    else:
        x0 = torch.cat([h] + embs, dim=-1)
        x = x0
        for c in tp["cross"]:
            x = x0 * ((x @ c["U"]) @ c["V"] + c["b"]) + x
    n = len(tp["top"])
    for i, (W, b) in enumerate(tp["top"]):
        z = x @ W + b
        x = torch.sigmoid(z) if i == n - 1 else torch.relu(z)
    return x


@pytest.mark.parametrize("interaction", ["dot", "cross"])
def test_dlrm_oracle_matches_torch_autograd(interaction):
    rng = np.random.default_rng(5)
    vocab, E, nd, B = [11, 7, 13], 4, 5, 9
    p = _params(rng, vocab, E, nd, (6, E), (8, 1), interaction)
    dense = rng.uniform(0, 0.9, size=(B, nd))
    ids = np.stack([rng.integers(0, v, size=B) for v in vocab], axis=1)
    dpred = rng.normal(size=(B, 1))
    cache = {}
    pred = O.dlrm_forward(p, dense, ids, interaction, cache)
    g = O.dlrm_backward(p, ids, dpred, cache, interaction)

    T = lambda a: torch.tensor(a, dtype=torch.float64, requires_grad=True)
    tp = dict(tables=[T(t) for t in p["tables"]], bottom=[(T(W), T(b)) for W, b in p["bottom"]],
              top=[(T(W), T(b)) for W, b in p["top"]], cross=[{k: T(v) for k, v in c.items()} for c in p["cross"]])
    tpred = _torch_forward(tp, torch.tensor(dense), torch.tensor(ids), interaction)
    np.testing.assert_allclose(pred, tpred.detach().numpy(), rtol=1e-10, atol=1e-12)
    tpred.backward(torch.tensor(dpred))
    for f in range(len(vocab)):
        np.testing.assert_allclose(g["tables"][f], tp["tables"][f].grad.numpy(), rtol=1e-8, atol=1e-10)
    for name in ("bottom", "top"):
        for (dW, db), (W, b) in zip(g[name], tp[name]):
            np.testing.assert_allclose(dW, W.grad.numpy(), rtol=1e-8, atol=1e-10)
            np.testing.assert_allclose(db, b.grad.numpy(), rtol=1e-8, atol=1e-10)
    for gc, c in zip(g["cross"], tp["cross"]):
        np.testing.assert_allclose(gc["dV"], c["V"].grad.numpy(), rtol=1e-8, atol=1e-10)
        np.testing.assert_allclose(gc["dU"], c["U"].grad.numpy(), rtol=1e-8, atol=1e-10)
        np.testing.assert_allclose(gc["db"], c["b"].grad.numpy(), rtol=1e-8, atol=1e-10)
