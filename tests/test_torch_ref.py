"""The torch-CPU restatement used as the CPU baseline is itself pinned to the numpy oracle."""
import numpy as np
import pytest
import torch

from oracle import np_oracle as O
from oracle import torch_ref as T


@pytest.mark.parametrize("opt", ["adamw", "adagrad", "sgd"])
def test_torch_ref_matches_numpy_oracle(opt):
    rng = np.random.default_rng(0)
    F, V, E, L, B = 3, 20, 4, 2, 16
    tables, cross, mlp = T.synthetic_c2(F, V, E, L, (8,), seed=3)
    for c in cross:
        c["b"] = rng.normal(size=c["b"].shape).astype(np.float32) * 0.1
    model = T.TorchDCN(tables, cross, mlp, lr=0.01, optimizer=opt)
    params = dict(tables=[t.copy() for t in tables], cross=[dict(V=c["V"].copy(), b=c["b"].copy()) for c in cross],
                  mlp=[(W.copy(), b.copy(), a) for W, b, a in mlp])
    flat = lambda P: P["tables"] + [a for c in P["cross"] for a in (c["V"], c["b"])] + [a for W, b, _ in P["mlp"] for a in (W, b)]
    st = [dict(m=np.zeros_like(a), v=np.zeros_like(a), acc=np.full_like(a, 0.1)) for a in flat(params)]
    for step in range(1, 3):
        ids = rng.integers(0, V, size=(B, F))
        y = rng.uniform(size=B).astype(np.float32)
        cache = {}
        pred = O.dcn_forward(params, ids, cache)
        loss_ref, dpred = O.mse_loss(pred, y)
        g = O.dcn_backward(params, ids, dpred, cache)
        gl = g["tables"] + [a for c in g["cross"] for a in (c["V"], c["b"])] + [a for dW, db in g["mlp"] for a in (dW, db)]
        new = []
        for a, ga, s in zip(flat(params), gl, st):
            if opt == "adamw":
                p2, s["m"], s["v"] = O.adamw_step(a, s["m"], s["v"], ga, step, lr=0.01)
            elif opt == "adagrad":
                p2, s["acc"] = O.adagrad_step(a, s["acc"], ga, lr=0.01)
            else:
                p2 = O.sgd_step(a, ga, 0.01)
            new.append(p2)
        params["tables"] = new[:F]
        k = F
        for c in params["cross"]:
            c["V"], c["b"] = new[k], new[k + 1]
            k += 2
        params["mlp"] = [(new[k + 2 * i], new[k + 2 * i + 1], params["mlp"][i][2]) for i in range(len(params["mlp"]))]
        loss = model.train_step(torch.tensor(ids), torch.tensor(y))
        np.testing.assert_allclose(loss, float(loss_ref), rtol=1e-5)
    for a, p in zip(flat(params), model.params):
        np.testing.assert_allclose(p.detach().numpy(), a, rtol=2e-4, atol=1e-6)
