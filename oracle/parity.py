"""Parity harness for the row-sharded DCN step — test infrastructure (oracle side), shared by tests/ and by the
self-check `bench.py --gpus N` runs before timing.  It replays K training steps of the GLOBAL batch through np_oracle
(examples/dcn.py wiring; table gradients per jax/test_utils.py:395-417, optimizer rules per np_oracle) and compares the
loss of every step and the final state of every table shard and dense weight with the sharded CUDA model.

Tolerance: 1e-5 of the tensor's max magnitude for losses, activations and one-step updates (the north star's bar).
Adam-type rules divide by sqrt(v) + eps, which turns a 1e-6 relative difference in a tiny gradient into a larger
difference of the update; parameters after K steps are therefore compared at `rel_params` (stated by the caller, 1e-5
for SGD / Adagrad, 5e-5 for AdamW over 3 steps) — the same amplification tests/test_gpu_model.py documents.

Two properties of the comparison that the callers rely on (both measured, benchmarks/sim_parity_probe*.py):
  * Adam-type rules make the check SENSITIVE: their update is ~lr whatever the gradient's size, so a wrong or missing
    gradient row shows as a percent-level parameter error.  SGD / Adagrad updates at lr = 0.01 are ~1e-6 of the
    parameter scale and would hide it — those optimizers are checked through the gradient rows themselves
    (tests/test_gpu_exchange.py::test_compact_gradient_rows_match_oracle).
  * A ReLU hidden layer makes it FRAGILE: when some pre-activation is within rounding of 0, the GPU (whose scatter-add
    order is not fixed) and numpy can disagree on the mask of that one unit, which changes one sample's gradient by a
    finite amount and, through Adam, a few parameters by ~lr.  The shadow configs therefore use a smooth hidden
    activation (tanh); ReLU is exercised with the mask taken from the implementation's own z at kernel level.
"""
from __future__ import annotations

import numpy as np

from . import np_oracle as O


def make_batches(vocab, per_rank_batch, world, steps, seed=100, bad_ids=False):
    out = []
    for step in range(steps):
        rng = np.random.default_rng(seed + step)
        gids = np.stack([rng.integers(0, v, size=per_rank_batch * world) for v in vocab], axis=1).astype(np.int64)
        if bad_ids and gids.shape[0] > 3:
            gids[0, 0] = -1                       # wraps to the last row
            gids[1, -1] = -int(vocab[-1])         # wraps to row 0
            gids[2, 0] = gids[2, 0]               # (a duplicate-prone position is left as is)
        gy = rng.uniform(size=per_rank_batch * world).astype(np.float32)
        out.append((gids, gy))
    return out


def _flat(P):
    return P["tables"] + [a for c in P["cross"] for a in (c["V"], c["b"])] + [a for W, b, _ in P["mlp"] for a in (W, b)]


def _unflat(P, new):
    nt = len(P["tables"])
    P["tables"] = list(new[:nt])
    k = nt
    for c in P["cross"]:
        c["V"], c["b"] = new[k], new[k + 1]
        k += 2
    P["mlp"] = [(new[k + 2 * i], new[k + 2 * i + 1], P["mlp"][i][2]) for i in range(len(P["mlp"]))]


class OracleTrainer:
    """K steps of the DCN training step on the global batch with one of the product's optimizers."""

    def __init__(self, params, optimizer="adamw", lr=0.01, **hyper):
        self.P, self.opt, self.lr, self.hyper = params, optimizer, lr, hyper
        self.nt = len(params["tables"])
        flat = _flat(params)
        self.m = [np.zeros_like(a) for a in flat]
        self.v = [np.zeros_like(a) for a in flat]
        self.acc = [np.full_like(a, hyper.get("initial_accumulator_value", 0.1)) for a in flat]
        self.lin = [np.zeros_like(a) for a in flat]
        self.step = 0

    def train(self, gids, gy):
        self.step += 1
        P = self.P
        cache = {}
        pred = O.dcn_forward(P, gids, cache)
        loss, dpred = O.mse_loss(pred, gy)
        g = O.dcn_backward(P, gids, dpred, cache)
        gl = g["tables"] + [a for c in g["cross"] for a in (c["V"], c["b"])] + [a for dW, db in g["mlp"] for a in (dW, db)]
        new = []
        looked_up = []
        for f, t in enumerate(P["tables"]):
            idx, ok = O.resolve_ids(gids[:, f], t.shape[0])
            mask = np.zeros((t.shape[0],), bool)
            mask[idx[ok]] = True
            looked_up.append(mask)
        for i, (a, ga) in enumerate(zip(_flat(P), gl)):
            table = i < self.nt
            if self.opt == "adamw":
                p2, self.m[i], self.v[i] = O.adamw_step(a, self.m[i], self.v[i], ga, self.step, lr=self.lr)
            elif self.opt == "adagrad":
                p2, self.acc[i] = O.adagrad_step(a, self.acc[i], ga, lr=self.lr)
            elif self.opt == "sgd":
                p2 = O.sgd_step(a, ga, lr=self.lr)
            elif self.opt == "lazy_adam":         # tables: per-row Adam; dense weights: plain Adam
                if table:
                    p2, self.m[i], self.v[i] = O.lazy_adam_step(a, self.m[i], self.v[i], ga, self.step, lr=self.lr, rows=looked_up[i])
                else:
                    p2, self.m[i], self.v[i] = O.adamw_step(a, self.m[i], self.v[i], ga, self.step, lr=self.lr, wd=0.0)
            elif self.opt == "ftrl":              # tables: FTRL on touched rows; dense weights: Adagrad (dense FTRL is out of scope)
                if table:
                    p2, self.acc[i], self.lin[i] = O.ftrl_step(a, self.acc[i], self.lin[i], ga, lr=self.lr, rows=looked_up[i], **{
                        k: v for k, v in self.hyper.items() if k in ("lr_power", "l1", "l2", "beta")})
                else:
                    p2, self.acc[i] = O.adagrad_step(a, self.acc[i], ga, lr=self.lr)
            else:
                raise ValueError(self.opt)
            new.append(p2)
        _unflat(P, new)
        return float(loss)


def params_of(model_tables_global, cross, mlp):
    """model_tables_global: list of unsharded (V, E) numpy tables; cross / mlp: the model's layer objects."""
    npy = lambda t: t.detach().float().cpu().numpy()
    return dict(tables=[t.copy() for t in model_tables_global],
                cross=[dict(V=npy(c.kernel), b=npy(c.bias)) for c in cross],
                mlp=[(npy(d.kernel), npy(d.bias), {0: None, 1: "relu", 2: "sigmoid", 3: "tanh", 4: "swish"}[d._act_id]) for d in mlp])


def max_rel(got, ref):
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    if got.shape != ref.shape:
        return float("inf")
    if not np.isfinite(got).all():
        return float("inf")
    scale = max(float(np.max(np.abs(ref))) if ref.size else 0.0, 1e-30)
    return float(np.max(np.abs(got - ref))) / scale if ref.size else 0.0
