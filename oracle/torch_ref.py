"""CPU ORACLE (port) — torch-CPU restatement of the reference's DCN training step, op for op as
Keras 3 would issue it under a CPU backend.  Test infrastructure / CPU baseline only (see
oracle/np_oracle.py header): imported by tests/ and by bench.py's cpu_baseline / --impl reference
leg, never by the product.

Op sequence mirrored (examples/dcn.py:418-449, 120-146; README.md:54-55):
  26 x keras.layers.Embedding (ops.take)  -> ops.concatenate(axis=1) -> FeatureCross x L
  (feature_cross.py:182-194: Dense matmul + bias, multiply, add) -> Dense(relu) x n -> Dense(1)
  -> MeanSquaredError -> backend autograd (dense (V,E) table gradients) -> AdamW on EVERY
  variable (Keras 3 rule: decoupled decay, folded bias correction, eps outside the sqrt).
It is validated against the numpy oracle in tests/test_torch_ref.py.  It is NOT the reference's own
JAX run (keras / jax are not installable here: SURVEY F3) and every number produced from it is
labelled kind="port".
"""
from __future__ import annotations

import math
import time

import numpy as np
import torch


class TorchDCN:
    def __init__(self, tables, cross, mlp, lr=0.01, wd=0.004, b1=0.9, b2=0.999, eps=1e-7, optimizer="adamw"):
        """tables: list of (V,E) arrays; cross: list of dict(V, b, U?); mlp: list of (W, b, act)."""
        t = lambda a: torch.tensor(np.asarray(a), dtype=torch.float32, requires_grad=True)
        self.tables = [t(a) for a in tables]
        self.cross = [{k: t(v) for k, v in c.items() if k in ("V", "b", "U") and v is not None} for c in cross]
        self.mlp = [(t(W), t(b), act) for W, b, act in mlp]
        self.params = list(self.tables) + [v for c in self.cross for v in c.values()] + [p for W, b, _ in self.mlp for p in (W, b)]
        self.lr, self.wd, self.b1, self.b2, self.eps = lr, wd, b1, b2, eps
        self.optimizer = optimizer
        self.m = [torch.zeros_like(p) for p in self.params] if optimizer == "adamw" else None
        self.v = [torch.zeros_like(p) for p in self.params] if optimizer == "adamw" else None
        self.acc = [torch.full_like(p, 0.1) for p in self.params] if optimizer == "adagrad" else None
        self.step_no = 0

    def forward(self, ids: torch.Tensor) -> torch.Tensor:
        embs = [torch.nn.functional.embedding(ids[:, f], tab) for f, tab in enumerate(self.tables)]
        x0 = torch.cat(embs, dim=1)
        xl = x0
        for c in self.cross:
            h = xl if "U" not in c else xl @ c["U"]
            xl = x0 * (h @ c["V"] + c["b"]) + xl
        h = xl
        for W, b, act in self.mlp:
            h = h @ W + b
            if act == "relu":
                h = torch.relu(h)
        return h

    def train_step(self, ids: torch.Tensor, labels: torch.Tensor) -> float:
        for p in self.params:
            p.grad = None
        pred = self.forward(ids)
        loss = torch.mean((pred.reshape(-1) - labels) ** 2)
        loss.backward()
        self.step_no += 1
        t = self.step_no
        with torch.no_grad():
            if self.optimizer == "adamw":
                alpha = self.lr * math.sqrt(1 - self.b2 ** t) / (1 - self.b1 ** t)
                for p, m, v in zip(self.params, self.m, self.v):
                    g = p.grad
                    p.sub_(p * (self.wd * self.lr))
                    m.add_((g - m) * (1 - self.b1))
                    v.add_((g * g - v) * (1 - self.b2))
                    p.sub_((m * alpha) / (torch.sqrt(v) + self.eps))
            elif self.optimizer == "adagrad":
                for p, a in zip(self.params, self.acc):
                    g = p.grad
                    a.add_(g * g)
                    p.sub_(self.lr * g / torch.sqrt(a + self.eps))
            else:
                for p in self.params:
                    p.sub_(self.lr * p.grad)
        return float(loss.detach())


def synthetic_c2(F=26, V=1_000_000, E=32, L=3, units=(192, 192), seed=1234):
    """C2 weights: tables ~U(-0.05,0.05), glorot-uniform kernels, zero biases."""
    g = torch.Generator().manual_seed(seed)
    D = F * E
    tables = [(torch.rand((V, E), generator=g) * 0.1 - 0.05).numpy() for _ in range(F)]

    def glorot(i, o):
        lim = math.sqrt(6.0 / (i + o))
        return ((torch.rand((i, o), generator=g) * 2 - 1) * lim).numpy()

    cross = [dict(V=glorot(D, D), b=np.zeros((D,), np.float32)) for _ in range(L)]
    mlp, k = [], D
    for u in units:
        mlp.append((glorot(k, u), np.zeros((u,), np.float32), "relu"))
        k = u
    mlp.append((glorot(k, 1), np.zeros((1,), np.float32), None))
    return tables, cross, mlp


def usable_cores() -> int:
    """Host threads this process can really run at once: min(cpu_count, scheduler affinity, cgroup CPU quota).  A box
    that reports 128 CPUs but grants a 16-CPU quota ran the 128-thread baseline 15x slower than 16 threads."""
    import os
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except (AttributeError, OSError):
        pass
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            txt = open(path).read().split()
            if path.endswith("cpu.max"):
                if txt[0] != "max":
                    n = min(n, max(1, int(int(txt[0]) / int(txt[1]) + 0.5)))
            else:
                quota = int(txt[0])
                period = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
                if quota > 0 and period > 0:
                    n = min(n, max(1, int(quota / period + 0.5)))
            break
        except (OSError, ValueError, IndexError):
            continue
    return max(1, n)


def time_cpu_baseline(B=65536, F=26, V=1_000_000, E=32, L=3, units=(192, 192), steps=3, warmup=1, optimizer="adamw",
                      threads=None, seed=1234, budget_s=None):
    """examples/s of the CPU restatement on this host.  Returns dict(value, cores, steps, ms_per_step).
    budget_s bounds the wall time: timed steps stop once the budget is spent (at least one is always run)."""
    threads = threads or usable_cores()
    torch.set_num_threads(threads)
    tables, cross, mlp = synthetic_c2(F, V, E, L, units, seed)
    model = TorchDCN(tables, cross, mlp, lr=0.01, optimizer=optimizer)
    g = torch.Generator().manual_seed(seed + 1)
    times = []
    t_start = time.perf_counter()
    for s in range(warmup + steps):
        ids = torch.randint(0, V, (B, F), generator=g)
        y = torch.rand((B,), generator=g)
        t0 = time.perf_counter()
        model.train_step(ids, y)
        dt = time.perf_counter() - t0
        if s >= warmup:
            times.append(dt)
        if budget_s is not None and times and (time.perf_counter() - t_start) + dt > budget_s:
            break
    sec = sum(times) / len(times)
    return dict(value=B / sec, cores=threads, steps=len(times), warmup=min(warmup, s), ms_per_step=sec * 1e3, batch=B)


def time_reference_arm(B=65536, F=26, V=1_000_000, E=32, L=3, units=(192, 192), steps=20, warmup=5, optimizer="adamw",
                       threads=None, seed=1234, budget_s=240.0, mem_cap_bytes=24e9):
    """`bench.py --impl reference`: EXACTLY `warmup` untimed + `steps` timed steps.  Every step is the full workload when
    `warmup + steps` of them fit `budget_s` (and the tables + optimizer state + dense gradients fit `mem_cap_bytes` of host
    memory); otherwise every step is a bounded sample of the same workload: batch AND vocabulary shrunk by the same
    power-of-two factor, which keeps the per-example cost (dense layers per example + the dense optimizer sweep per
    example, V*F*E/B) unchanged.  The factor is found by timing one step at each candidate size (that step counts as the
    first warm-up step of the size finally used)."""
    threads = threads or usable_cores()
    torch.set_num_threads(threads)
    slots = {"adamw": 4, "adagrad": 3, "sgd": 2}.get(optimizer, 4)          # p + grad (+ m, v | acc), fp32

    def build(scale):
        b, v = max(int(B * scale), 64), max(int(V * scale), 64)
        tables, cross, mlp = synthetic_c2(F, v, E, L, units, seed)
        return TorchDCN(tables, cross, mlp, lr=0.01, optimizer=optimizer), b, v

    g = torch.Generator().manual_seed(seed + 1)

    def one(model, b, v):
        ids = torch.randint(0, v, (b, F), generator=g)
        y = torch.rand((b,), generator=g)
        t0 = time.perf_counter()
        model.train_step(ids, y)
        return time.perf_counter() - t0

    scale = 1.0
    while scale > 1.0 / 65536 and F * V * scale * E * 4.0 * slots > mem_cap_bytes:
        scale /= 2
    probes = []
    while True:
        model, b, v = build(scale)
        t1 = one(model, b, v)
        probes.append((scale, t1))
        if t1 * (steps + max(warmup, 1) - 1) <= budget_s or scale <= 1.0 / 65536:
            break
        need = t1 * (steps + warmup) / budget_s
        del model
        while need > 1.0 and scale > 1.0 / 65536:
            scale /= 2
            need /= 2
    for _ in range(max(warmup - 1, 0)):
        one(model, b, v)
    times = [one(model, b, v) for _ in range(steps)]
    sec = sum(times) / len(times)
    sample = (f"every step = the full batch of {B} examples over the full {F} x {V}-row tables" if scale == 1.0 else
              f"every step = a 1/{int(round(1 / scale))} sample: batch {b} over {F} x {v}-row tables (batch and vocabulary shrunk "
              f"together so the per-example cost incl. the dense optimizer sweep is unchanged; probe steps: "
              + ", ".join(f"1/{int(round(1 / sc))}: {t:.2f} s" for sc, t in probes) + ")")
    return dict(value=b / sec, cores=threads, steps=steps, warmup=warmup, ms_per_step=sec * 1e3, batch=b, vocab=v, scale=scale,
                sample=sample)
