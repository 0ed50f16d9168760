"""DCN-v2 ranking model — the caller of the hot path, wired as examples/dcn.py:418-449 builds it
(per-feature Embedding -> concatenate -> FeatureCross -> Dense(relu)... -> Dense(1)) with the
stacked cross of README.md:54-55 / examples/ml_perf/model.py:332-336 (`xl = layer(x0, xl)`).

Two equivalent execution paths over the SAME weights:
  * `forward(ids)` / autograd: composes the public layers (FeatureCross, Dense, fused gather) — the
    drop-in API path;
  * `train_on_batch(ids, labels)`: the whole step (gather, L cross layers, MLP, loss, every backward
    contraction, embedding scatter-add, optimizer sweeps) as a fixed sequence of C-ABI launches on
    preallocated buffers — no autograd engine, no aten math, no allocation in the loop.
Both are checked against each other and against the CPU oracle in tests/.

All embedding tables live in ONE (sum V_padded, E) arena (each table padded to a multiple of 32 rows
so its slice of the touched bitmap is word aligned); all dense weights live in one flat buffer, so
an optimizer step is two launches.
"""
from __future__ import annotations

from typing import Sequence

import torch

from . import _lib as L
from . import initializers, ops, optimizers
from ._lib import check, lib, ptr, stream
from .layers import Dense, FeatureCross


def _round_up(a: int, b: int) -> int:
    return (a + b - 1) // b * b


class DCN(torch.nn.Module):
    def __init__(self, vocab_sizes: Sequence[int], embedding_dim: int = 32, num_cross_layers: int = 3,
                 projection_dim: int | None = None, dense_units: Sequence[int] = (192, 192),
                 diag_scale: float = 0.0, pre_activation=None, loss: str = "mse", seed: int = 0,
                 device: str = "cuda", embeddings_initializer="uniform", dense_activation: str = "relu"):
        super().__init__()
        self.vocab_sizes = [int(v) for v in vocab_sizes]
        self.F = len(self.vocab_sizes)
        self.E = int(embedding_dim)
        self.D = self.F * self.E
        self.L = int(num_cross_layers)
        self.P = projection_dim
        self.loss_kind = ops.LOSS_KIND[loss]
        self.device_ = torch.device(device)
        self.world = 1
        self.rank = 0
        self._init_tables(seed, embeddings_initializer)
        # ---- cross + MLP layers (public layer classes) --------------------------
        self.cross = torch.nn.ModuleList([
            FeatureCross(projection_dim=projection_dim, diag_scale=diag_scale, pre_activation=pre_activation,
                         kernel_initializer=initializers.GlorotUniform(seed=seed + 1 + i), device=device,
                         name=f"cross_{i}")
            for i in range(self.L)])
        if dense_activation not in (None, "linear", "relu", "sigmoid", "tanh"):
            raise ValueError(f"DCN's fused step supports dense_activation in linear / relu / sigmoid / tanh, got {dense_activation!r}")
        units = list(dense_units) + [1]
        acts = [dense_activation] * len(dense_units) + [None]      # examples/dcn.py:445 uses relu
        self.mlp = torch.nn.ModuleList([
            Dense(u, activation=a, kernel_initializer=initializers.GlorotUniform(seed=seed + 101 + i), device=device,
                  name=f"dense_{i}")
            for i, (u, a) in enumerate(zip(units, acts))])
        for c in self.cross:
            c.build((None, self.D))
        k = self.D
        for d in self.mlp:
            d.build((None, k))
            k = d.units
        self._flatten_dense()
        ops.ensure_gemm_workspace()            # per-device scratch of the tcgen05 engines, registered on first use
        self._bufs = {}
        self._plan = None
        self._plan_key = None

    def _init_tables(self, seed, embeddings_initializer):
        self.row_off = []
        off = 0
        for v in self.vocab_sizes:
            self.row_off.append(off)
            off += _round_up(v, 32)
        self.total_rows = off
        g = torch.Generator(device="cpu").manual_seed(seed)
        init = initializers.get(embeddings_initializer)
        emb = torch.zeros((self.total_rows, self.E), dtype=torch.float32)
        for f, v in enumerate(self.vocab_sizes):
            if isinstance(init, initializers.RandomUniform):
                emb[self.row_off[f]:self.row_off[f] + v] = (
                    torch.rand((v, self.E), generator=g) * (init.maxval - init.minval) + init.minval)
            else:
                emb[self.row_off[f]:self.row_off[f] + v] = init((v, self.E))
        self.emb = torch.nn.Parameter(emb.to(self.device_))
        self.emb_grad = torch.zeros_like(self.emb)                      # gradient arena (persistently zero)
        self.emb_touched = torch.zeros((self.total_rows // 32,), dtype=torch.int32, device=self.device_)
        self.emb._krs_arena = self.emb_grad
        self.emb._krs_touched = self.emb_touched
        # rows that have ever received a gradient (AdamW sweeps the others with a decay-only update, krs_adamw_cold)
        self.emb_ever = torch.zeros_like(self.emb_touched)
        self.emb._krs_ever = self.emb_ever

    # ------------------------------------------------------------------ parameters
    def tables(self):
        return [self.emb[self.row_off[f]:self.row_off[f] + v] for f, v in enumerate(self.vocab_sizes)]

    def dense_params(self):
        out = []
        for c in self.cross:
            out.extend(c.weights)
        for d in self.mlp:
            out.extend(d.weights)
        return out

    def _flatten_dense(self):
        """Re-home every dense weight as a view of one flat buffer (+ a flat gradient buffer)."""
        ps = self.dense_params()
        offs, n = [], 0
        for p in ps:
            offs.append(n)
            n += _round_up(p.numel(), 4)          # keep every view 16-byte aligned
        flat = torch.zeros((n,), dtype=torch.float32, device=self.device_)
        gflat = torch.zeros_like(flat)
        self._dense_views, self._dense_grad_views = [], []
        for p, o in zip(ps, offs):
            v = flat[o:o + p.numel()].view(p.shape)
            v.copy_(p.data)
            p.data = v
            self._dense_views.append(v)
            self._dense_grad_views.append(gflat[o:o + p.numel()].view(p.shape))
        self.dense_flat = flat
        self.dense_grad_flat = gflat
        self._dense_index = {id(p): i for i, p in enumerate(ps)}

    def _g(self, p):
        return self._dense_grad_views[self._dense_index[id(p)]]

    # ------------------------------------------------------------------ inputs
    def _feature_list(self, ids: torch.Tensor):
        tabs = self.tables()
        return [dict(table=tabs[f], ids=ids[:, f], weights=None, combiner="sum") for f in range(self.F)]

    def _as_ids(self, ids):
        if isinstance(ids, dict):   # examples/dcn.py feeds a dict of per-feature (B,) tensors
            ids = torch.stack([ids[k] for k in ids], dim=1)
        if ids.dtype not in (torch.int32, torch.int64):
            ids = ids.to(torch.int32)
        return ids

    # ------------------------------------------------------------------ layer-API path
    def forward(self, ids, sparse_arena: bool = False) -> torch.Tensor:
        ids = self._as_ids(ids)
        feats = self._feature_list(ids)
        if sparse_arena:
            plan = ops.GatherPlan(feats)
            x0 = _ArenaGather.apply(plan, self, self.emb)
        else:
            x0 = ops.gather_concat(feats, sparse_arena=False)
        xl = x0
        for i, c in enumerate(self.cross):
            xl = c(x0) if i == 0 else c(x0, xl)
        h = xl
        for d in self.mlp:
            h = d(h)
        return h

    # ------------------------------------------------------------------ fused training step
    def _step_buffers(self, B: int):
        b = self._bufs.get(B)
        if b is not None:
            return b
        dev, f32 = self.device_, torch.float32
        D, L_ = self.D, self.L
        mk = lambda *s: torch.empty(s, device=dev, dtype=f32)
        b = dict(
            ids=torch.empty((B, self.F), device=dev, dtype=torch.int32),
            labels=mk(B), xs=[mk(B, D) for _ in range(L_ + 1)], h2=[mk(B, D) for _ in range(L_)],
            z=[mk(B, D) if self.cross[i]._act_id != 0 else None for i in range(L_)],
            hproj=[mk(B, self.P) if self.P is not None else None for _ in range(L_)],
            hs=[mk(B, d.units) for d in self.mlp], loss=mk(1), dpred=mk(B),
            ga=mk(B, D), gb=mk(B, D), dx0=mk(B, D), dz=mk(B, D),
            dh=mk(B, self.P) if self.P is not None else None,
            mg=[mk(B, d.units) for d in self.mlp], mdz=[mk(B, d.units) for d in self.mlp],
        )
        self._bufs[B] = b
        b["plan"] = self._make_plan(b["ids"])
        return b

    def _make_plan(self, ids: torch.Tensor):
        """Fused-gather descriptor table for a static ids buffer, with the gradient arena wired in."""
        plan = ops.GatherPlan(self._feature_list(ids))
        for f in range(self.F):
            plan.arr[f].grad = self.emb_grad[self.row_off[f]:].data_ptr()
            plan.arr[f].touched = self.emb_touched[self.row_off[f] // 32:].data_ptr()
        return plan

    def forward_backward(self, ids: torch.Tensor, labels: torch.Tensor, denom: int = 0):
        """Forward + backward of one batch through the C ABI only.  ids (B,F) int32 and labels (B,) may
        be pinned-host or device tensors; they are copied into the static device buffers (this IS the
        per-step host->device transfer).  Gradients land in emb_grad/emb_touched and dense_grad_flat."""
        B = ids.shape[0]
        b = self._step_buffers(B)
        b["ids"].copy_(ids, non_blocking=True)
        b["labels"].copy_(labels.reshape(-1), non_blocking=True)
        return self._run_step(b, B, denom)

    def _run_step(self, b, B: int, denom: int = 0):
        """Forward + backward on the staged batch (ids / labels already in the static buffers)."""
        s = stream()
        self._gather_into(b, B, s)
        cur = self._dense_step(b, B, denom, s)
        # cur holds dL/dx0 (B, D): scatter-add into the embedding arena
        self._scatter_from(b, B, cur, s)
        return b["loss"]

    def _dense_step(self, b, B: int, denom: int, s):
        """Everything between the gather and the scatter: cross stack, MLP, loss and their backward.  Returns the
        buffer (ga or gb) that holds dL/dx0."""
        D, P = self.D, (self.P or 0)
        xs = b["xs"]
        for i, c in enumerate(self.cross):
            x_in = xs[0] if i == 0 else xs[i]
            check(lib.krs_cross_fwd(ptr(xs[0]), ptr(x_in), ptr(c.down_proj_kernel), ptr(c.kernel), ptr(c.bias),
                                    float(c.diag_scale or 0.0), c._act_id, ptr(xs[i + 1]), ptr(b["h2"][i]),
                                    ptr(b["z"][i]), ptr(b["hproj"][i]), B, D, P, s))
        h = xs[self.L]
        for i, d in enumerate(self.mlp):
            check(lib.krs_dense_fwd(ptr(h), ptr(d.kernel), ptr(d.bias), d._act_id, ptr(b["hs"][i]), B, h.shape[1],
                                    d.units, s))
            h = b["hs"][i]
        check(lib.krs_loss_fwd_bwd(ptr(h), ptr(b["labels"]), ptr(b["loss"]), ptr(b["dpred"]), B, self.loss_kind,
                                   int(denom), s))
        # ---- backward -----------------------------------------------------------
        g = b["dpred"]
        for i in range(len(self.mlp) - 1, -1, -1):
            d = self.mlp[i]
            x_in = xs[self.L] if i == 0 else b["hs"][i - 1]
            dx = b["ga"] if i == 0 else b["mg"][i - 1]
            check(lib.krs_dense_bwd(ptr(g), ptr(x_in), ptr(d.kernel), ptr(b["hs"][i]), d._act_id, ptr(dx),
                                    ptr(self._g(d.kernel)), ptr(self._g(d.bias)) if d.bias is not None else None,
                                    ptr(b["mdz"][i]), B, x_in.shape[1], d.units, s))
            g = dx
        # g == ga holds dL/dx_L.  Cross layers top-down; dx0 accumulates across layers.
        cur, nxt = b["ga"], b["gb"]
        for i in range(self.L - 1, -1, -1):
            c = self.cross[i]
            flags = 0
            if i < self.L - 1:
                flags |= L.CROSS_ACC_DX0
            if i == 0:
                flags |= L.CROSS_SAME_INPUT      # x is x0: dx <- total dL/dx0 (incl. accumulated dx0)
            x_in = xs[0] if i == 0 else xs[i]
            check(lib.krs_cross_bwd(ptr(cur), ptr(xs[0]), ptr(x_in), ptr(c.down_proj_kernel), ptr(c.kernel),
                                    ptr(b["h2"][i]), ptr(b["z"][i]), ptr(b["hproj"][i]), float(c.diag_scale or 0.0),
                                    c._act_id, ptr(b["dx0"]), ptr(nxt),
                                    ptr(self._g(c.down_proj_kernel)) if c.down_proj_kernel is not None else None,
                                    ptr(self._g(c.kernel)), ptr(self._g(c.bias)) if c.bias is not None else None,
                                    ptr(b["dz"]), ptr(b["dh"]), B, D, P, flags, s))
            cur, nxt = nxt, cur
        return cur

    def _gather_into(self, b, B, s):
        """Fused multi-table gather of the staged ids into xs[0] (overridden by the row-sharded model)."""
        plan = b["plan"]
        check(lib.krs_gather_fwd(plan.arr, plan.F, B, ptr(b["xs"][0]), self.D, 0, s))

    def _scatter_from(self, b, B, cur, s):
        plan = b["plan"]
        check(lib.krs_gather_bwd(plan.arr, plan.F, B, ptr(cur), self.D, s))

    def train_on_batch(self, ids, labels, optimizer: optimizers.Optimizer, denom: int = 0):
        loss = self.forward_backward(ids, labels, denom)
        optimizer.iterations += 1
        if getattr(optimizer, "_hyper_dev", None) is not None:
            optimizer.advance_device_hyper()       # keep the device-resident step / alpha in lock-step
        with torch.no_grad():
            self._update_tables(optimizer)
            self._sync_gradients()
            optimizer._update(self.dense_flat, self.dense_grad_flat, None)
        self._end_of_step()
        return loss

    def _update_tables(self, optimizer):
        optimizer._update(self.emb, self.emb_grad, self.emb_touched)

    def train_on_batch_graph(self, ids, labels, optimizer: optimizers.Optimizer, denom: int = 0):
        """Same step as train_on_batch, replayed from a CUDA graph captured on first use (per batch size):
        one graph launch instead of ~45 kernel launches + tensor-map encodes, which removes the host-side gaps
        between kernels.  Only the H2D copies of ids / labels stay outside the graph."""
        B = ids.shape[0]
        b = self._step_buffers(B)
        key = ("graph", id(optimizer), int(denom))
        if key not in b:
            optimizer.enable_device_hyper(self.device_)
            # this batch's step runs eagerly (it also performs the library's lazy one-time initialisation);
            # the graph captured right after it serves every later batch of this size
            loss = self.train_on_batch(ids, labels, optimizer, denom)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._run_step(b, B, denom)
                optimizer.advance_device_hyper()
                with torch.no_grad():
                    self._update_tables(optimizer)
                    self._sync_gradients()
                    optimizer._update(self.dense_flat, self.dense_grad_flat, None)
                self._end_of_step()
            b[key] = g                         # capturing does not execute: host and device step counters unchanged
            return loss
        b["ids"].copy_(ids, non_blocking=True)
        b["labels"].copy_(labels.reshape(-1), non_blocking=True)
        b[key].replay()
        optimizer.iterations += 1
        return b["loss"]

    def _end_of_step(self):
        """Hook for cross-rank ordering at the end of a step (row-sharded model)."""

    def _sync_gradients(self):
        """Single GPU: nothing to exchange."""

    def predict(self, ids: torch.Tensor) -> torch.Tensor:
        with torch.no_grad():
            return self.forward(ids)


class _ArenaGather(torch.autograd.Function):
    """Layer-API gather whose backward writes the model's gradient arena (no dense (V,E) grad)."""

    @staticmethod
    def forward(ctx, plan, model, emb):
        ctx.plan, ctx.model = plan, model
        return plan.forward()

    @staticmethod
    def backward(ctx, gout):
        m, plan = ctx.model, ctx.plan
        grads = [m.emb_grad[m.row_off[f]:] for f in range(m.F)]
        touched = [m.emb_touched[m.row_off[f] // 32:] for f in range(m.F)]
        plan.backward(gout, grads, touched)
        m.emb._krs_arena_dirty = True
        return None, None, None
