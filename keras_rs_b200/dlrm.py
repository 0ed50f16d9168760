"""DLRM — the examples/ml_perf model wiring (examples/ml_perf/model.py:60-212) on the layers of this package.

    dense (B, 13) --bottom MLP (relu ... relu)--> (B, E)  \
    ids   (B, F)  --fused multi-table gather----> (B, F*E) --> interaction --> top MLP (relu ... sigmoid) --> (B, 1)

interaction = "dot"   : classic DLRM / BASELINE.json config C3 — DotInteraction over [bottom, e_1 .. e_F] (F + 1 features
                        of E dims, lower triangle without the diagonal), concatenated behind the bottom output;
interaction = "cross" : ml_perf's DLRM-DCNv2 — DCNBlock (model.py:286-336) of low-rank FeatureCross layers over
                        concat([bottom, e_1 .. e_F]) (model.py:204-208), D = E * (F + 1) = 3456 for the ml_perf shape.
Tables use the `sum` combiner on one-hot ids (main.py:165).  Every layer is the public class of this package; the model
runs through torch.autograd over the same C-ABI kernels, `train_on_batch` adds BCE (main.py:201-210) and the optimizer.
Initial values follow model.py:226-266 (VarianceScaling(1.0, fan_in, uniform) kernels AND biases) and
model.py:317-325 (GlorotUniform cross kernels, zero biases)."""
from __future__ import annotations

from typing import Sequence

import torch

from . import initializers, ops, optimizers
from .layers import Dense, DotInteraction, FeatureCross


class DLRM(torch.nn.Module):
    def __init__(self, vocab_sizes: Sequence[int], embedding_dim: int = 128, num_dense: int = 13,
                 bottom_mlp_dims: Sequence[int] = (512, 256, 128), top_mlp_dims: Sequence[int] = (1024, 1024, 512, 256, 1),
                 interaction: str = "dot", num_dcn_layers: int = 3, dcn_projection_dim: int | None = 512, seed: int = 0,
                 device: str = "cuda"):
        super().__init__()
        if interaction not in ("dot", "cross"):
            raise ValueError(f"`interaction` must be 'dot' or 'cross', got {interaction!r}")
        if int(bottom_mlp_dims[-1]) != int(embedding_dim):
            raise ValueError("the bottom MLP must end in `embedding_dim` units (its output is one of the interacting features)")
        self.vocab_sizes = [int(v) for v in vocab_sizes]
        self.F, self.E, self.num_dense, self.interaction = len(self.vocab_sizes), int(embedding_dim), int(num_dense), interaction
        self.device_ = torch.device(device)
        vs = lambda i: initializers.VarianceScaling(scale=1.0, mode="fan_in", distribution="uniform", seed=seed + i)

        def mlp(dims, final_activation, base):
            acts = ["relu"] * (len(dims) - 1) + [final_activation]
            return torch.nn.ModuleList([Dense(int(u), activation=a, kernel_initializer=vs(base + 2 * i),
                                              bias_initializer=vs(base + 2 * i + 1), device=device, name=f"dense_{base}_{i}")
                                        for i, (u, a) in enumerate(zip(dims, acts))])

        self.bottom_mlp = mlp(bottom_mlp_dims, "relu", 100)                     # model.py:105-112
        self.top_mlp = mlp(top_mlp_dims, "sigmoid", 200)                         # model.py:157-164
        # tables: U(-0.05, 0.05), the keras.layers.Embedding default (the ml_perf configs override it per table)
        self.tables = torch.nn.ParameterList([
            torch.nn.Parameter(initializers.RandomUniform(-0.05, 0.05, seed=seed + 1000 + f)((v, self.E)).to(self.device_))
            for f, v in enumerate(self.vocab_sizes)])
        self.dot = DotInteraction() if interaction == "dot" else None
        self.cross = torch.nn.ModuleList()
        if interaction == "cross":
            self.cross = torch.nn.ModuleList([
                FeatureCross(projection_dim=dcn_projection_dim, kernel_initializer=initializers.GlorotUniform(seed=seed + 300 + i),
                             bias_initializer="zeros", device=device, name=f"cross_{i}") for i in range(int(num_dcn_layers))])
        # build eagerly so that parameters exist before the first call
        k = self.num_dense
        for d in self.bottom_mlp:
            d.build((None, k)); k = d.units
        n_feat = self.F + 1
        k = self.E + n_feat * (n_feat - 1) // 2 if interaction == "dot" else self.E * n_feat
        for c in self.cross:
            c.build((None, k))
        for d in self.top_mlp:
            d.build((None, k)); k = d.units

    # ------------------------------------------------------------------ forward (model.py:175-212)
    def forward(self, dense: torch.Tensor, ids: torch.Tensor, sparse_arena: bool = False) -> torch.Tensor:
        """sparse_arena=True: the tables' gradients go to their persistent arenas + touched bitmaps (row-sparse optimizer
        updates, no (V, E) gradient is materialised) instead of dense `.grad` tensors."""
        if ids.dtype not in (torch.int32, torch.int64):
            ids = ids.to(torch.int32)
        h = dense
        for d in self.bottom_mlp:
            h = d(h)
        emb = ops.gather_concat([dict(table=self.tables[f], ids=ids[:, f], weights=None, combiner="sum") for f in range(self.F)],
                                sparse_arena=sparse_arena)                      # (B, F*E): lookup + concat in one kernel
        if self.interaction == "dot":
            # one (B, (F+1)*E) buffer [bottom | e_1 .. e_F]: DotInteraction reads its features as strided views of it and its
            # backward writes ONE gradient buffer (ops.dot_interaction_packed) — same kernels, same order as DotInteraction()(list)
            packed = torch.cat([h, emb], dim=-1)
            x = torch.cat([h, ops.dot_interaction_packed(packed, self.F + 1, self.E)], dim=-1)
        else:
            x0 = torch.cat([h, emb], dim=-1)                                    # model.py:204-207
            x = x0
            for c in self.cross:                                                # DCNBlock.call, model.py:332-336
                x = c(x0, x)
        for d in self.top_mlp:
            x = d(x)
        return x

    def parameters_list(self):
        ps = list(self.tables)
        for d in list(self.bottom_mlp) + list(self.top_mlp):
            ps += [d.kernel] + ([d.bias] if d.bias is not None else [])
        for c in self.cross:
            ps += [w for w in (c.down_proj_kernel, c.kernel, c.bias) if w is not None]
        return ps

    def train_on_batch(self, dense, ids, labels, optimizer: optimizers.Optimizer) -> torch.Tensor:
        """forward -> BCE (main.py:201-210) -> backward -> optimizer on every variable; returns the loss tensor."""
        params = self.parameters_list()
        optimizer.zero_grad(params)
        pred = self.forward(dense, ids, sparse_arena=True)
        loss = ops.loss_fn(pred, labels, "bce")
        loss.backward()
        optimizer.apply(params)
        return loss.detach()

    def train_on_batch_graph(self, dense, ids, labels, optimizer: optimizers.Optimizer) -> torch.Tensor:
        """The step of train_on_batch replayed from a CUDA graph captured on first use (per batch shape and optimizer).
        The eager step is host-bound here (public layers through autograd: ~150 launches whose Python dispatch takes longer
        than the kernels run); the replay is one launch.  Inputs are copied into the graph's static buffers, the returned
        loss tensor is the graph's (valid until the next replay).  AdamW / Adam read step and bias correction from the
        optimizer's device-resident hyper-parameters; lazy Adam (`sparse_rows=True`) computes them on the host and is refused."""
        if getattr(optimizer, "sparse_rows", False):
            raise ValueError("train_on_batch_graph: lazy Adam (sparse_rows=True) takes its bias correction from the host step counter "
                             "and cannot be replayed; use train_on_batch")
        if not hasattr(self, "_graphs"):
            self._graphs = {}
        key = (tuple(dense.shape), tuple(ids.shape), ids.dtype, id(optimizer))
        st = self._graphs.get(key)
        if st is None:
            optimizer.enable_device_hyper(self.device_)
            s_dense, s_ids, s_y = dense.clone(), ids.clone(), labels.clone()
            # this batch's step runs eagerly (library one-time initialisation, optimizer slots, arenas); the graph captured
            # right after it serves every later batch of this shape
            loss = self.train_on_batch(s_dense, s_ids, s_y, optimizer)
            torch.cuda.synchronize()
            params = self.parameters_list()
            optimizer.zero_grad(params)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                pred = self.forward(s_dense, s_ids, sparse_arena=True)
                g_loss = ops.loss_fn(pred, s_y, "bce")
                g_loss.backward()
                optimizer.apply(params)
            optimizer.iterations -= 1          # capturing does not execute: the host step counter stays where it was
            self._graphs[key] = (g, s_dense, s_ids, s_y, g_loss.detach())
            return loss
        g, s_dense, s_ids, s_y, g_loss = st
        s_dense.copy_(dense, non_blocking=True)
        s_ids.copy_(ids, non_blocking=True)
        s_y.copy_(labels, non_blocking=True)
        g.replay()
        optimizer.iterations += 1
        return g_loss
