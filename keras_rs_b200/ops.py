"""Host-side glue: torch.autograd.Function wrappers whose forward/backward call the C ABI.

No torch.nn / aten math runs on the hot path here — torch provides device memory (torch.empty),
the current stream and the autograd graph only.
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import torch

from . import _lib as L
from ._lib import check, lib, ptr, require_cuda, stream


def _c(t: torch.Tensor) -> torch.Tensor:
    return t if t.is_contiguous() else t.contiguous()


_GEMM_WS: dict[int, torch.Tensor] = {}    # device index -> scratch registered with krs_gemm_set_workspace


def ensure_gemm_workspace(nbytes: int = 32 << 20) -> None:
    """Registers a caller-owned scratch buffer for the tcgen05 engines (precomputed low-order TF32 plane of small B
    operands, include/krs_b200.h krs_gemm_set_workspace) for the CURRENT device, once.  Called lazily where weights are
    created / GEMMs are issued — never at import time (importing the package must not create a CUDA context, and a rank
    that calls torch.cuda.set_device after the import must get its buffer on its own GPU).  The library keeps one
    registration per device; GEMMs that use it must be issued on one stream at a time per device."""
    if not torch.cuda.is_available() or lib.krs_get_gemm_engine() == 0:
        return
    dev = torch.cuda.current_device()
    if dev not in _GEMM_WS:
        _GEMM_WS[dev] = torch.empty(nbytes, dtype=torch.uint8, device=f"cuda:{dev}")
        check(lib.krs_gemm_set_workspace(_GEMM_WS[dev].data_ptr(), _GEMM_WS[dev].numel()))


def set_gemm_engine(name: str) -> None:
    """'ffma' (exact fp32 FMA products), 'tcgen05' (tensor pipe, 3xTF32 split, operands from shared memory) or
    'tcgen05_ts' (same, A operand kept in tensor memory).  Process-wide switch; touches no CUDA state."""
    check(lib.krs_set_gemm_engine({"ffma": 0, "tcgen05": 1, "tcgen05_ts": 2}[name]))
    if torch.cuda.is_available() and torch.cuda.is_initialized():
        ensure_gemm_workspace()            # switching engines at run time: the current device gets its scratch now


def get_gemm_engine() -> str:
    return ["ffma", "tcgen05", "tcgen05_ts"][lib.krs_get_gemm_engine()]


# ----------------------------------------------------------------------------- gather
class GatherPlan:
    """Pre-built krs_feature_t array for one fused multi-table lookup.

    features: list of dicts(table=Tensor(V,E), ids=Tensor (B,) or (B,H), weights=Tensor|None,
    combiner=str).  The output is the concatenation over features along axis 1
    (examples/dcn.py:437)."""

    def __init__(self, features: Sequence[dict]):
        self.F = len(features)
        if self.F == 0:
            raise ValueError("gather: need at least one feature")
        self.arr = (L.KrsFeature * self.F)()
        self.keep = []  # keep tensors alive
        off = 0
        B = None
        for i, f in enumerate(features):
            table = f["table"]
            ids = f["ids"]
            w = f.get("weights")
            require_cuda(table, "embedding table")
            if not table.is_contiguous() or table.dim() != 2:
                raise ValueError("embedding table must be a contiguous (vocab, dim) tensor")
            if not isinstance(ids, torch.Tensor) or not ids.is_cuda:
                raise L.KrsError("ids must be a CUDA tensor (keras_rs_b200 has no CPU path)")
            if ids.dtype not in (torch.int32, torch.int64):
                ids = ids.to(torch.int32)  # Keras Embedding casts non-int ids to int32
            if ids.dim() not in (1, 2):
                raise ValueError(f"ids must be rank 1 or 2, got rank {ids.dim()}")
            if ids.dim() == 2 and ids.stride(1) != 1:
                ids = ids.contiguous()
            b = ids.shape[0]
            if B is None:
                B = b
            elif b != B:
                raise ValueError("all features must share the batch dimension")
            H = 1 if ids.dim() == 1 else ids.shape[1]
            d = self.arr[i]
            d.table = table.data_ptr()
            d.ids = ids.data_ptr()
            d.ids_stride = ids.stride(0) if B > 0 else max(H, 1)
            if w is not None:
                require_cuda(w, "weights")
                if tuple(w.shape) != tuple(ids.shape):
                    raise ValueError(
                        f"The shape of `weights`: {tuple(w.shape)} is not compatible with the shape of "
                        f"`inputs` after embedding: {tuple(ids.shape) + (table.shape[1],)}.")
                if w.stride() != ids.stride():
                    w = w.contiguous()
                    if ids.stride() != w.stride():
                        ids = ids.contiguous()
                        d.ids = ids.data_ptr()
                        d.ids_stride = ids.stride(0) if B > 0 else max(H, 1)
                d.weights = w.data_ptr()
            else:
                d.weights = None
            d.grad = None
            d.touched = None
            d.vocab = table.shape[0]
            d.hotness = H
            d.dim = table.shape[1]
            d.out_offset = off
            d.combiner = L.COMBINER[f.get("combiner", "mean")]
            d.ids_i64 = 1 if ids.dtype == torch.int64 else 0
            d.reduce = 1 if ids.dim() == 2 else 0
            d.shard_tables = None
            d.shard_grads = None
            d.shard_touched = None
            d.num_shards = 1
            off += table.shape[1]
            self.keep.append((table, ids, w))
        self.B = int(B)
        self.out_dim = off
        self.tables = [f["table"] for f in features]

    def forward(self, out: torch.Tensor | None = None, variant: int = 0) -> torch.Tensor:
        if out is None:
            out = torch.empty((self.B, self.out_dim), device=self.tables[0].device, dtype=torch.float32)
        check(lib.krs_gather_fwd(self.arr, self.F, self.B, out.data_ptr(), out.stride(0) if self.B else self.out_dim,
                                 variant, stream()))
        return out

    def backward(self, gout: torch.Tensor, grads: Sequence[torch.Tensor],
                 touched: Sequence[torch.Tensor | None]) -> None:
        """Scatter-add gout (B, out_dim) into per-feature dense arenas `grads[i]` (V,E) (accumulating)."""
        gout = _c(gout)
        for i in range(self.F):
            self.arr[i].grad = grads[i].data_ptr()
            self.arr[i].touched = None if touched[i] is None else touched[i].data_ptr()
        check(lib.krs_gather_bwd(self.arr, self.F, self.B, gout.data_ptr(), gout.stride(0) if self.B else self.out_dim,
                                 stream()))


def ensure_arena(table: torch.Tensor):
    """Persistent zero-initialised (V,E) gradient arena + touched bitmap attached to a table."""
    if getattr(table, "_krs_arena", None) is None:
        table._krs_arena = torch.zeros_like(table)
        table._krs_touched = torch.zeros(((table.shape[0] + 31) // 32,), dtype=torch.int32, device=table.device)
    return table._krs_arena, table._krs_touched


class _GatherFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, plan: GatherPlan, sparse_arena: bool, variant: int, *tables):
        ctx.plan = plan
        ctx.sparse_arena = sparse_arena
        return plan.forward(None, variant)

    @staticmethod
    def backward(ctx, gout):
        plan = ctx.plan
        if ctx.sparse_arena:
            grads, touched = [], []
            for t in plan.tables:
                a, b = ensure_arena(t)
                t._krs_arena_dirty = True
                grads.append(a)
                touched.append(b)
            plan.backward(gout, grads, touched)
            return (None, None, None) + tuple(None for _ in plan.tables)
        # dense-gradient mode (what the reference's non-TPU path produces: SURVEY a1)
        uniq = {}
        grads = []
        for t in plan.tables:
            if id(t) not in uniq:
                uniq[id(t)] = torch.zeros_like(t)
            grads.append(uniq[id(t)])
        plan.backward(gout, grads, [None] * plan.F)
        seen = set()
        out = []
        for t in plan.tables:  # a shared table appears once per feature: return its grad once
            if id(t) in seen:
                out.append(None)
            else:
                seen.add(id(t))
                out.append(uniq[id(t)])
        return (None, None, None) + tuple(out)


def gather_concat(features: Sequence[dict], sparse_arena: bool = False, variant: int = 0) -> torch.Tensor:
    plan = GatherPlan(features)
    return _GatherFn.apply(plan, sparse_arena, variant, *plan.tables)


# ----------------------------------------------------------------------------- FeatureCross
class _CrossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x0, x, U, V, b, diag, act, same_input):
        B, D = x0.shape
        P = 0 if U is None else U.shape[1]
        need_grad = any(t is not None and t.requires_grad for t in (x0, x, U, V, b))
        y = torch.empty_like(x0)
        h2 = torch.empty_like(x0) if need_grad else None
        z = torch.empty_like(x0) if (need_grad and act != 0) else None
        hproj = torch.empty((B, P), device=x0.device, dtype=torch.float32) if U is not None else None
        check(lib.krs_cross_fwd(ptr(x0), ptr(x), ptr(U), ptr(V), ptr(b), float(diag), act, ptr(y), ptr(h2), ptr(z),
                                ptr(hproj), B, D, P, stream()))
        ctx.save_for_backward(x0, x, U, V, b, h2, z, hproj)
        ctx.diag, ctx.act, ctx.same_input = float(diag), act, same_input
        return y

    @staticmethod
    def backward(ctx, gy):
        x0, x, U, V, b, h2, z, hproj = ctx.saved_tensors
        gy = _c(gy)
        B, D = x0.shape
        P = 0 if U is None else U.shape[1]
        dx0 = torch.empty_like(x0)
        dx = torch.empty_like(x0)
        dV = torch.empty_like(V)
        dU = torch.empty_like(U) if U is not None else None
        db = torch.empty_like(b) if b is not None else None
        dz = torch.empty_like(x0)
        dh = torch.empty((B, P), device=x0.device, dtype=torch.float32) if U is not None else None
        flags = L.CROSS_SAME_INPUT if ctx.same_input else 0
        check(lib.krs_cross_bwd(ptr(gy), ptr(x0), ptr(x), ptr(U), ptr(V), ptr(h2), ptr(z), ptr(hproj), ctx.diag,
                                ctx.act, ptr(dx0), ptr(dx), ptr(dU), ptr(dV), ptr(db), ptr(dz), ptr(dh), B, D, P,
                                flags, stream()))
        if ctx.same_input:
            return dx, None, dU, dV, db, None, None, None
        return dx0, dx, dU, dV, db, None, None, None


class _CrossCombineFn(torch.autograd.Function):
    """y = x0 * (a + diag*x) + x for a user-supplied (callable) pre_activation output `a`."""

    @staticmethod
    def forward(ctx, x0, x, a, diag):
        y = torch.empty_like(x0)
        check(lib.krs_cross_combine_fwd(ptr(x0), ptr(x), ptr(a), float(diag), ptr(y), x0.numel(), stream()))
        ctx.save_for_backward(x0, x, a)
        ctx.diag = float(diag)
        return y

    @staticmethod
    def backward(ctx, gy):
        x0, x, a = ctx.saved_tensors
        gy = _c(gy)
        dx0, dx, da = torch.empty_like(x0), torch.empty_like(x0), torch.empty_like(x0)
        check(lib.krs_cross_combine_bwd(ptr(gy), ptr(x0), ptr(x), ptr(a), ctx.diag, ptr(dx0), ptr(dx), ptr(da),
                                        x0.numel(), stream()))
        return dx0, dx, da, None


def feature_cross(x0, x, U, V, b, diag_scale, act: int, same_input: bool):
    ensure_gemm_workspace()
    return _CrossFn.apply(x0, x, U, V, b, diag_scale or 0.0, act, same_input)


def cross_combine(x0, x, a, diag_scale):
    return _CrossCombineFn.apply(x0, x, a, diag_scale or 0.0)


# ----------------------------------------------------------------------------- Dense
class _DenseFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, W, b, act):
        B, K = x.shape
        N = W.shape[1]
        ctx.K = K
        if K % 4 != 0 and K > 64 and lib.krs_get_gemm_engine() != 0:
            # TMA needs 16-byte row strides: a width like DLRM's 128 + 351 = 479 would fall back to the FFMA engine (measured
            # 2.13 ms vs 0.31 ms for the forward at B = 65536, N = 1024).  Zero-pad x and W to the next multiple of 4; the
            # padded column / row contribute exact zeros, and the caller still sees (B, K) / (K, N) shapes and gradients.
            K4 = (K + 3) // 4 * 4
            xp = torch.zeros((B, K4), device=x.device, dtype=torch.float32)
            xp[:, :K] = x
            Wp = torch.zeros((K4, N), device=W.device, dtype=torch.float32)
            Wp[:K] = W
            x, W, K = xp, Wp, K4
        y = torch.empty((B, N), device=x.device, dtype=torch.float32)
        check(lib.krs_dense_fwd(ptr(x), ptr(W), ptr(b), act, ptr(y), B, K, N, stream()))
        ctx.save_for_backward(x, W, b, y)
        ctx.act = act
        return y

    @staticmethod
    def backward(ctx, gy):
        x, W, b, y = ctx.saved_tensors
        gy = _c(gy)
        B, K = x.shape
        N = W.shape[1]
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        dW = torch.empty_like(W)
        db = torch.empty_like(b) if b is not None else None
        dz = torch.empty_like(gy)
        check(lib.krs_dense_bwd(ptr(gy), ptr(x), ptr(W), ptr(y), ctx.act, ptr(dx), ptr(dW), ptr(db), ptr(dz), B, K, N,
                                stream()))
        if K != ctx.K:                                      # padded width: hand back the caller's shapes
            dx = dx[:, :ctx.K] if dx is not None else None
            dW = dW[:ctx.K]
        return dx, dW, db, None


def dense(x, W, b, act: int):
    ensure_gemm_workspace()
    return _DenseFn.apply(x, W, b, act)


def linear_no_bias(x, W):
    """x @ W through the C ABI GEMM (used for callable pre_activations and tests)."""
    return _DenseFn.apply(x, W, None, 0)


def sgemm(A: torch.Tensor, Bm: torch.Tensor, transA=False, transB=False, out=None, accumulate=False):
    M = A.shape[1] if transA else A.shape[0]
    K = A.shape[0] if transA else A.shape[1]
    N = Bm.shape[0] if transB else Bm.shape[1]
    if out is None:
        out = torch.empty((M, N), device=A.device, dtype=torch.float32)
    check(lib.krs_sgemm(ptr(A), ptr(Bm), ptr(out), M, N, K, int(transA), int(transB), int(accumulate), stream()))
    return out


# ----------------------------------------------------------------------------- DotInteraction
def _dot_args(tensors):
    n = len(tensors)
    ptrs = (C.c_void_p * n)(*[t.data_ptr() for t in tensors])
    strides = (C.c_int64 * n)(*[t.stride(0) if t.shape[0] > 1 else t.shape[1] for t in tensors])
    return ptrs, strides


def _as_rows(t: torch.Tensor) -> torch.Tensor:
    # rows must be unit-stride along the feature dim; row stride is free (views of a concat buffer)
    return t if t.stride(1) == 1 else t.contiguous()


class _DotFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, self_interaction, skip_gather, *inputs):
        inputs = [_as_rows(t) for t in inputs]
        n = len(inputs)
        B, E = inputs[0].shape
        out_dim = n * n if skip_gather else (n * (n + 1) // 2 if self_interaction else n * (n - 1) // 2)
        out = torch.empty((B, out_dim), device=inputs[0].device, dtype=torch.float32)
        ptrs, strides = _dot_args(inputs)
        check(lib.krs_dot_fwd(ptrs, strides, n, E, B, int(self_interaction), int(skip_gather), ptr(out), stream()))
        ctx.save_for_backward(*inputs)
        ctx.cfg = (self_interaction, skip_gather)
        return out

    @staticmethod
    def backward(ctx, gout):
        inputs = ctx.saved_tensors
        self_interaction, skip_gather = ctx.cfg
        gout = _c(gout)
        n = len(inputs)
        B, E = inputs[0].shape
        buf = torch.empty((B, n * E), device=gout.device, dtype=torch.float32)
        grads = [buf[:, i * E:(i + 1) * E] for i in range(n)]
        ptrs, strides = _dot_args(inputs)
        gptrs, gstrides = _dot_args(grads)
        check(lib.krs_dot_bwd(ptrs, strides, ptr(gout), gptrs, gstrides, n, E, B, int(self_interaction),
                              int(skip_gather), stream()))
        return (None, None) + tuple(grads)


def dot_interaction(inputs, self_interaction: bool, skip_gather: bool):
    return _DotFn.apply(self_interaction, skip_gather, *inputs)


class _DotPackedFn(torch.autograd.Function):
    """DotInteraction over the n features of ONE (B, n*E) buffer (feature j = columns j*E .. (j+1)*E): the same kernels
    with pointer + row stride per feature, but a single autograd input — the backward writes one (B, n*E) gradient
    instead of n tensors that autograd would each embed into a zero-filled full-size buffer (27 x 0.9 GB at C3)."""

    @staticmethod
    def forward(ctx, x, n, E, self_interaction, skip_gather):
        x = _c(x)
        B = x.shape[0]
        out_dim = n * n if skip_gather else (n * (n + 1) // 2 if self_interaction else n * (n - 1) // 2)
        out = torch.empty((B, out_dim), device=x.device, dtype=torch.float32)
        ptrs = (C.c_void_p * n)(*[x.data_ptr() + 4 * j * E for j in range(n)])
        strides = (C.c_int64 * n)(*[n * E] * n)
        check(lib.krs_dot_fwd(ptrs, strides, n, E, B, int(self_interaction), int(skip_gather), ptr(out), stream()))
        ctx.save_for_backward(x)
        ctx.cfg = (n, E, self_interaction, skip_gather)
        return out

    @staticmethod
    def backward(ctx, gout):
        (x,) = ctx.saved_tensors
        n, E, self_interaction, skip_gather = ctx.cfg
        gout = _c(gout)
        B = x.shape[0]
        dx = torch.empty_like(x)
        ptrs = (C.c_void_p * n)(*[x.data_ptr() + 4 * j * E for j in range(n)])
        gptrs = (C.c_void_p * n)(*[dx.data_ptr() + 4 * j * E for j in range(n)])
        strides = (C.c_int64 * n)(*[n * E] * n)
        check(lib.krs_dot_bwd(ptrs, strides, ptr(gout), gptrs, strides, n, E, B, int(self_interaction), int(skip_gather), stream()))
        return dx, None, None, None, None


def dot_interaction_packed(x: torch.Tensor, n: int, E: int, self_interaction: bool = False, skip_gather: bool = False):
    if x.dim() != 2 or x.shape[1] != n * E:
        raise ValueError(f"dot_interaction_packed: expected a (B, {n * E}) buffer, got {tuple(x.shape)}")
    return _DotPackedFn.apply(x, n, E, self_interaction, skip_gather)


# ----------------------------------------------------------------------------- retrieval
def set_topk_engine(name: str) -> None:
    """'auto' (tensor pipe for large problems), 'ffma' (exact-fp32 FMA score tiles) or 'tcgen05' (whenever eligible)."""
    check(lib.krs_set_topk_engine({"auto": 0, "ffma": 1, "tcgen05": 2}[name]))


def top_k_scores(q: torch.Tensor, cand: torch.Tensor, cand_ids: torch.Tensor | None, k: int):
    """Streaming Q @ C^T + exact top-k (never materialises the score matrix)."""
    q = _c(q)
    cand = _c(cand)
    nq, d = q.shape
    nc = cand.shape[0]
    ws_bytes = lib.krs_topk_workspace_bytes(nq, nc, d, k)
    ws = torch.empty((max(ws_bytes, 16),), dtype=torch.uint8, device=q.device)
    top_s = torch.empty((nq, k), dtype=torch.float32, device=q.device)
    top_i = torch.empty((nq, k), dtype=torch.int32, device=q.device)
    check(lib.krs_topk(ptr(q), ptr(cand), ptr(cand_ids), ptr(top_s), ptr(top_i), nq, nc, d, k, ptr(ws), ws.numel(),
                       stream()))
    return top_s, top_i


# ----------------------------------------------------------------------------- loss
LOSS_KIND = {"mse": 0, "mean_squared_error": 0, "bce": 1, "binary_crossentropy": 1, "bce_logits": 2}


class _LossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, label, kind):
        p = _c(pred).reshape(-1)
        y = _c(label).reshape(-1)
        loss = torch.empty((1,), device=pred.device, dtype=torch.float32)
        dpred = torch.empty_like(p)
        check(lib.krs_loss_fwd_bwd(ptr(p), ptr(y), ptr(loss), ptr(dpred), p.numel(), kind, 0, stream()))
        ctx.save_for_backward(dpred)
        ctx.shape = pred.shape
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        (dpred,) = ctx.saved_tensors
        # g is the scalar upstream gradient (1.0 for loss.backward()); scaling it in is a 1-element op
        return (dpred * g).reshape(ctx.shape), None, None


def loss_fn(pred, label, kind: str = "mse"):
    return _LossFn.apply(pred, label, LOSS_KIND[kind])
