"""Optimizers with the Keras 3 update rules the examples use (AdamW: examples/dcn.py:127; Adagrad:
examples/ml_perf/main.py:203; SGD), applied by fused CUDA sweeps (csrc/optim.cu).

A parameter that owns a gradient arena (`_krs_arena` + `_krs_touched`, written by the fused
embedding backward) is updated straight from the arena: AdamW still visits every row (that is what
the reference's dense update does) but only reads gradient rows that were touched; SGD / Adagrad
visit touched rows only — identical to the dense update because their step is zero where g = 0.
Otherwise the dense `.grad` is used."""
from __future__ import annotations

import ctypes as C
from typing import Iterable

import torch

from ._lib import check, lib, ptr, stream


class Optimizer:
    def __init__(self, learning_rate: float):
        self.learning_rate = float(learning_rate)
        self.iterations = 0
        self._state: dict[int, dict] = {}

    def _slots(self, p: torch.Tensor, names: Iterable[str], init: float = 0.0):
        st = self._state.get(id(p))
        if st is None:
            st = {n: torch.full_like(p, init) if init else torch.zeros_like(p) for n in names}
            self._state[id(p)] = st
        return st

    def _grad_of(self, p):
        """The fused embedding backward writes a parameter's gradient either into its arena (sparse_arena=True /
        train_on_batch: p.grad stays None) or, on the plain autograd path, into p.grad.  A dense p.grad wins; if a shared
        table received both kinds in one step the arena rows are folded into the dense gradient (off the hot path)."""
        arena = getattr(p, "_krs_arena", None)
        if p.grad is None:
            if arena is None:
                return None, None
            p._krs_arena_dirty = False           # the sweep consumes (and re-zeroes) the arena
            return arena, p._krs_touched
        g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
        if arena is not None and getattr(p, "_krs_arena_dirty", False):
            g = g + arena
            arena.zero_()
            p._krs_touched.zero_()
            p._krs_arena_dirty = False
        return g, None

    def apply(self, params: Iterable[torch.Tensor]) -> None:
        self.iterations += 1
        if getattr(self, "_hyper_dev", None) is not None:
            self.advance_device_hyper()
        with torch.no_grad():
            for p in params:
                g, touched = self._grad_of(p)
                if g is None:
                    continue
                self._update(p, g, touched)

    step = apply

    def enable_device_hyper(self, device="cuda"):
        return None

    def advance_device_hyper(self):
        pass

    def zero_grad(self, params: Iterable[torch.Tensor]) -> None:
        for p in params:
            p.grad = None

    def _update(self, p, g, touched):
        raise NotImplementedError

    # ---- compact gradient rows (row-sharded tables, keras_rs_b200/sharded.py) ----
    def _rows_apply(self, p, cs, kind, hyper, s1=None, s2=None):
        h = (C.c_float * 8)(*([float(x) for x in hyper] + [0.0] * (8 - len(hyper))))
        check(lib.krs_rows_apply(ptr(p), ptr(s1), ptr(s2), ptr(cs.compact), ptr(cs.uniq_rows), ptr(cs.n_unique),
                                 cs.cap_rows, p.shape[-1], kind, h, ptr(cs.touched), cs.touched.numel(), stream()))

    def _opt_apply(self, p, g, touched, kind, hyper, s1=None, s2=None):
        """krs_opt_apply: one of the four row rules from an arena (touched rows only) or a dense gradient."""
        h = (C.c_float * 8)(*([float(x) for x in hyper] + [0.0] * (8 - len(hyper))))
        row_len = p.shape[-1] if (touched is not None and p.dim() >= 2) else 1
        check(lib.krs_opt_apply(ptr(p), ptr(s1), ptr(s2), ptr(g), ptr(touched), p.numel(), row_len, kind, h, stream()))

    def _update_compact(self, p, cs):
        """p: (rows, E) table shard; cs: CompactGrads (one gradient row per distinct touched row, in row order)."""
        raise NotImplementedError(f"{type(self).__name__} has no compact-row update")


OPT_SGD, OPT_ADAGRAD, OPT_ADAM, OPT_FTRL = 0, 1, 2, 3


class AdamW(Optimizer):
    def __init__(self, learning_rate=0.001, weight_decay=0.004, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
        super().__init__(learning_rate)
        self.weight_decay, self.beta_1, self.beta_2, self.epsilon = weight_decay, beta_1, beta_2, epsilon

    def _update(self, p, g, touched):
        st = self._slots(p, ("m", "v"))
        if getattr(self, "sparse_rows", False) and touched is not None:   # lazy Adam from the arena: touched rows only
            self._opt_apply(p, g, touched, OPT_ADAM, [self.learning_rate, self.beta_1, self.beta_2, self.epsilon, self._alpha()],
                            st["m"], st["v"])
            return
        row_len = p.shape[-1] if (touched is not None and p.dim() >= 2) else 1
        ever = getattr(p, "_krs_ever", None) if touched is not None else None
        if ever is not None and not getattr(p, "_krs_ever_owner", None) in (None, id(self)):
            ever = None                     # the bitmap describes the moments of ONE optimizer instance
        if ever is not None:
            # rows that never received a gradient hold m = v = 0: decay-only update, 8 instead of 24 bytes per parameter
            p._krs_ever_owner = id(self)
            check(lib.krs_adamw_cold(ptr(p), ptr(st["m"]), ptr(st["v"]), ptr(g), ptr(touched), ptr(ever), p.numel(), row_len,
                                     self.learning_rate, self.beta_1, self.beta_2, self.epsilon, self.weight_decay,
                                     max(self.iterations, 1), ptr(getattr(self, "_hyper_dev", None)), stream()))
            return
        check(lib.krs_adamw(ptr(p), ptr(st["m"]), ptr(st["v"]), ptr(g), ptr(touched), p.numel(), row_len,
                            self.learning_rate, self.beta_1, self.beta_2, self.epsilon, self.weight_decay,
                            max(self.iterations, 1), ptr(getattr(self, "_hyper_dev", None)), stream()))

    def _alpha(self):
        t = max(self.iterations, 1)
        return self.learning_rate * (1.0 - self.beta_2 ** t) ** 0.5 / (1.0 - self.beta_1 ** t)

    def _update_compact(self, p, cs):
        st = self._slots(p, ("m", "v"))
        if getattr(self, "sparse_rows", False):      # lazy Adam: only rows with gradient (the SparseCore form)
            self._rows_apply(p, cs, OPT_ADAM, [self.learning_rate, self.beta_1, self.beta_2, self.epsilon, self._alpha()],
                             st["m"], st["v"])
            return
        ever = getattr(p, "_krs_ever", None)
        if ever is not None and not getattr(p, "_krs_ever_owner", None) in (None, id(self)):
            ever = None                     # the bitmap describes the moments of ONE optimizer instance
        if ever is not None:
            p._krs_ever_owner = id(self)
        check(lib.krs_adamw_compact(ptr(p), ptr(st["m"]), ptr(st["v"]), ptr(cs.compact), ptr(cs.touched), ptr(cs.wordprefix),
                                    ptr(cs.blockbase), ptr(ever), p.numel(), p.shape[-1], self.learning_rate, self.beta_1,
                                    self.beta_2, self.epsilon, self.weight_decay, max(self.iterations, 1), stream()))

    # ---- device-resident hyper-parameters (CUDA-graph replay of the step) ----
    def enable_device_hyper(self, device="cuda"):
        """Keep [lr, b1, b2, eps, wd, alpha, step] on the device; `advance_device_hyper()` (a 1-thread kernel)
        then replaces the host-side step counter, so a captured step needs no per-replay parameters."""
        if getattr(self, "_hyper_dev", None) is None:
            self._hyper_dev = torch.tensor([self.learning_rate, self.beta_1, self.beta_2, self.epsilon,
                                            self.weight_decay, 0.0, float(self.iterations)], dtype=torch.float32,
                                           device=device)
        return self._hyper_dev

    def advance_device_hyper(self):
        check(lib.krs_adam_hyper_advance(ptr(self._hyper_dev), stream()))


class Adam(AdamW):
    """keras Adam.  sparse_rows=True: rows without gradient are left alone (moments included) — the per-row form the
    reference's SparseCore path applies (jax/config_conversion.py:259-268); available for compact gradient rows."""

    def __init__(self, learning_rate=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-7, sparse_rows=False):
        super().__init__(learning_rate, 0.0, beta_1, beta_2, epsilon)
        self.sparse_rows = bool(sparse_rows)


class Adagrad(Optimizer):
    def __init__(self, learning_rate=0.001, initial_accumulator_value=0.1, epsilon=1e-7):
        super().__init__(learning_rate)
        self.initial_accumulator_value, self.epsilon = initial_accumulator_value, epsilon

    def _update(self, p, g, touched):
        st = self._slots(p, ("acc",), self.initial_accumulator_value)
        row_len = p.shape[-1] if (touched is not None and p.dim() >= 2) else 1
        check(lib.krs_sgd_adagrad(ptr(p), ptr(st["acc"]), ptr(g), ptr(touched), p.numel(), row_len,
                                  self.learning_rate, self.epsilon, 1, stream()))


    def _update_compact(self, p, cs):
        st = self._slots(p, ("acc",), self.initial_accumulator_value)
        self._rows_apply(p, cs, OPT_ADAGRAD, [self.learning_rate, self.epsilon], st["acc"])


class Ftrl(Optimizer):
    """keras Ftrl with l2_shrinkage_regularization_strength = 0 (the only form the reference's embedding path accepts,
    jax/config_conversion.py:269-285), applied to the rows that received gradient."""

    def __init__(self, learning_rate=0.001, learning_rate_power=-0.5, initial_accumulator_value=0.1,
                 l1_regularization_strength=0.0, l2_regularization_strength=0.0, beta=0.0):
        super().__init__(learning_rate)
        if learning_rate_power > 0:
            raise ValueError(f"`learning_rate_power` needs to be negative or zero. Received: {learning_rate_power}.")
        self.learning_rate_power, self.initial_accumulator_value = learning_rate_power, initial_accumulator_value
        self.l1, self.l2, self.beta = l1_regularization_strength, l2_regularization_strength, beta

    def _ftrl_state(self, p):
        st = self._state.get(id(p))
        if st is None:
            st = self._state[id(p)] = {"accum": torch.full_like(p, self.initial_accumulator_value), "linear": torch.zeros_like(p)}
        return st

    def _update(self, p, g, touched):
        st = self._ftrl_state(p)
        self._opt_apply(p, g, touched, OPT_FTRL, [self.learning_rate, self.learning_rate_power, self.l1, self.l2, self.beta],
                        st["accum"], st["linear"])

    def _update_compact(self, p, cs):
        st = self._ftrl_state(p)
        self._rows_apply(p, cs, OPT_FTRL, [self.learning_rate, self.learning_rate_power, self.l1, self.l2, self.beta],
                         st["accum"], st["linear"])


class SGD(Optimizer):
    def __init__(self, learning_rate=0.01):
        super().__init__(learning_rate)

    def _update_compact(self, p, cs):
        self._rows_apply(p, cs, OPT_SGD, [self.learning_rate])

    def _update(self, p, g, touched):
        row_len = p.shape[-1] if (touched is not None and p.dim() >= 2) else 1
        check(lib.krs_sgd_adagrad(ptr(p), None, ptr(g), ptr(touched), p.numel(), row_len, self.learning_rate, 0.0, 0,
                                  stream()))


def get(name_or_opt, **kw):
    if isinstance(name_or_opt, Optimizer):
        return name_or_opt
    return {"adamw": AdamW, "adam": Adam, "adagrad": Adagrad, "sgd": SGD, "ftrl": Ftrl}[str(name_or_opt).lower()](**kw)
