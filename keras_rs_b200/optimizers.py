"""Optimizers with the Keras 3 update rules the examples use (AdamW: examples/dcn.py:127; Adagrad:
examples/ml_perf/main.py:203; SGD), applied by fused CUDA sweeps (csrc/optim.cu).

A parameter that owns a gradient arena (`_krs_arena` + `_krs_touched`, written by the fused
embedding backward) is updated straight from the arena: AdamW still visits every row (that is what
the reference's dense update does) but only reads gradient rows that were touched; SGD / Adagrad
visit touched rows only — identical to the dense update because their step is zero where g = 0.
Otherwise the dense `.grad` is used."""
from __future__ import annotations

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
        arena = getattr(p, "_krs_arena", None)
        if arena is not None:
            return arena, p._krs_touched
        if p.grad is None:
            return None, None
        g = p.grad
        return (g if g.is_contiguous() else g.contiguous()), None

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


class AdamW(Optimizer):
    def __init__(self, learning_rate=0.001, weight_decay=0.004, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
        super().__init__(learning_rate)
        self.weight_decay, self.beta_1, self.beta_2, self.epsilon = weight_decay, beta_1, beta_2, epsilon

    def _update(self, p, g, touched):
        st = self._slots(p, ("m", "v"))
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
    def __init__(self, learning_rate=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
        super().__init__(learning_rate, 0.0, beta_1, beta_2, epsilon)


class Adagrad(Optimizer):
    def __init__(self, learning_rate=0.001, initial_accumulator_value=0.1, epsilon=1e-7):
        super().__init__(learning_rate)
        self.initial_accumulator_value, self.epsilon = initial_accumulator_value, epsilon

    def _update(self, p, g, touched):
        st = self._slots(p, ("acc",), self.initial_accumulator_value)
        row_len = p.shape[-1] if (touched is not None and p.dim() >= 2) else 1
        check(lib.krs_sgd_adagrad(ptr(p), ptr(st["acc"]), ptr(g), ptr(touched), p.numel(), row_len,
                                  self.learning_rate, self.epsilon, 1, stream()))


class SGD(Optimizer):
    def __init__(self, learning_rate=0.01):
        super().__init__(learning_rate)

    def _update(self, p, g, touched):
        row_len = p.shape[-1] if (touched is not None and p.dim() >= 2) else 1
        check(lib.krs_sgd_adagrad(ptr(p), None, ptr(g), ptr(touched), p.numel(), row_len, self.learning_rate, 0.0, 0,
                                  stream()))


def get(name_or_opt, **kw):
    if isinstance(name_or_opt, Optimizer):
        return name_or_opt
    return {"adamw": AdamW, "adam": Adam, "adagrad": Adagrad, "sgd": SGD}[str(name_or_opt).lower()](**kw)
