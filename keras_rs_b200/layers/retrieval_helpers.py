"""HardNegativeMining, RemoveAccidentalHits, SamplingProbabilityCorrection — the two-tower TRAINING helpers of
keras_rs.layers, with the reference's names, constructor arguments, call signatures, error
messages and arithmetic, running on the row kernels of csrc/rowops.cu (krs_row_topk with a label boost, krs_row_scatter
for its backward, krs_remove_accidental_hits, krs_sampling_prob_correction).  CUDA tensors only — like the rest of the
package there is no CPU path.  Gradients flow to `logits` (row selection: scatter of the selected columns; the two additive
corrections: identity); labels, ids and sampling probabilities receive none.

References: hard_negative_mining.py:43-94, remove_accidental_hits.py:32-97,
sampling_probability_correction.py:39-63 (all under keras_rs/src/layers/retrieval/)."""
from __future__ import annotations

from typing import Any

import numpy as np
import torch

from .. import _lib as L
from .._lib import check, lib, ptr, stream
from .base import Layer, register

# hard_negative_mining.py:9  (ml_dtypes.finfo("float32").max / 100.0)
MAX_FLOAT = float(np.finfo(np.float32).max) / 100.0
# remove_accidental_hits.py:9 (ml_dtypes.finfo("float32").smallest_normal / 100.0) — a SUBNORMAL number: added to an
# fp32 logit it changes nothing unless the logit is itself ~0, exactly as in the reference
SMALLEST_FLOAT = float(np.finfo(np.float32).tiny) / 100.0


def _shapes_compatible(a, b) -> bool:
    """keras_utils.check_shapes_compatible (utils/keras_utils.py:54-64) for static shapes."""
    return len(a) == len(b) and all(int(x) == int(y) for x, y in zip(a, b))


class _RowSelectFn(torch.autograd.Function):
    """top-k columns of every row ordered by logits + labels * MAX_FLOAT; returns the selected logits and labels."""

    @staticmethod
    def forward(ctx, x, y, k):
        rows, n = x.shape
        out_l = torch.empty((rows, k), device=x.device, dtype=torch.float32)
        out_y = torch.empty((rows, k), device=x.device, dtype=torch.float32)
        idx = torch.empty((rows, k), device=x.device, dtype=torch.int32)
        check(lib.krs_row_topk(ptr(x), rows, n, n, ptr(y), n, MAX_FLOAT, k, ptr(out_l), ptr(idx), ptr(y), n, ptr(out_y), None, 0, None,
                               stream()))
        ctx.save_for_backward(idx)
        ctx.n = n
        ctx.mark_non_differentiable(out_y)
        return out_l, out_y

    @staticmethod
    def backward(ctx, g_l, g_y):
        (idx,) = ctx.saved_tensors
        rows, k = idx.shape
        g = g_l.contiguous()
        dx = torch.empty((rows, ctx.n), device=g.device, dtype=torch.float32)
        check(lib.krs_row_scatter(ptr(g), ptr(idx), rows, k, ctx.n, ptr(dx), stream()))
        return dx, None, None


class _AdditiveFn(torch.autograd.Function):
    """logits + (a correction that does not depend on the logits): identity gradient."""

    @staticmethod
    def forward(ctx, x, kind, aux, ids, per_row, scalar):
        out = torch.empty_like(x)
        if kind == "hits":
            rows, n = x.shape
            check(lib.krs_remove_accidental_hits(ptr(x), ptr(aux), ptr(ids), 1 if ids.dtype == torch.int64 else 0, per_row, rows, n,
                                                 scalar, ptr(out), stream()))
        else:
            check(lib.krs_sampling_prob_correction(ptr(x), ptr(aux), x.numel(), max(aux.numel(), 1), scalar, ptr(out), stream()))
        return out

    @staticmethod
    def backward(ctx, g):
        return g, None, None, None, None, None


@register("keras_rs.layers.HardNegativeMining")
class HardNegativeMining(Layer):
    """Keeps, per row, the positive candidate and the `num_hard_negatives` highest-scoring negatives
    (hard_negative_mining.py:43-94).  The reference asks `top_k(..., sorted=False)`; this returns the selected columns in
    descending order of `logits + labels * MAX_FLOAT`, one valid instance of that unspecified order."""

    def __init__(self, num_hard_negatives: int, **kwargs: Any) -> None:
        super().__init__(**kwargs)
        self._num_hard_negatives = num_hard_negatives
        self.built = True

    def call(self, logits: torch.Tensor, labels: torch.Tensor):
        L.require_cuda(logits, "logits")
        L.require_cuda(labels, "labels", dtype=None)
        num_logits = logits.shape[-1]
        num_sampled = min(self._num_hard_negatives + 1, num_logits)              # :70-75
        lead = logits.shape[:-1]
        x = logits.reshape(-1, num_logits).contiguous()
        y = labels.reshape(-1, num_logits).to(torch.float32).contiguous()
        out_logits, out_labels = _RowSelectFn.apply(x, y, num_sampled)           # :88-94
        return out_logits.reshape(*lead, num_sampled), out_labels.reshape(*lead, num_sampled).to(labels.dtype)

    def compute_output_shape(self, logits_shape, labels_shape=None):
        out = tuple(logits_shape[:-1]) + (min(self._num_hard_negatives + 1, logits_shape[-1]),)
        return out, out

    def get_config(self) -> dict[str, Any]:
        config = super().get_config()
        config.update({"num_hard_negatives": self._num_hard_negatives})
        return config


@register("keras_rs.layers.RemoveAccidentalHits")
class RemoveAccidentalHits(Layer):
    """Adds SMALLEST_FLOAT to the logits of negatives that carry the positive candidate's id
    (remove_accidental_hits.py:32-97), literally: the positive id is `take(candidate_ids, argmax(labels))` on the
    FLATTENED id tensor (`ops.take` without an axis, :92-93), `duplicate = (ids == positive_id) - labels` (:94-96)."""

    def __init__(self, **kwargs: Any) -> None:
        super().__init__(**kwargs)
        self.built = True

    def call(self, logits: torch.Tensor, labels: torch.Tensor, candidate_ids: torch.Tensor) -> torch.Tensor:
        labels_shape, logits_shape, ids_shape = tuple(labels.shape), tuple(logits.shape), tuple(candidate_ids.shape)
        if not _shapes_compatible(labels_shape, logits_shape):
            raise ValueError("`labels` and `logits` should have the same shape. Received: "
                             f"`labels.shape` = {labels_shape}, `logits.shape` = {logits_shape}.")
        ids_rank = len(ids_shape)
        if not _shapes_compatible(labels_shape[len(labels_shape) - ids_rank:] if ids_rank else (), ids_shape):
            raise ValueError("`candidate_ids` should have the same shape as the last dimensions of `labels`. Received: "
                             f"`candidate_ids.shape` = {ids_shape}, `labels.shape` = {labels_shape}.")
        L.require_cuda(logits, "logits")
        L.require_cuda(labels, "labels", dtype=None)
        L.require_cuda(candidate_ids, "candidate_ids", dtype=None)
        n = logits_shape[-1] if logits_shape else 1
        x = logits.reshape(-1, n).contiguous()
        y = labels.reshape(-1, n).to(torch.float32).contiguous()
        ids = candidate_ids if candidate_ids.dtype in (torch.int32, torch.int64) else candidate_ids.to(torch.int64)
        # ids broadcast over the leading dims of labels (:84-90).  The kernel takes one shared row of ids or one row per
        # logits row; anything in between is expanded.  The positive id is taken from the FLATTENED ids (:92-93).
        if ids_rank <= 1:
            ids2, per_row = ids.reshape(-1).contiguous(), 0
        else:
            ids2 = ids.reshape((1,) * (len(labels_shape) - ids_rank) + ids_shape).expand(labels_shape).reshape(-1, n).contiguous()
            per_row = 1
        return _AdditiveFn.apply(x, "hits", y, ids2, per_row, SMALLEST_FLOAT).reshape(logits_shape)

    def compute_output_shape(self, logits_shape, *unused):
        return tuple(logits_shape)


@register("keras_rs.layers.SamplingProbabilityCorrection")
class SamplingProbabilityCorrection(Layer):
    """`logits - log(clip(p, epsilon, 1))` (sampling_probability_correction.py:39-58)."""

    def __init__(self, epsilon: float = 1e-6, **kwargs: Any) -> None:
        super().__init__(**kwargs)
        self.epsilon = epsilon
        self.built = True

    def call(self, logits: torch.Tensor, candidate_sampling_probability: torch.Tensor) -> torch.Tensor:
        L.require_cuda(logits, "logits")
        L.require_cuda(candidate_sampling_probability, "candidate_sampling_probability", dtype=None)
        p = candidate_sampling_probability.to(torch.float32)
        if p.dim() > logits.dim() or tuple(logits.shape[logits.dim() - p.dim():]) != tuple(p.shape):
            p = p.expand(logits.shape)                                            # general numpy broadcasting, off the fast path
        p = p.contiguous()
        return _AdditiveFn.apply(logits.contiguous(), "prob", p, None, 0, float(self.epsilon))

    def compute_output_shape(self, logits_shape, *unused):
        return tuple(logits_shape)

    def get_config(self) -> dict[str, Any]:
        config = super().get_config()
        config.update({"epsilon": self.epsilon})
        return config
