"""HardNegativeMining, RemoveAccidentalHits, SamplingProbabilityCorrection — the two-tower TRAINING helpers of
keras_rs.layers ("next": they feed the loss of the retrieval model, not the hot path
gather -> cross -> dense this package accelerates).

These are thin host-side mirrors with the reference's names, constructor arguments, call signatures, error messages and
arithmetic, composed from torch tensor ops on whatever device the inputs live on; they launch no kernel of
libkrs_b200.so.  Fusing them into the score epilogue of the tensor-pipe scorer (csrc/topk.cu) is the planned CUDA form.

References: hard_negative_mining.py:43-94, remove_accidental_hits.py:32-97,
sampling_probability_correction.py:39-63 (all under keras_rs/src/layers/retrieval/)."""
from __future__ import annotations

from typing import Any

import numpy as np
import torch

from .base import Layer, register

# hard_negative_mining.py:9  (ml_dtypes.finfo("float32").max / 100.0)
MAX_FLOAT = float(np.finfo(np.float32).max) / 100.0
# remove_accidental_hits.py:9 (ml_dtypes.finfo("float32").smallest_normal / 100.0) — a SUBNORMAL number: added to an
# fp32 logit it changes nothing unless the logit is itself ~0, exactly as in the reference
SMALLEST_FLOAT = float(np.finfo(np.float32).tiny) / 100.0


def _shapes_compatible(a, b) -> bool:
    """keras_utils.check_shapes_compatible (utils/keras_utils.py:54-64) for static shapes."""
    return len(a) == len(b) and all(int(x) == int(y) for x, y in zip(a, b))


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
        num_logits = logits.shape[-1]
        num_sampled = min(self._num_hard_negatives + 1, num_logits)              # :70-75
        boosted = logits + labels.to(logits.dtype) * MAX_FLOAT                   # :88
        _, indices = torch.topk(boosted, k=num_sampled, dim=-1, largest=True, sorted=True)
        return torch.take_along_dim(logits, indices, dim=-1), torch.take_along_dim(labels, indices, dim=-1)

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
        ids = candidate_ids.reshape((1,) * (len(labels_shape) - ids_rank) + ids_shape)       # :84-90
        positive_indices = torch.argmax(labels, dim=-1, keepdim=True)                       # :91
        positive_ids = candidate_ids.reshape(-1)[positive_indices]                          # :92-93 (flattened take)
        duplicate = (positive_ids == ids).to(labels.dtype) - labels                         # :94-96
        return logits + duplicate.to(logits.dtype) * SMALLEST_FLOAT                         # :97

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
        p = candidate_sampling_probability.to(logits.dtype)
        return logits - torch.log(torch.clamp(p, self.epsilon, 1.0))

    def compute_output_shape(self, logits_shape, *unused):
        return tuple(logits_shape)

    def get_config(self) -> dict[str, Any]:
        config = super().get_config()
        config.update({"epsilon": self.epsilon})
        return config
