"""DotInteraction — drop-in for keras_rs.layers.DotInteraction
(keras_rs/src/layers/feature_interaction/dot_interaction.py:12-234): same ctor (:84-94), same
validation errors (:152-167), same output ordering (row-major lower triangle, :118-132) and
`skip_gather` zero-masked variant (:182-192).  No weights."""
from __future__ import annotations

from typing import Any

import torch

from .. import _lib as L
from .. import ops
from .base import Layer, register


@register("keras_rs.layers.DotInteraction")
class DotInteraction(Layer):
    def __init__(self, self_interaction: bool = False, skip_gather: bool = False, **kwargs: Any) -> None:
        super().__init__(**kwargs)
        self.self_interaction = self_interaction
        self.skip_gather = skip_gather

    def _get_lower_triangular_indices(self, num_features: int) -> list[int]:
        out = []                                                       # dot_interaction.py:118-132
        for i in range(num_features):
            k = i + 1 if self.self_interaction else i
            for j in range(k):
                out.append(i * num_features + j)
        return out

    def build(self, *a):
        self.built = True

    def call(self, inputs: list[torch.Tensor]) -> torch.Tensor:
        shape = tuple(inputs[0].shape)
        for idx, t in enumerate(inputs):
            if len(shape) != 2:                                        # :155-160
                raise ValueError("All feature tensors inside `inputs` should have rank 2. "
                                 f"Received rank {len(shape)} at index {idx}.")
            if tuple(t.shape) != shape:                                # :162-167
                raise ValueError("All feature tensors in `inputs` should have the same shape. Found at least one "
                                 f"conflict: shape = {shape} at index 0 and shape = {tuple(t.shape)} at index {idx}.")
        for t in inputs:
            L.require_cuda(t, "inputs")
        if len(inputs) > 32:
            raise ValueError(f"DotInteraction supports at most 32 features per call, got {len(inputs)}")
        return ops.dot_interaction(list(inputs), self.self_interaction, self.skip_gather)

    def compute_output_shape(self, input_shape):
        n = len(input_shape)                                           # :207-222
        b = input_shape[0][0]
        d = n * (n + 1) // 2 if self.self_interaction else n * (n - 1) // 2
        if self.skip_gather:
            d = n * n
        return (b, d)

    def get_config(self) -> dict[str, Any]:
        c = super().get_config()
        c.update(self_interaction=self.self_interaction, skip_gather=self.skip_gather)
        return c
