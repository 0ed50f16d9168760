"""Embedding / EmbedReduce — the lookup layers of the hot path.

Embedding stands in for keras.layers.Embedding as used at examples/dcn.py:430-435
(`ops.take(embeddings, ids, axis=0)`); EmbedReduce is the drop-in for keras_rs.layers.EmbedReduce
(keras_rs/src/layers/embedding/embed_reduce.py:13-309): same ctor (:133-160), same
`call(inputs, weights=None)` semantics for dense inputs (the torch backend of the reference supports
dense inputs only, embed_reduce_test.py:37-43) and the same errors (:155-159, :184-190).
A single layer call is a fused gather with F=1; multi-table fusion is `DistributedEmbedding` /
`ops.gather_concat`."""
from __future__ import annotations

from typing import Any

import torch

from .. import _lib as L
from .. import initializers, ops
from .base import Layer, register

SUPPORTED_COMBINERS = ("mean", "sum", "sqrtn")   # embed_reduce.py:10


class Embedding(Layer):
    def __init__(self, input_dim: int, output_dim: int, embeddings_initializer="uniform",
                 embeddings_regularizer=None, embeddings_constraint=None, mask_zero: bool = False,
                 weights=None, sparse_grad_arena: bool = False, **kwargs: Any) -> None:
        super().__init__(**kwargs)
        self.input_dim = int(input_dim)
        self.output_dim = int(output_dim)
        self.embeddings_initializer = (initializers.Constant(weights) if weights is not None
                                       else initializers.get(embeddings_initializer))
        self.embeddings_regularizer = embeddings_regularizer
        self.embeddings_constraint = embeddings_constraint
        self.mask_zero = mask_zero
        self.sparse_grad_arena = sparse_grad_arena
        # keras.layers.Embedding builds eagerly in __init__-time `build(None)`; do the same so that
        # `layer.embeddings` can be shared right after construction (basic_retrieval.py:249-257).
        self.embeddings = self.add_weight("embeddings", (self.input_dim, self.output_dim), self.embeddings_initializer)
        self.built = True

    def build(self, *a):
        self.built = True

    def _lookup(self, inputs: torch.Tensor, weights, combiner: str) -> torch.Tensor:
        if not isinstance(inputs, torch.Tensor):
            inputs = torch.as_tensor(inputs)
        if not inputs.is_cuda:
            raise L.KrsError("inputs must live on a CUDA device (keras_rs_b200 has no CPU path)")
        if inputs.dtype not in (torch.int32, torch.int64):
            inputs = inputs.to(torch.int32)          # keras Embedding: non-int ids are cast to int32
        feat = dict(table=self.embeddings, ids=inputs, weights=weights, combiner=combiner)
        return ops.gather_concat([feat], sparse_arena=self.sparse_grad_arena)

    def call(self, inputs: torch.Tensor) -> torch.Tensor:
        ids = inputs
        shape = tuple(ids.shape)
        flat = ids.reshape(-1)
        out = self._lookup(flat.contiguous(), None, "sum")
        return out.reshape(*shape, self.output_dim)

    def compute_output_shape(self, input_shape):
        return tuple(input_shape) + (self.output_dim,)

    def get_config(self):
        c = super().get_config()
        c.update(input_dim=self.input_dim, output_dim=self.output_dim,
                 embeddings_initializer=initializers.serialize(self.embeddings_initializer), mask_zero=self.mask_zero)
        return c


@register("keras_rs.layers.EmbedReduce")
class EmbedReduce(Embedding):
    def __init__(self, input_dim: int, output_dim: int, embeddings_initializer="uniform",
                 embeddings_regularizer=None, embeddings_constraint=None, mask_zero: bool = False, weights=None,
                 combiner: str = "mean", **kwargs: Any) -> None:
        super().__init__(input_dim, output_dim, embeddings_initializer=embeddings_initializer,
                         embeddings_regularizer=embeddings_regularizer, embeddings_constraint=embeddings_constraint,
                         mask_zero=mask_zero, weights=weights, **kwargs)
        if combiner not in SUPPORTED_COMBINERS:                        # embed_reduce.py:155-159
            raise ValueError(f"Invalid `combiner`: '{combiner}', use one of {', '.join(SUPPORTED_COMBINERS)}.")
        self.combiner = combiner

    def call(self, inputs: torch.Tensor, weights: torch.Tensor | None = None) -> torch.Tensor:
        if not isinstance(inputs, torch.Tensor):
            inputs = torch.as_tensor(inputs)
        if inputs.dim() > 2:
            raise ValueError("EmbedReduce expects a 1D tensor to embed or a 2D tensor to embed and reduce "
                             f"(embed_reduce.py:100), got rank {inputs.dim()}")
        if weights is not None:
            if not isinstance(weights, torch.Tensor):
                weights = torch.as_tensor(weights)
            x_shape = tuple(inputs.shape) + (self.output_dim,)
            wr = weights.dim()
            if wr > len(x_shape) or tuple(x_shape[:wr]) != tuple(weights.shape):   # :182-190
                raise ValueError(f"The shape of `weights`: {tuple(weights.shape)} is not compatible with the shape "
                                 f"of `inputs` after embedding: {x_shape}.")
            weights = weights.to(device=self.embeddings.device, dtype=torch.float32)
            if wr < inputs.dim():      # (B,) weights for (B,H) ids: broadcast over H (expand_dims :244-248)
                weights = weights.reshape(-1, 1).expand(tuple(inputs.shape)).contiguous()
        return self._lookup(inputs, weights, self.combiner)

    def compute_output_shape(self, input_shape, weights_shape=None):
        if len(input_shape) <= 1:
            return tuple(input_shape) + (self.output_dim,)
        return tuple(input_shape[:-1]) + (self.output_dim,)

    def get_config(self):
        c = super().get_config()
        c.update(combiner=self.combiner)
        return c
