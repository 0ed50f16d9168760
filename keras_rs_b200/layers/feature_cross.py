"""FeatureCross — drop-in for keras_rs.layers.FeatureCross
(keras_rs/src/layers/feature_interaction/feature_cross.py:14-222).

Same constructor arguments and defaults (:93-108), same lazy build and weight order
([down_proj kernel (D,P)]?, kernel (P|D, D), bias (D)?; feature_cross_test.py:28-32,41-47), same
errors (:124-128, :175-179), same call(x0, x=None) contract for rank-N inputs.  The arithmetic is
one GEMM with the cross fused in its epilogue (csrc/cross_dense.cu), not keras.ops.
"""
from __future__ import annotations

from typing import Any

import torch

from .. import _lib as L
from .. import initializers, ops
from .base import Layer, register
from .dense import resolve_activation


@register("keras_rs.layers.FeatureCross")
class FeatureCross(Layer):
    def __init__(self, projection_dim: int | None = None, diag_scale: float | None = 0.0, use_bias: bool = True,
                 pre_activation=None, kernel_initializer="glorot_uniform", bias_initializer="zeros",
                 kernel_regularizer=None, bias_regularizer=None, **kwargs: Any) -> None:
        super().__init__(**kwargs)
        self.projection_dim = projection_dim
        self.diag_scale = diag_scale
        self.use_bias = use_bias
        self.pre_activation = pre_activation
        self._act_id, self._act_fn, self._act_name = resolve_activation(pre_activation)
        self.kernel_initializer = initializers.get(kernel_initializer)
        self.bias_initializer = initializers.get(bias_initializer)
        self.kernel_regularizer = kernel_regularizer
        self.bias_regularizer = bias_regularizer
        self.supports_masking = True
        if self.diag_scale is not None and self.diag_scale < 0.0:      # feature_cross.py:124-128
            raise ValueError(f"`diag_scale` should be non-negative. Received: `diag_scale={self.diag_scale}`")

    def build(self, input_shape, *unused) -> None:
        last_dim = int(input_shape[-1])                                 # :131
        if self.projection_dim is not None:                            # :133-140 (no bias, no activation)
            self.down_proj_kernel = self.add_weight(
                "down_proj_kernel", (last_dim, int(self.projection_dim)),
                initializers.clone_initializer(self.kernel_initializer))
        else:
            self.down_proj_kernel = None
        k_in = last_dim if self.projection_dim is None else int(self.projection_dim)
        self.kernel = self.add_weight("kernel", (k_in, last_dim), initializers.clone_initializer(self.kernel_initializer))
        self.bias = (self.add_weight("bias", (last_dim,), initializers.clone_initializer(self.bias_initializer))
                     if self.use_bias else None)
        self.built = True

    def call(self, x0: torch.Tensor, x: torch.Tensor | None = None) -> torch.Tensor:
        same = x is None or x is x0
        if x is None:                                                  # :172-173
            x = x0
        if tuple(x0.shape) != tuple(x.shape):                          # :175-179
            raise ValueError("`x0` and `x` should have the same shape. Received: "
                             f"`x.shape` = {tuple(x.shape)}, `x0.shape` = {tuple(x0.shape)}")
        L.require_cuda(x0, "x0")
        L.require_cuda(x, "x")
        shape = x0.shape
        D = shape[-1]
        x0_2 = x0.reshape(-1, D)
        x_2 = x0_2 if same else x.reshape(-1, D)
        if not x0_2.is_contiguous():
            x0_2 = x0_2.contiguous()
            if same:
                x_2 = x0_2
        if not x_2.is_contiguous():
            x_2 = x_2.contiguous()
        diag = self.diag_scale if self.diag_scale else 0.0             # `if self.diag_scale:` :191
        if self._act_fn is None:
            y = ops.feature_cross(x0_2, x_2, self.down_proj_kernel, self.kernel, self.bias, diag, self._act_id, same)
        else:
            # arbitrary Python callable as pre_activation (e.g. ops.zeros_like in
            # feature_cross_test.py:75-79): GEMM+bias in the kernel, callable by the caller, cross in
            # the combine kernel.
            h = x_2 if self.down_proj_kernel is None else ops.linear_no_bias(x_2, self.down_proj_kernel)
            z = ops.dense(h, self.kernel, self.bias, 0)
            a = self._act_fn(z)
            if not isinstance(a, torch.Tensor):
                a = torch.as_tensor(a, device=z.device, dtype=z.dtype)
            y = ops.cross_combine(x0_2, x_2, a.contiguous(), diag)
        return y.reshape(shape)

    def compute_output_shape(self, x0_shape, x_shape=None):
        return tuple(x0_shape)

    def get_config(self) -> dict[str, Any]:
        c = super().get_config()
        c.update(projection_dim=self.projection_dim, diag_scale=self.diag_scale, use_bias=self.use_bias,
                 pre_activation=self._act_name,
                 kernel_initializer=initializers.serialize(self.kernel_initializer),
                 bias_initializer=initializers.serialize(self.bias_initializer),
                 kernel_regularizer=self.kernel_regularizer, bias_regularizer=self.bias_regularizer)
        return c
