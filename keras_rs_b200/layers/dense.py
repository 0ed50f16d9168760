"""Dense — stand-in for keras.layers.Dense on the hot path: y = act(x @ kernel + bias), kernel
(in, out) (examples/dcn.py:444-447, examples/ml_perf/model.py:214-266).  One GEMM whose epilogue
applies bias + activation (csrc/cross_dense.cu)."""
from __future__ import annotations

import torch

from .. import _lib as L
from .. import initializers, ops
from .base import Layer


def resolve_activation(act):
    """keras.activations.get: None -> linear; known names run inside the GEMM epilogue; any other
    callable is applied by the caller on the pre-activation (user code, outside the kernels)."""
    if act is None or act == "linear":
        return 0, None, "linear"
    if isinstance(act, str):
        if act not in L.ACT:
            raise ValueError(f"Unknown activation function: {act!r}")
        return L.ACT[act], None, act
    if callable(act):
        name = getattr(act, "__name__", None)
        if name in L.ACT and name not in (None,):
            return L.ACT[name], None, name
        return 0, act, name or repr(act)
    raise ValueError(f"Could not interpret activation: {act!r}")


class Dense(Layer):
    def __init__(self, units, activation=None, use_bias=True, kernel_initializer="glorot_uniform",
                 bias_initializer="zeros", kernel_regularizer=None, bias_regularizer=None, **kwargs):
        super().__init__(**kwargs)
        self.units = int(units)
        self.use_bias = use_bias
        self.activation = activation
        self._act_id, self._act_fn, self._act_name = resolve_activation(activation)
        if self._act_id == L.ACT["swish"]:
            # the Dense backward keeps only y = act(z), from which swish'(z) cannot be recovered (krs_dense_bwd rejects it):
            # swish / silu therefore runs as linear GEMM + the activation applied (and differentiated) on the pre-activation,
            # exactly like a user callable.  FeatureCross keeps z and runs swish inside its epilogue.
            self._act_id, self._act_fn = 0, torch.nn.functional.silu
        self.kernel_initializer = initializers.get(kernel_initializer)
        self.bias_initializer = initializers.get(bias_initializer)
        self.kernel_regularizer = kernel_regularizer
        self.bias_regularizer = bias_regularizer

    def build(self, input_shape):
        k = int(input_shape[-1])
        self.kernel = self.add_weight("kernel", (k, self.units), self.kernel_initializer)
        self.bias = self.add_weight("bias", (self.units,), self.bias_initializer) if self.use_bias else None
        self.built = True

    def call(self, x):
        L.require_cuda(x, "inputs")
        lead = x.shape[:-1]
        x2 = x.reshape(-1, x.shape[-1])
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        y = ops.dense(x2, self.kernel, self.bias, self._act_id)
        if self._act_fn is not None:
            y = self._act_fn(y)
        return y.reshape(*lead, self.units)

    def compute_output_shape(self, input_shape):
        return tuple(input_shape[:-1]) + (self.units,)

    def get_config(self):
        c = super().get_config()
        c.update(units=self.units, activation=self._act_name, use_bias=self.use_bias,
                 kernel_initializer=initializers.serialize(self.kernel_initializer),
                 bias_initializer=initializers.serialize(self.bias_initializer))
        return c
