"""Minimal stand-ins for keras.initializers (the reference resolves names with
keras.initializers.get, feature_cross.py:116-117).  Values are generated on the host with a
seeded torch.Generator and copied to the device once at build time — not on the hot path."""
from __future__ import annotations

import math
from typing import Callable

import torch


def _fans(shape):
    if len(shape) < 1:
        return 1, 1
    if len(shape) == 1:
        return shape[0], shape[0]
    return shape[0], shape[1]


class Initializer:
    name = "initializer"

    def __init__(self, seed: int | None = None):
        self.seed = seed

    def _gen(self):
        g = torch.Generator(device="cpu")
        if self.seed is not None:
            g.manual_seed(int(self.seed))
        else:
            g.seed()
        return g

    def __call__(self, shape, dtype=torch.float32):
        raise NotImplementedError

    def get_config(self):
        return {"class_name": type(self).__name__, "config": {"seed": self.seed}}

    def clone(self):
        """keras_rs.utils.clone_initializer (keras_utils.py:30-51): fresh instance, fresh seed."""
        return type(self)(**self.get_config()["config"])


class Zeros(Initializer):
    def __init__(self, seed=None):
        super().__init__(None)

    def __call__(self, shape, dtype=torch.float32):
        return torch.zeros(tuple(shape), dtype=dtype)

    def get_config(self):
        return {"class_name": "Zeros", "config": {}}


class Ones(Initializer):
    def __init__(self, seed=None):
        super().__init__(None)

    def __call__(self, shape, dtype=torch.float32):
        return torch.ones(tuple(shape), dtype=dtype)

    def get_config(self):
        return {"class_name": "Ones", "config": {}}


class GlorotUniform(Initializer):
    def __call__(self, shape, dtype=torch.float32):
        fi, fo = _fans(shape)
        limit = math.sqrt(6.0 / (fi + fo))
        return (torch.rand(tuple(shape), generator=self._gen(), dtype=dtype) * 2 - 1) * limit


class RandomUniform(Initializer):
    """keras 'uniform' — the keras.layers.Embedding default, U(-0.05, 0.05)."""

    def __init__(self, minval=-0.05, maxval=0.05, seed=None):
        super().__init__(seed)
        self.minval, self.maxval = minval, maxval

    def __call__(self, shape, dtype=torch.float32):
        return torch.rand(tuple(shape), generator=self._gen(), dtype=dtype) * (self.maxval - self.minval) + self.minval

    def get_config(self):
        return {"class_name": "RandomUniform", "config": {"minval": self.minval, "maxval": self.maxval, "seed": self.seed}}


class VarianceScaling(Initializer):
    """TableConfig default: VarianceScaling(mode="fan_out") (distributed_embedding_config.py:54-61);
    ml_perf MLPs use VarianceScaling uniform (examples/ml_perf/model.py:214-266)."""

    def __init__(self, scale=1.0, mode="fan_in", distribution="truncated_normal", seed=None):
        super().__init__(seed)
        self.scale, self.mode, self.distribution = scale, mode, distribution

    def __call__(self, shape, dtype=torch.float32):
        fi, fo = _fans(shape)
        n = {"fan_in": fi, "fan_out": fo, "fan_avg": (fi + fo) / 2.0}[self.mode]
        s = self.scale / max(1.0, n)
        g = self._gen()
        if self.distribution == "uniform":
            limit = math.sqrt(3.0 * s)
            return (torch.rand(tuple(shape), generator=g, dtype=dtype) * 2 - 1) * limit
        std = math.sqrt(s)
        if self.distribution == "truncated_normal":
            std = std / 0.87962566103423978
            t = torch.empty(tuple(shape), dtype=dtype)
            torch.nn.init.trunc_normal_(t, 0.0, std, -2 * std, 2 * std, generator=g)
            return t
        return torch.randn(tuple(shape), generator=g, dtype=dtype) * std

    def get_config(self):
        return {"class_name": "VarianceScaling",
                "config": {"scale": self.scale, "mode": self.mode, "distribution": self.distribution, "seed": self.seed}}


class Constant(Initializer):
    """An explicit tensor/array (e.g. `weights=` of Embedding, candidate embeddings)."""

    def __init__(self, value, seed=None):
        super().__init__(None)
        self.value = value

    def __call__(self, shape, dtype=torch.float32):
        t = torch.as_tensor(self.value).detach().to("cpu", dtype)
        return t.reshape(tuple(shape)).clone()

    def get_config(self):
        return {"class_name": "Constant", "config": {"value": "<tensor>"}}

    def clone(self):
        return Constant(self.value)


_BY_NAME = {
    "zeros": Zeros, "ones": Ones, "glorot_uniform": GlorotUniform, "uniform": RandomUniform,
    "random_uniform": RandomUniform, "variance_scaling": VarianceScaling,
    "Zeros": Zeros, "Ones": Ones, "GlorotUniform": GlorotUniform, "RandomUniform": RandomUniform,
    "VarianceScaling": VarianceScaling,
}


def get(identifier) -> Initializer | Callable:
    if identifier is None:
        return None
    if isinstance(identifier, Initializer):
        return identifier
    if isinstance(identifier, str):
        if identifier not in _BY_NAME:
            raise ValueError(f"Unknown initializer: {identifier!r}")
        return _BY_NAME[identifier]()
    if isinstance(identifier, dict):
        return _BY_NAME[identifier["class_name"]](**identifier.get("config", {}))
    if callable(identifier):
        return identifier
    raise ValueError(f"Could not interpret initializer identifier: {identifier!r}")


def clone_initializer(init):
    """keras_rs/src/utils/keras_utils.py:30-51."""
    if isinstance(init, Initializer):
        return init.clone()
    return init


def serialize(init):
    if isinstance(init, Initializer):
        return init.get_config()
    return getattr(init, "__name__", repr(init))
