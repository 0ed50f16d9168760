"""A small Keras-`Layer`-shaped base on torch.nn.Module (under KERAS_BACKEND=torch a Keras layer IS
a torch.nn.Module).  Only what the hot-path layers of keras-rs use: lazy `build`, ordered
`weights`, `get_config` / `from_config`, and the `keras_rs>Name` registration names of
keras_rs/src/api_export.py:14-23."""
from __future__ import annotations

import torch

from .. import initializers

_REGISTRY: dict[str, type] = {}
_UID: dict[str, int] = {}

DEFAULT_DEVICE = "cuda"


def register(path: str):
    """Equivalent of @keras_rs_export("keras_rs.layers.X") -> registered name 'keras_rs>X'."""
    def deco(cls):
        cls._keras_name = "keras_rs>" + path.split(".")[-1]
        _REGISTRY[cls._keras_name] = cls
        return cls
    return deco


def _auto_name(cls_name: str) -> str:
    import re
    snake = re.sub(r"(?<!^)(?=[A-Z])", "_", cls_name).lower()
    n = _UID.get(snake, 0)
    _UID[snake] = n + 1
    return snake if n == 0 else f"{snake}_{n}"


class Layer(torch.nn.Module):
    def __init__(self, name: str | None = None, dtype=None, trainable: bool = True, device=None, **kwargs):
        if kwargs:
            raise TypeError(f"Unrecognized keyword arguments passed to {type(self).__name__}: {kwargs}")
        super().__init__()
        self.name = name or _auto_name(type(self).__name__)
        self.built = False
        self.trainable = trainable
        self._dtype_name = "float32" if dtype is None else str(dtype)
        self._device = device or DEFAULT_DEVICE
        self._weight_order: list[str] = []
        self.supports_masking = False

    # ---- weights --------------------------------------------------------
    def add_weight(self, name, shape, initializer, trainable=True, dtype=torch.float32):
        init = initializers.get(initializer)
        value = init(tuple(shape), dtype=dtype) if callable(init) else torch.zeros(tuple(shape), dtype=dtype)
        p = torch.nn.Parameter(value.to(self._device).contiguous(), requires_grad=trainable and dtype.is_floating_point)
        self.register_parameter(name, p)
        self._weight_order.append(name)
        return p

    @property
    def weights(self):
        out = []
        for n in self._weight_order:
            out.append(getattr(self, n))
        for child in self.children():
            if isinstance(child, Layer):
                out.extend(child.weights)
        return out

    @property
    def trainable_weights(self):
        return [w for w in self.weights if w.requires_grad]

    @property
    def non_trainable_weights(self):
        return [w for w in self.weights if not w.requires_grad]

    @property
    def compute_dtype(self):
        return self._dtype_name

    def get_weights(self):
        return [w.detach().cpu().numpy() for w in self.weights]

    def set_weights(self, values):
        ws = self.weights
        if len(ws) != len(values):
            raise ValueError(f"Layer {self.name} expects {len(ws)} weights, got {len(values)}")
        with torch.no_grad():
            for w, v in zip(ws, values):
                t = torch.as_tensor(v, dtype=w.dtype)
                if tuple(t.shape) != tuple(w.shape):
                    raise ValueError(f"Shape mismatch for weight: {tuple(t.shape)} vs {tuple(w.shape)}")
                w.copy_(t)

    # ---- protocol -------------------------------------------------------
    def build(self, *shapes):
        self.built = True

    def call(self, *args, **kwargs):
        raise NotImplementedError

    def _shape_of(self, a):
        if isinstance(a, torch.Tensor):
            return tuple(a.shape)
        if isinstance(a, (list, tuple)):
            return [self._shape_of(x) for x in a]
        if isinstance(a, dict):
            return {k: self._shape_of(v) for k, v in a.items()}
        return None

    def forward(self, *args, **kwargs):
        if not self.built:
            self.build(*[self._shape_of(a) for a in args])
            self.built = True
        return self.call(*args, **kwargs)

    def get_config(self):
        return {"name": self.name, "trainable": self.trainable, "dtype": self._dtype_name}

    @classmethod
    def from_config(cls, config):
        return cls(**config)


def serialize(layer: Layer) -> dict:
    return {"module": "keras_rs_b200.layers", "class_name": type(layer).__name__,
            "registered_name": getattr(type(layer), "_keras_name", type(layer).__name__),
            "config": layer.get_config()}


def deserialize(obj: dict) -> Layer:
    cls = _REGISTRY.get(obj.get("registered_name"))
    if cls is None:
        raise ValueError(f"Unknown layer: {obj.get('registered_name')}")
    return cls.from_config(dict(obj["config"]))
