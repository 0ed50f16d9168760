"""TableConfig / FeatureConfig / DistributedEmbedding — drop-ins for
keras_rs/src/layers/embedding/distributed_embedding_config.py:13-132 and the default-device half of
keras_rs/src/layers/embedding/base_distributed_embedding.py (:468-938).

The reference loops over features in Python and issues one EmbedReduce lookup per feature
(:910-928).  Here the nested feature structure is flattened once and EVERY feature of a call goes
through one fused multi-table kernel; the result is handed back in the caller's nesting
(:802-808).  `call(..., concat=True)` additionally returns the concatenated (B, sum E) activation
the models feed to FeatureCross / DotInteraction without any copy."""
from __future__ import annotations

import dataclasses
from typing import Any

import numpy as np
import torch

from .. import _lib as L
from .. import initializers, ops
from .base import Layer, register
from .embedding import SUPPORTED_COMBINERS


@dataclasses.dataclass(eq=False, unsafe_hash=False)
class TableConfig:
    """distributed_embedding_config.py:13-87."""
    name: str
    vocabulary_size: int
    embedding_dim: int
    initializer: Any = dataclasses.field(default_factory=lambda: initializers.VarianceScaling(mode="fan_out"))
    optimizer: Any = "adam"
    combiner: str = "mean"
    placement: str = "auto"
    max_ids_per_partition: int = 256
    max_unique_ids_per_partition: int = 256

    def get_config(self):
        return dict(name=self.name, vocabulary_size=self.vocabulary_size, embedding_dim=self.embedding_dim,
                    initializer=initializers.serialize(initializers.get(self.initializer)), optimizer=self.optimizer,
                    combiner=self.combiner, placement=self.placement,
                    max_ids_per_partition=self.max_ids_per_partition,
                    max_unique_ids_per_partition=self.max_unique_ids_per_partition)

    @classmethod
    def from_config(cls, config):
        config = dict(config)
        config["initializer"] = initializers.get(config["initializer"])
        return cls(**config)


@dataclasses.dataclass(eq=False, unsafe_hash=False)
class FeatureConfig:
    """distributed_embedding_config.py:91-132."""
    name: str
    table: TableConfig
    input_shape: tuple
    output_shape: tuple

    def get_config(self):
        return dict(name=self.name, table=self.table.get_config(), input_shape=tuple(self.input_shape),
                    output_shape=tuple(self.output_shape))

    @classmethod
    def from_config(cls, config):
        config = dict(config)
        config["table"] = TableConfig.from_config(config["table"])
        return cls(**config)


def ragged_to_dense_inputs(inputs, weights=None, dense_row_length: int | None = None, device="cuda"):
    """base_distributed_embedding.py:31-92 (_ragged_to_dense_inputs): a ragged batch of id lists (object ndarray, list of
    lists, or a CSR pair `(values, row_splits)`) becomes a dense (B, L) id tensor padded with 0 and a float32 weight
    tensor that is 0 on the padding (1, or the given weights, on real ids) — so `sum` / `mean` / `sqrtn` combiners see
    exactly the ragged row.  L = dense_row_length or the longest row; longer rows are an error.  Non-ragged inputs are
    returned unchanged.  The padding is built with one vectorised scatter (no per-row Python loop)."""
    x, w = inputs, weights
    csr = isinstance(x, tuple) and len(x) == 2 and not isinstance(x[0], (list, tuple)) and np.ndim(x[1]) == 1 and np.ndim(x[0]) == 1
    if isinstance(x, torch.Tensor) or (isinstance(x, np.ndarray) and x.dtype != object):
        return inputs, weights
    if csr:
        values, splits = np.asarray(x[0]), np.asarray(x[1]).astype(np.int64)
        wvals = None if w is None else np.asarray(w[0] if isinstance(w, tuple) else w, dtype=np.float32).reshape(-1)
    else:
        rows = [np.asarray(r) for r in x]
        if len(rows) == 0 or all(r.ndim == 0 for r in rows):
            return inputs, weights                              # a plain list of scalars: not ragged
        lens = np.array([len(r) for r in rows], dtype=np.int64)
        splits = np.concatenate([[0], np.cumsum(lens)])
        values = np.concatenate(rows) if splits[-1] else np.zeros((0,), np.int64)
        wvals = None if w is None else np.concatenate([np.asarray(r, dtype=np.float32) for r in w])
    lens = np.diff(splits)
    B = len(lens)
    L_ = int(dense_row_length) if dense_row_length is not None else (int(lens.max()) if B else 0)
    if B and int(lens.max()) > L_:
        raise ValueError(f"ragged row of length {int(lens.max())} exceeds the dense row length {L_}")
    if wvals is not None and len(wvals) != len(values):
        raise ValueError("ragged `weights` must have the same row lengths as `inputs`")
    row = np.repeat(np.arange(B), lens)
    col = np.arange(len(values)) - np.repeat(splits[:-1], lens)
    ids = np.zeros((B, L_), dtype=values.dtype if values.dtype.kind in "iu" else np.int64)
    ids[row, col] = values
    wts = np.zeros((B, L_), dtype=np.float32)
    wts[row, col] = 1.0 if wvals is None else wvals
    ids_t = torch.from_numpy(ids if ids.dtype in (np.int32, np.int64) else ids.astype(np.int64))
    return ids_t.to(device), torch.from_numpy(wts).to(device)


def _flatten(struct, path=()):
    """Deterministic flattening of nested dict / list / tuple structures -> [(path, leaf)]."""
    if isinstance(struct, dict):
        out = []
        for k in sorted(struct.keys(), key=str):
            out.extend(_flatten(struct[k], path + (k,)))
        return out
    if isinstance(struct, (list, tuple)) and not isinstance(struct, torch.Size):
        out = []
        for i, v in enumerate(struct):
            out.extend(_flatten(v, path + (i,)))
        return out
    return [(path, struct)]


def _is_ragged_leaf(v) -> bool:
    if isinstance(v, torch.Tensor):
        return False
    if isinstance(v, np.ndarray):
        return v.dtype == object
    if isinstance(v, tuple) and len(v) == 2 and np.ndim(v[0]) == 1 and np.ndim(v[1]) == 1 and not isinstance(v[0], (list, tuple)):
        return True                                            # CSR (values, row_splits)
    return isinstance(v, (list, tuple)) and len(v) > 0 and isinstance(v[0], (list, tuple, np.ndarray))


def _flatten_inputs(struct, flat_configs):
    """Like _flatten, but stops at the leaves of the FEATURE structure (a ragged feature is itself a list of lists)."""
    out = []
    for path, _ in flat_configs:
        cur = struct
        for k in path:
            cur = cur[k]
        out.append((path, cur))
    return out


def _has_ragged(inputs, flat_configs) -> bool:
    try:
        return any(_is_ragged_leaf(v) for _, v in _flatten_inputs(inputs, flat_configs))
    except (KeyError, IndexError, TypeError):
        return False


def _pack_like(struct, leaves_iter):
    if isinstance(struct, dict):
        # preserve the caller's key order while consuming leaves in the sorted order used by _flatten
        vals = {k: None for k in struct}
        for k in sorted(struct.keys(), key=str):
            vals[k] = _pack_like(struct[k], leaves_iter)
        return vals
    if isinstance(struct, (list, tuple)):
        return type(struct)(_pack_like(v, leaves_iter) for v in struct)
    return next(leaves_iter)


@register("keras_rs.layers.DistributedEmbedding")
class DistributedEmbedding(Layer):
    def __init__(self, feature_configs, table_stacking="auto", sparse_grad_arena: bool = False, **kwargs: Any) -> None:
        super().__init__(**kwargs)
        self.feature_configs = feature_configs
        self.table_stacking = table_stacking
        self.sparse_grad_arena = sparse_grad_arena
        self._flat = _flatten(feature_configs)
        if not self._flat:
            raise ValueError("`feature_configs` must contain at least one FeatureConfig")
        self._tables: list[TableConfig] = []
        self._feature_table: list[int] = []
        for path, fc in self._flat:
            if not isinstance(fc, FeatureConfig):
                raise ValueError(f"Expected FeatureConfig at {path}, got {type(fc)}")
            t = fc.table
            if t.placement not in ("auto", "default_device"):
                # base_distributed_embedding.py:560-567,990-1013: sparsecore placement is TPU-only
                raise ValueError(f"Placement '{t.placement}' (sparsecore) is not supported on this backend; "
                                 "use 'auto' or 'default_device'.")
            if t.combiner not in SUPPORTED_COMBINERS:
                raise ValueError(f"Invalid `combiner`: '{t.combiner}', use one of {', '.join(SUPPORTED_COMBINERS)}.")
            for i, existing in enumerate(self._tables):     # one table per TableConfig OBJECT (:836-852)
                if existing is t:
                    self._feature_table.append(i)
                    break
            else:
                self._tables.append(t)
                self._feature_table.append(len(self._tables) - 1)
        names = [t.name for t in self._tables]
        if len(set(names)) != len(names):
            raise ValueError(f"Table names must be unique, got {names}")
        self._table_params: list[torch.nn.Parameter] = []

    def build(self, *a) -> None:
        if self._table_params:
            self.built = True
            return
        for t in self._tables:
            p = self.add_weight(f"table_{t.name}".replace(".", "_"), (int(t.vocabulary_size), int(t.embedding_dim)),
                                t.initializer)
            self._table_params.append(p)
        self.built = True

    def get_embedding_tables(self) -> dict[str, torch.Tensor]:
        """base_distributed_embedding.py:930-938."""
        if not self.built:
            self.build()
        return {t.name: p for t, p in zip(self._tables, self._table_params)}

    def preprocess(self, inputs, weights=None, training: bool = False):
        """base_distributed_embedding.py:630-729: on the default device this is pure structure
        shuffling; it returns the dict form that `call` also accepts (:721-738)."""
        fin = {p: v for p, v in _flatten_inputs(inputs, self._flat)}
        fw = None if weights is None else {p: v for p, v in _flatten_inputs(weights, self._flat)}
        # ragged / CSR features are densified to their configured valence with a 0-weight mask (:862-908)
        use_w = fw is not None
        new_w = {}
        for path, fc in self._flat:
            valence = None if len(fc.input_shape) <= 1 else fc.input_shape[1]
            x, w = ragged_to_dense_inputs(fin[path], None if fw is None else fw[path], valence, device=self._device)
            use_w = use_w or (w is not None)
            fin[path], new_w[path] = x, w
        d = {"inputs": fin}
        if use_w:
            d["weights"] = new_w
        return {"preprocessed_inputs_per_placement": {"default_device": d}}

    def _check_shape(self, fc: FeatureConfig, ids: torch.Tensor):
        """Input rank / static dims vs FeatureConfig.input_shape (:1141-1188)."""
        exp = tuple(fc.input_shape)
        got = tuple(ids.shape)
        if len(exp) != len(got) or any(e is not None and e != g for e, g in zip(exp[1:], got[1:])):
            raise ValueError(f"Feature '{fc.name}': input shape {got} is incompatible with the configured "
                             f"input_shape {exp}")

    def call(self, inputs, weights=None, training: bool = False, concat: bool = False):
        if isinstance(inputs, dict) and "preprocessed_inputs_per_placement" in inputs:
            pp = inputs["preprocessed_inputs_per_placement"]["default_device"]
            flat_in = [pp["inputs"][p] for p, _ in self._flat]
            flat_w = [pp["weights"][p] for p, _ in self._flat] if "weights" in pp else None
        elif _has_ragged(inputs, self._flat):
            return self.call(self.preprocess(inputs, weights, training), training=training, concat=concat)
        else:
            fi = _flatten(inputs)
            if [p for p, _ in fi] != [p for p, _ in self._flat]:
                raise ValueError("`inputs` must have the same nested structure as `feature_configs`")
            flat_in = [v for _, v in fi]
            flat_w = None
            if weights is not None:
                fw = _flatten(weights)
                if [p for p, _ in fw] != [p for p, _ in self._flat]:
                    raise ValueError("`weights` must have the same nested structure as `feature_configs`")
                flat_w = [v for _, v in fw]
        feats = []
        for i, ((path, fc), ids) in enumerate(zip(self._flat, flat_in)):
            if not isinstance(ids, torch.Tensor):
                ids = torch.as_tensor(ids)
            if not ids.is_cuda:
                raise L.KrsError("inputs must live on a CUDA device (keras_rs_b200 has no CPU path)")
            self._check_shape(fc, ids)
            if ids.dim() == 2 and ids.shape[1] == 1 and len(fc.output_shape) == 2:
                pass  # (B,1) ids reduce over one element — identical to a 1-hot lookup
            w = None if flat_w is None else flat_w[i]
            if w is not None and not isinstance(w, torch.Tensor):
                w = torch.as_tensor(w)
            if w is not None:
                w = w.to(device=ids.device, dtype=torch.float32)
            tcfg = self._tables[self._feature_table[i]]
            feats.append(dict(table=self._table_params[self._feature_table[i]], ids=ids, weights=w,
                              combiner=tcfg.combiner))
        out = ops.gather_concat(feats, sparse_arena=self.sparse_grad_arena)
        if concat:
            return out
        views, off = [], 0
        for i in range(len(feats)):
            e = self._tables[self._feature_table[i]].embedding_dim
            views.append(out[:, off:off + e])
            off += e
        return _pack_like(self.feature_configs, iter(views))

    # ------------------------------------------------------------------ per-table optimizers (:172-186)
    def table_optimizers(self):
        """One optimizer per TableConfig, built from `TableConfig.optimizer` (a name or an instance).  The supported set is
        the reference's (jax/config_conversion.py:211-288): SGD, Adagrad, Adam, Ftrl — in their row-sparse forms, i.e. only
        rows that received gradient are touched (Adam is the per-row "lazy" form the SparseCore path applies)."""
        from .. import optimizers as O_
        if getattr(self, "_table_opts", None) is None:
            opts = []
            for t in self._tables:
                o = t.optimizer
                if isinstance(o, str):
                    name = o.lower()
                    if name not in ("sgd", "adagrad", "adam", "ftrl"):
                        raise ValueError(f"Unsupported optimizer type {o!r}. Optimizer must be one of [Adagrad, Adam, Ftrl, SGD].")
                    o = O_.Adam(sparse_rows=True) if name == "adam" else O_.get(name)
                elif not isinstance(o, (O_.SGD, O_.Adagrad, O_.Adam, O_.Ftrl)) or type(o) is O_.AdamW:
                    raise ValueError(f"Unsupported optimizer type {type(o)}. Optimizer must be one of [Adagrad, Adam, Ftrl, SGD].")
                if isinstance(o, O_.Adam):
                    o.sparse_rows = True
                opts.append(o)
            self._table_opts = opts
        return self._table_opts

    def apply_table_gradients(self) -> None:
        """Applies every table's own optimizer to the gradient its arena received (needs sparse_grad_arena=True: the fused
        backward leaves per-table gradient arenas + touched bitmaps instead of dense (V, E) gradients)."""
        if not self.sparse_grad_arena:
            raise ValueError("apply_table_gradients needs DistributedEmbedding(..., sparse_grad_arena=True)")
        for p, opt in zip(self._table_params, self.table_optimizers()):
            opt.apply([p])

    def get_config(self):
        c = super().get_config()
        # shared TableConfigs are de-duplicated by index (base_distributed_embedding.py:1053-1139)
        c.update(tables=[t.get_config() for t in self._tables],
                 features=[dict(path=list(p), name=fc.name, table=self._feature_table[i],
                                input_shape=tuple(fc.input_shape), output_shape=tuple(fc.output_shape))
                           for i, (p, fc) in enumerate(self._flat)],
                 table_stacking=self.table_stacking)
        return c

    @classmethod
    def from_config(cls, config):
        config = dict(config)
        tables = [TableConfig.from_config(t) for t in config.pop("tables")]
        feats = config.pop("features")
        struct: Any = {}
        flat = []
        for f in feats:
            flat.append((tuple(f["path"]), FeatureConfig(f["name"], tables[f["table"]], tuple(f["input_shape"]),
                                                         tuple(f["output_shape"]))))
        if len(flat) == 1 and flat[0][0] == ():
            struct = flat[0][1]
        else:
            for path, fc in flat:
                cur = struct
                for k in path[:-1]:
                    cur = cur.setdefault(k, {})
                cur[path[-1]] = fc
        return cls(struct, **config)
