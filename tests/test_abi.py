"""CPU: the C-ABI library loads and exports every symbol include/krs_b200.h declares; host-side
validation logic that needs no GPU."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "krs_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(krs_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_expected_entry_points():
    syms = header_symbols()
    for s in ["krs_gather_fwd", "krs_gather_bwd", "krs_cross_fwd", "krs_cross_bwd", "krs_dot_fwd", "krs_dot_bwd",
              "krs_dense_fwd", "krs_dense_bwd", "krs_topk", "krs_adamw", "krs_sgd_adagrad", "krs_mod_route"]:
        assert s in syms


def test_library_exports_every_declared_symbol():
    from keras_rs_b200 import _lib
    main = ctypes.CDLL(_lib.LIB_PATH)
    for s in header_symbols():
        assert hasattr(main, s), f"{s} declared in krs_b200.h but not exported by libkrs_b200.so"
    assert sorted(_lib.EXPORTED) == sorted(header_symbols())


def test_version_and_error_string():
    from keras_rs_b200 import _lib
    assert _lib.lib.krs_version() >= 100
    assert isinstance(_lib.lib.krs_last_error(), bytes)
    assert _lib.lib.krs_set_gemm_engine(7) != 0
    assert b"engine" in _lib.lib.krs_last_error()


def test_cpu_tensors_are_rejected_loudly():
    import torch
    from keras_rs_b200 import _lib
    from keras_rs_b200.layers import DotInteraction
    with pytest.raises(_lib.KrsError):
        DotInteraction()([torch.ones(2, 3), torch.ones(2, 3)])


def test_layer_validation_without_gpu():
    import torch
    from keras_rs_b200.layers import BruteForceRetrieval, DotInteraction, EmbedReduce, FeatureCross, Retrieval
    with pytest.raises(ValueError):
        FeatureCross(diag_scale=-1.0)                                   # feature_cross_test.py:62-65
    with pytest.raises(ValueError):
        DotInteraction()([torch.ones(3), torch.ones(3)])                # dot_interaction_test.py:93-98
    with pytest.raises(ValueError):
        DotInteraction()([torch.ones(1, 3), torch.ones(1, 4)])          # :100-105
    with pytest.raises(ValueError):
        BruteForceRetrieval(candidate_ids=torch.arange(3))              # brute_force_retrieval.py:72-77

    class NoCall(Retrieval):
        def update_candidates(self, candidate_embeddings, candidate_ids=None):
            pass

    class NoUpdate(Retrieval):
        def call(self, inputs):
            pass

    with pytest.raises(TypeError):                                      # retrieval_test.py:50-66
        NoCall(k=5)
    with pytest.raises(TypeError):
        NoUpdate(k=5)

    class Dummy(Retrieval):
        def update_candidates(self, candidate_embeddings, candidate_ids=None):
            pass

        def call(self, inputs):
            pass

    import json
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "retrieval.json")))
    layer = Dummy(k=5)
    for e in g["errors"]:
        emb = None if e["emb_shape"] is None else torch.zeros(e["emb_shape"])
        ids = None if e["ids_shape"] is None else torch.zeros(e["ids_shape"], dtype=torch.int32)
        with pytest.raises(ValueError, match=e["regex"]):
            layer._validate_candidate_embeddings_and_ids(emb, ids)


def test_serialization_roundtrip_cpu():
    from keras_rs_b200.layers import DotInteraction, FeatureCross, deserialize, serialize
    a = FeatureCross(projection_dim=None, pre_activation="swish")       # feature_cross_test.py:90-93
    b = deserialize(serialize(a))
    ca, cb = a.get_config(), b.get_config()
    ca.pop("name"); cb.pop("name")
    assert ca == cb
    assert serialize(a)["registered_name"] == "keras_rs>FeatureCross"
    d = DotInteraction(self_interaction=True, skip_gather=True)
    assert deserialize(serialize(d)).get_config()["skip_gather"] is True


def test_distributed_embedding_config_cpu():
    from keras_rs_b200.layers import DistributedEmbedding, FeatureConfig, TableConfig
    t = TableConfig("t", 10, 4)
    assert (t.combiner, t.placement, t.optimizer) == ("mean", "auto", "adam")   # distributed_embedding_config.py:54-61
    with pytest.raises(ValueError, match="sparsecore"):                  # distributed_embedding_test.py:195-198
        DistributedEmbedding({"f": FeatureConfig("f", TableConfig("s", 10, 4, placement="sparsecore"), (8,), (8, 4))})
    fc = {"a": FeatureConfig("a", t, (8,), (8, 4)), "b": FeatureConfig("b", t, (8,), (8, 4))}
    layer = DistributedEmbedding(fc)
    assert len(layer._tables) == 1                                        # shared table -> one variable (:640-652)
    cfg = layer.get_config()
    assert len(cfg["tables"]) == 1 and [f["table"] for f in cfg["features"]] == [0, 0]
