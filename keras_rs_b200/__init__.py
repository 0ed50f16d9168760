"""keras_rs_b200 — B200-native (sm_100a) implementation of the keras-rs hot path
Embedding gather -> FeatureCross | DotInteraction -> Dense stack (+ BruteForceRetrieval).

Public surface mirrors `keras_rs`: `keras_rs_b200.layers.{FeatureCross, DotInteraction, Retrieval,
BruteForceRetrieval, EmbedReduce, DistributedEmbedding, TableConfig, FeatureConfig}`.
Importing this package loads libkrs_b200.so and FAILS if it is missing — there is no CPU fallback.
"""
from . import _lib  # noqa: F401  (raises ImportError when the CUDA extension is not built)
from . import initializers, layers, ops, optimizers  # noqa: F401
from .ops import get_gemm_engine, set_gemm_engine  # noqa: F401

import os as _os

# dense contractions: tensor pipe (tcgen05, 3xTF32 split, fp32-level accuracy; "tcgen05_ts" keeps the A operand in
# tensor memory, "tcgen05" reads both operands from shared memory) unless KRS_GEMM_ENGINE=ffma asks for the exact-fp32
# FMA engine.  Shapes the tensor-core kernel does not take fall through to FFMA automatically.
set_gemm_engine(_os.environ.get("KRS_GEMM_ENGINE", "tcgen05_ts"))

__version__ = "0.1.0"


def version() -> str:
    return __version__
