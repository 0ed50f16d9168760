"""keras_rs_b200.layers — same public names as keras_rs.layers (keras_rs/api/layers/__init__.py:7-36)
for the hot path, plus the Keras core layers that path composes (Dense, Embedding)."""
from .base import Layer, deserialize, serialize
from .dense import Dense
from .distributed_embedding import DistributedEmbedding, FeatureConfig, TableConfig
from .dot_interaction import DotInteraction
from .embedding import EmbedReduce, Embedding
from .feature_cross import FeatureCross
from .retrieval import BruteForceRetrieval, CandidateShardedRetrieval, Retrieval, merge_top_k
from .retrieval_helpers import HardNegativeMining, RemoveAccidentalHits, SamplingProbabilityCorrection

__all__ = ["Layer", "Dense", "Embedding", "EmbedReduce", "DistributedEmbedding", "TableConfig", "FeatureConfig",
           "FeatureCross", "DotInteraction", "Retrieval", "BruteForceRetrieval", "CandidateShardedRetrieval", "merge_top_k", "HardNegativeMining", "RemoveAccidentalHits",
           "SamplingProbabilityCorrection", "serialize", "deserialize"]
