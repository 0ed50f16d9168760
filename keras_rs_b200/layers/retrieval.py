"""Retrieval / BruteForceRetrieval — drop-ins for keras_rs.layers.Retrieval
(keras_rs/src/layers/retrieval/retrieval.py:11-127) and BruteForceRetrieval
(brute_force_retrieval.py:12-148).  Validation messages follow retrieval.py:42-68 (regexes pinned by
retrieval_test.py:21-40).  Scoring + top-k is one streaming kernel (csrc/topk.cu): the
(num_queries, num_candidates) score matrix is never materialised."""
from __future__ import annotations

import abc
from typing import Any

import torch

from .. import _lib as L
from .. import ops
from .base import Layer, register


@register("keras_rs.layers.Retrieval")
class Retrieval(Layer, abc.ABC):
    def __init__(self, k: int = 10, return_scores: bool = True, **kwargs: Any) -> None:
        super().__init__(**kwargs)
        self.k = k
        self.return_scores = return_scores

    def _validate_candidate_embeddings_and_ids(self, candidate_embeddings, candidate_ids=None) -> None:
        if candidate_embeddings is None:
            raise ValueError("`candidate_embeddings` is required.")
        if len(candidate_embeddings.shape) != 2:
            raise ValueError("`candidate_embeddings` must be a tensor of rank 2 (num_candidates, embedding_size), "
                             f"received `candidate_embeddings` with shape {tuple(candidate_embeddings.shape)}")
        if candidate_embeddings.shape[0] < self.k:
            raise ValueError(f"The number of candidates provided ({candidate_embeddings.shape[0]}) is less than the "
                             f"number of candidates to retrieve (k={self.k}).")
        if candidate_ids is not None and candidate_ids.shape[0] != candidate_embeddings.shape[0]:
            raise ValueError("The `candidate_embeddings` and `candidate_is` tensors must have the same number of "
                             f"rows, got tensors of shape {tuple(candidate_embeddings.shape)} and "
                             f"{tuple(candidate_ids.shape)}.")

    @abc.abstractmethod
    def update_candidates(self, candidate_embeddings, candidate_ids=None) -> None:
        pass

    @abc.abstractmethod
    def call(self, inputs):
        pass

    def compute_score(self, query_embedding: torch.Tensor, candidate_embedding: torch.Tensor) -> torch.Tensor:
        """matmul(query, transpose(candidates)) (retrieval.py:101-117) — materialises the scores; the
        retrieval layers below never call this on their hot path."""
        return ops.sgemm(query_embedding.contiguous(), candidate_embedding.contiguous(), transB=True)

    def get_config(self) -> dict[str, Any]:
        c = super().get_config()
        # the reference stores the bound method under "return_scores" (retrieval.py:119-127, a bug);
        # we store the boolean it evidently meant.
        c.update(k=self.k, return_scores=self.return_scores)
        return c


@register("keras_rs.layers.BruteForceRetrieval")
class BruteForceRetrieval(Retrieval):
    def __init__(self, candidate_embeddings=None, candidate_ids=None, k: int = 10, return_scores: bool = True,
                 **kwargs: Any) -> None:
        super().__init__(k=k, return_scores=return_scores, **kwargs)
        self.candidate_embeddings = None
        self.candidate_ids = None
        if candidate_embeddings is None:
            if candidate_ids is not None:                              # brute_force_retrieval.py:72-77
                raise ValueError("You cannot provide `candidate_ids` without providing `candidate_embeddings`")
        else:
            self.update_candidates(candidate_embeddings, candidate_ids)

    def _to_dev(self, t, dtype):
        t = torch.as_tensor(t)
        return t.detach().to(device=self._device, dtype=dtype).contiguous()

    def update_candidates(self, candidate_embeddings, candidate_ids=None) -> None:
        self._validate_candidate_embeddings_and_ids(candidate_embeddings, candidate_ids)
        if self.candidate_embeddings is not None:                      # :97-109 update in place
            with torch.no_grad():
                self.candidate_embeddings.copy_(self._to_dev(candidate_embeddings, torch.float32))
            if self.candidate_ids is None:
                if candidate_ids is not None:
                    raise ValueError("New `candidate_ids` cannot be provided as previous candidates did not have "
                                     "candidate IDs")
            elif candidate_ids is not None:
                with torch.no_grad():
                    self.candidate_ids.copy_(self._to_dev(candidate_ids, torch.int32))
        else:                                                          # :110-123 creation (non-trainable)
            self.candidate_embeddings = torch.nn.Parameter(self._to_dev(candidate_embeddings, torch.float32),
                                                           requires_grad=False)
            if candidate_ids is not None:
                self.candidate_ids = torch.nn.Parameter(self._to_dev(candidate_ids, torch.int32), requires_grad=False)
        self.built = True

    def build(self, *a):
        self.built = True

    @property
    def weights(self):
        return [w for w in (self.candidate_embeddings, self.candidate_ids) if w is not None]

    def call(self, inputs: torch.Tensor):
        if self.candidate_embeddings is None:
            raise ValueError("`candidate_embeddings` is required.")
        cand = self.candidate_embeddings
        if isinstance(cand, torch.Tensor) and cand.requires_grad:
            cand = cand.detach()                                       # shared Embedding variable (basic_retrieval.py:249-257)
        L.require_cuda(inputs, "inputs")
        L.require_cuda(cand, "candidate_embeddings")
        self._validate_candidate_embeddings_and_ids(cand, self.candidate_ids)
        q = inputs.detach()
        lead = q.shape[:-1]
        q2 = q.reshape(-1, q.shape[-1])
        ids = None if self.candidate_ids is None else self.candidate_ids.detach().to(torch.int32)
        top_scores, top_ids = ops.top_k_scores(q2, cand, ids, self.k)
        top_scores = top_scores.reshape(*lead, self.k)
        top_ids = top_ids.reshape(*lead, self.k)
        if self.return_scores:
            return top_scores, top_ids
        return top_ids

    def compute_output_shape(self, input_shape):
        s = tuple(input_shape[:-1]) + (self.k,)
        return (s, s) if self.return_scores else s


def merge_top_k(scores: torch.Tensor, ids: torch.Tensor, k: int):
    """Exact top-k of the concatenated per-shard result lists: scores, ids (nq, m) -> (nq, k), scores descending, ties ->
    lowest position first (csrc/rowops.cu krs_row_topk).  With shards that hold contiguous, increasing id ranges and lists
    concatenated in shard order, position order IS id order, so the merged result equals the unsharded top-k."""
    from .._lib import check, lib, ptr, stream
    L.require_cuda(scores, "scores")
    L.require_cuda(ids, "ids", dtype=torch.int32)
    scores, ids = scores.contiguous(), ids.contiguous()
    nq, m = scores.shape
    out_s = torch.empty((nq, k), device=scores.device, dtype=torch.float32)
    out_pos = torch.empty((nq, k), device=scores.device, dtype=torch.int32)
    out_i = torch.empty((nq, k), device=scores.device, dtype=torch.int32)
    check(lib.krs_row_topk(ptr(scores), nq, m, m, None, 0, 0.0, k, ptr(out_s), ptr(out_pos), None, 0, None, ptr(ids), m, ptr(out_i),
                           stream()))
    return out_s, out_i


class CandidateShardedRetrieval(torch.nn.Module):
    """BruteForceRetrieval over candidates split row-wise across the GPUs of one box (the multi-GPU caller in
    the reference is examples/data_parallel_retrieval.py:145-165, which replicates the candidates — at C4 size, 1e7 x 64
    floats, every GPU would stream the same 2.56 GB per query batch; sharded, each GPU streams 1/S of it).

    Every rank holds the contiguous slice [first_id, first_id + n_local) of the candidate matrix and receives the SAME query
    batch.  Per call: local exact top-k on the tensor-pipe scorer (csrc/topk.cu) with global candidate ids, all-gather of
    the (nq, k) lists (NCCL: a real collective, S * nq * k * 8 bytes), one krs_row_topk merge.  `group=None` with
    world_size 1 (or no process group) degenerates to the local search."""

    def __init__(self, local_candidates: torch.Tensor, first_id: int, k: int = 10, candidate_ids: torch.Tensor | None = None,
                 return_scores: bool = True, group=None):
        super().__init__()
        L.require_cuda(local_candidates, "local_candidates")
        self.k, self.return_scores, self.group = int(k), return_scores, group
        self.candidates = local_candidates.detach().contiguous()
        n = self.candidates.shape[0]
        if candidate_ids is None:
            candidate_ids = torch.arange(first_id, first_id + n, device=self.candidates.device, dtype=torch.int32)
        self.ids = candidate_ids.to(torch.int32).contiguous()

    def local_top_k(self, q: torch.Tensor):
        k = min(self.k, self.candidates.shape[0])
        return ops.top_k_scores(q.detach().contiguous(), self.candidates, self.ids, k)

    def forward(self, q: torch.Tensor):
        import torch.distributed as dist
        s, i = self.local_top_k(q)
        world = dist.get_world_size(self.group) if (dist.is_available() and dist.is_initialized()) else 1
        if world > 1:
            if s.shape[1] != self.k:
                raise ValueError("CandidateShardedRetrieval: every shard must hold at least k candidates")
            gs = torch.empty((world,) + tuple(s.shape), device=s.device, dtype=s.dtype)
            gi = torch.empty((world,) + tuple(i.shape), device=i.device, dtype=i.dtype)
            dist.all_gather_into_tensor(gs, s, group=self.group)
            dist.all_gather_into_tensor(gi, i, group=self.group)
            s, i = merge_top_k(gs.permute(1, 0, 2).reshape(s.shape[0], -1), gi.permute(1, 0, 2).reshape(i.shape[0], -1), self.k)
        return (s, i) if self.return_scores else i
