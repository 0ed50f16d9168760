"""Host-side bookkeeping for row-wise MOD table sharding (no GPU needed; covered by the world_size-2
gloo tests).  Layout precedent: the reference's TPU SparseCore path — row r of a table lives on shard
r % S at local row r // S (keras_rs/src/layers/embedding/jax/embedding_utils.py:187-197
sharding_strategy="MOD"; tensorflow/distributed_embedding.py:316-328 mod-shard reassembly)."""
from __future__ import annotations

from typing import Sequence


def round_up(a: int, b: int) -> int:
    return (a + b - 1) // b * b


def local_vocab(vocab: int, shard: int, num_shards: int) -> int:
    """Rows of a `vocab`-row table owned by `shard` (rows shard, shard+S, shard+2S, ...)."""
    if shard >= vocab:
        return 0
    return (vocab - shard + num_shards - 1) // num_shards


def shard_row_offsets(vocab_sizes: Sequence[int], shard: int, num_shards: int, align: int = 32):
    """Row offset of every table inside one shard's arena (each table padded to `align` rows so its
    slice of the touched bitmap is word aligned) and the arena's total rows.  Offsets depend on the
    shard only through local_vocab, so every rank can compute every peer's layout."""
    offs, off = [], 0
    for v in vocab_sizes:
        offs.append(off)
        off += round_up(max(local_vocab(v, shard, num_shards), 1), align)
    return offs, off


def owner_and_local(row: int, num_shards: int):
    return row % num_shards, row // num_shards


def global_row(shard: int, local: int, num_shards: int) -> int:
    return local * num_shards + shard
