"""ShardedDCN — the DCN-v2 step with every embedding table row-sharded (MOD) across the GPUs of one
NVSwitch box (BASELINE configs[4] / SURVEY §8e), dense layers data-parallel.

B200-first exchange: there is NO id all-to-all and no staged row all-to-all.  Every rank exports its
table / gradient / bitmap arenas with cudaIpc; the fused gather kernel addresses
`peer_arena[id % S] + (id // S) * E` directly, so each remote row crosses NVLink exactly once inside
the kernel that writes the concatenated activation, and the backward scatter-add pushes row
gradients to their owners with remote atomics in one kernel as well.  NCCL (through
torch.distributed) is used only where a collective is really needed: the all-reduce of the ~9 MB of
dense gradients (which also orders "all scatters done" before the owners' optimizer sweeps) and a
1-element all-reduce as the stream-ordered barrier between an optimizer sweep and the next step's
remote reads.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from . import _lib as L
from . import initializers, ops
from ._lib import check, lib
from .dcn import DCN
from .sharding import local_vocab, shard_row_offsets


class _Raw:
    """Expose a raw device allocation to torch through __cuda_array_interface__."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 3, "strides": None}


def _ipc_tensor(shape, dtype):
    """cudaMalloc + cudaIpcGetMemHandle; returns (tensor aliasing the allocation, 64-byte handle)."""
    n = 1
    for s in shape:
        n *= int(s)
    itemsize = 4
    p = C.c_void_p()
    handle = (C.c_ubyte * 64)()
    check(lib.krs_ipc_alloc(C.byref(p), max(n * itemsize, 256), handle))
    typestr = "<f4" if dtype == torch.float32 else "<i4"
    t = torch.as_tensor(_Raw(p.value, shape, typestr), device="cuda")
    t.zero_()
    return t, bytes(handle), p.value


class ShardedDCN(DCN):
    def __init__(self, vocab_sizes, rank: int, world: int, **kw):
        self.rank_, self.world_ = int(rank), int(world)
        super().__init__(vocab_sizes, **kw)
        self.rank, self.world = self.rank_, self.world_

    # ------------------------------------------------------------------ tables
    def _init_tables(self, seed, embeddings_initializer):
        S, me = self.world_, self.rank_
        self.row_off, self.total_rows = shard_row_offsets(self.vocab_sizes, me, S)
        self.peer_row_off = [shard_row_offsets(self.vocab_sizes, s, S)[0] for s in range(S)]
        emb, h_emb, p_emb = _ipc_tensor((self.total_rows, self.E), torch.float32)
        grad, h_grad, p_grad = _ipc_tensor((self.total_rows, self.E), torch.float32)
        touched, h_t, p_t = _ipc_tensor((self.total_rows // 32,), torch.int32)
        # identical global tables on every rank (same seed), keep only the local MOD shard
        g = torch.Generator(device="cuda").manual_seed(seed)
        init = initializers.get(embeddings_initializer)
        for f, v in enumerate(self.vocab_sizes):
            if isinstance(init, initializers.RandomUniform):
                full = torch.rand((v, self.E), device="cuda", generator=g) * (init.maxval - init.minval) + init.minval
            else:
                full = init((v, self.E)).cuda()
            lv = local_vocab(v, me, S)
            emb[self.row_off[f]:self.row_off[f] + lv] = full[me::S]
            del full
        self.emb = torch.nn.Parameter(emb)
        self.emb_grad, self.emb_touched = grad, touched
        self.emb._krs_arena, self.emb._krs_touched = grad, touched
        # exchange handles, open the peers' arenas
        mine = dict(emb=h_emb, grad=h_grad, touched=h_t)
        allh = [None] * S
        dist.all_gather_object(allh, mine)
        self._peer_ptrs = {"emb": [], "grad": [], "touched": []}
        self._opened = []
        for s in range(S):
            for key, local_ptr in (("emb", p_emb), ("grad", p_grad), ("touched", p_t)):
                if s == me:
                    self._peer_ptrs[key].append(local_ptr)
                else:
                    q = C.c_void_p()
                    hb = (C.c_ubyte * 64).from_buffer_copy(allh[s][key])
                    check(lib.krs_ipc_open(hb, C.byref(q)))
                    self._peer_ptrs[key].append(q.value)
                    self._opened.append(q.value)
        # device arrays of per-feature shard pointers: [F][S]
        F = self.F
        mk = lambda key, scale, div: torch.tensor(
            [[self._peer_ptrs[key][s] + (self.peer_row_off[s][f] // div) * scale for s in range(S)] for f in range(F)],
            dtype=torch.int64, device="cuda")
        self._shard_tables = mk("emb", self.E * 4, 1)
        self._shard_grads = mk("grad", self.E * 4, 1)
        self._shard_touched = mk("touched", 4, 32)
        dist.barrier()

    def tables(self):
        return [self.emb[self.row_off[f]:self.row_off[f] + local_vocab(v, self.rank_, self.world_)]
                for f, v in enumerate(self.vocab_sizes)]

    def _make_plan(self, ids: torch.Tensor):
        S = self.world_
        dummy = self.emb[:1]
        feats = [dict(table=dummy, ids=ids[:, f], weights=None, combiner="sum") for f in range(self.F)]
        plan = ops.GatherPlan(feats)
        off = 0
        for f, v in enumerate(self.vocab_sizes):
            d = plan.arr[f]
            d.table = None
            d.vocab = v                      # GLOBAL vocabulary: ids are global rows
            d.dim = self.E
            d.out_offset = off
            d.num_shards = S
            d.shard_tables = self._shard_tables[f].data_ptr()
            d.shard_grads = self._shard_grads[f].data_ptr()
            d.shard_touched = self._shard_touched[f].data_ptr()
            off += self.E
        plan.out_dim = off
        return plan

    def _feature_list(self, ids):
        raise RuntimeError("ShardedDCN builds its gather plan with _make_plan")

    def forward(self, ids, sparse_arena: bool = False):
        ids = self._as_ids(ids)
        plan = self._make_plan(ids.contiguous())
        x0 = plan.forward()
        xl = x0
        for i, c in enumerate(self.cross):
            xl = c(x0) if i == 0 else c(x0, xl)
        h = xl
        for d in self.mlp:
            h = d(h)
        return h

    # ------------------------------------------------------------------ collectives
    def _sync_gradients(self):
        # all-reduce of the dense gradients; being stream ordered after this rank's scatter kernel it
        # also guarantees every rank's remote atomics have landed before any owner's optimizer sweep
        dist.all_reduce(self.dense_grad_flat)

    def train_on_batch(self, ids, labels, optimizer, denom: int = 0):
        loss = super().train_on_batch(ids, labels, optimizer, denom)
        # owners have updated their shards: order that before the next step's remote reads
        if not hasattr(self, "_tick"):
            self._tick = torch.zeros((1,), device="cuda")
        dist.all_reduce(self._tick)
        return loss

    def close(self):
        for p in self._opened:
            lib.krs_ipc_close(C.c_void_p(p))
        self._opened = []
