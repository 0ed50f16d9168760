"""ShardedDCN — the DCN-v2 step with every embedding table row-sharded (MOD) across the GPUs of one
NVSwitch box (BASELINE configs[4] / SURVEY §8e), dense layers data-parallel.

Exchange design (measured, benchmarks/p2p_probe.py on 2 x B200 over NV18): random 128-byte rows read
straight out of a peer's HBM crawl at ~60 GB/s (the remote path cannot keep enough translations /
requests in flight), while MONOTONIC remote access runs at NVLink speed (660-730 GB/s).  So all random
traffic is kept local and NVLink only carries position-ordered streams:

  forward   1. owners: for every requester s, gather the rows THIS rank owns (id % S == me) from the local
               shard into a staging buffer laid out in the requester's own (b, f) order
               (gather kernel, KRS_SHARD_OWNER — random reads are local; the requester's ids are read
               sequentially over NVLink)
            2. requesters: pull row (b, f) from owner id % S at staging position b*F+f
               (gather kernel, KRS_SHARD_POSITION — remote reads are monotonic), writing the concatenated
               activation directly
  backward  3. requesters: push gradient row (b, f) into the owner's gradient staging at position b*F+f
               with plain 16-byte stores (push_rows_kernel — no atomics cross NVLink)
            4. owners: scatter-add their staged rows into the local gradient arena (local atomics), then
               run the optimizer on the local shard only.
NCCL (torch.distributed) is used where a collective is really needed: the all-reduce of the flat dense
gradient buffer (which also orders step 3 before step 4) and 1-element all-reduces as stream-ordered
barriers between the phases.  Staging and id buffers are cudaMalloc'd and shared with cudaIpc.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from . import _lib as L
from . import initializers, ops
from ._lib import check, lib, ptr
from .dcn import DCN
from .sharding import local_vocab, shard_row_offsets

SHARD_OWNER, SHARD_POSITION = 1, 2


class _Raw:
    """Expose a raw device allocation to torch through __cuda_array_interface__."""

    def __init__(self, p: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(p), False),
                                         "version": 3, "strides": None}


def _ipc_tensor(shape, dtype):
    """cudaMalloc + cudaIpcGetMemHandle; returns (tensor aliasing the allocation, 64-byte handle, pointer)."""
    n = 1
    for s in shape:
        n *= int(s)
    p = C.c_void_p()
    handle = (C.c_ubyte * 64)()
    check(lib.krs_ipc_alloc(C.byref(p), max(n * 4, 256), handle))
    typestr = "<f4" if dtype == torch.float32 else "<i4"
    t = torch.as_tensor(_Raw(p.value, shape, typestr), device="cuda")
    t.zero_()
    return t, bytes(handle), p.value


class ShardedDCN(DCN):
    def __init__(self, vocab_sizes, rank: int, world: int, **kw):
        self.rank_, self.world_ = int(rank), int(world)
        self._opened = []
        self._owned = []
        super().__init__(vocab_sizes, **kw)
        self.rank, self.world = self.rank_, self.world_
        self._tick = torch.zeros((1,), device="cuda")

    # ------------------------------------------------------------------ tables (local shard only)
    def _init_tables(self, seed, embeddings_initializer):
        S, me = self.world_, self.rank_
        self.row_off, self.total_rows = shard_row_offsets(self.vocab_sizes, me, S)
        emb = torch.zeros((self.total_rows, self.E), dtype=torch.float32, device="cuda")
        # identical global tables on every rank (same seed); keep only the local MOD shard
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
        self.emb_grad = torch.zeros_like(emb)
        self.emb_touched = torch.zeros((self.total_rows // 32,), dtype=torch.int32, device="cuda")
        self.emb._krs_arena, self.emb._krs_touched = self.emb_grad, self.emb_touched

    def tables(self):
        return [self.emb[self.row_off[f]:self.row_off[f] + local_vocab(v, self.rank_, self.world_)]
                for f, v in enumerate(self.vocab_sizes)]

    # ------------------------------------------------------------------ per-batch-size shared buffers + plans
    def _make_plan(self, ids):
        return None        # sharded plans are built in _step_buffers (they need the peers' buffers)

    def _open(self, handle: bytes) -> int:
        q = C.c_void_p()
        hb = (C.c_ubyte * 64).from_buffer_copy(handle)
        check(lib.krs_ipc_open(hb, C.byref(q)))
        self._opened.append(q.value)
        return q.value

    def _step_buffers(self, B: int):
        b = self._bufs.get(B)
        if b is not None:
            return b
        b = super()._step_buffers(B)
        S, me, F, E = self.world_, self.rank_, self.F, self.E
        rows = B * F
        ids, h_ids, p_ids = _ipc_tensor((B, F), torch.int32)
        stage, h_st, p_st = _ipc_tensor((S, rows, E), torch.float32)      # [requester][position][E]: rows I own
        gstage, h_gs, p_gs = _ipc_tensor((S, rows, E), torch.float32)     # [requester][position][E]: their grads
        self._owned += [p_ids, p_st, p_gs]
        b["ids"], b["stage"], b["gstage"] = ids, stage, gstage
        allh = [None] * S
        dist.all_gather_object(allh, dict(ids=h_ids, stage=h_st, gstage=h_gs))
        peer = {k: [(p if s == me else self._open(allh[s][k])) for s in range(S)]
                for k, p in (("ids", p_ids), ("stage", p_st), ("gstage", p_gs))}
        region = rows * E * 4
        # device pointer tables: where MY rows live inside every owner's staging buffers
        b["pull_ptrs"] = torch.tensor([peer["stage"][o] + me * region for o in range(S)], dtype=torch.int64, device="cuda")
        b["push_ptrs"] = torch.tensor([peer["gstage"][o] + me * region for o in range(S)], dtype=torch.int64, device="cuda")
        tabs = [self.emb[self.row_off[f]:] for f in range(F)]
        grads = [self.emb_grad[self.row_off[f]:] for f in range(F)]
        touched = [self.emb_touched[self.row_off[f] // 32:] for f in range(F)]

        def base_plan():
            plan = ops.GatherPlan([dict(table=tabs[f], ids=ids[:, f], weights=None, combiner="sum") for f in range(F)])
            for f in range(F):
                d = plan.arr[f]
                d.vocab = self.vocab_sizes[f]          # GLOBAL vocabulary: ids are global rows
                d.num_shards = S
            return plan

        owner_fwd, owner_bwd = [], []
        for s in range(S):                              # s = requesting rank
            pf, pb = base_plan(), base_plan()
            for f in range(F):
                for plan in (pf, pb):
                    d = plan.arr[f]
                    d.ids = peer["ids"][s] + f * 4      # column f of rank s's (B, F) int32 id matrix
                    d.ids_stride = F
                    d.shard_mode = SHARD_OWNER | (me << 8)
                pb.arr[f].grad = grads[f].data_ptr()
                pb.arr[f].touched = touched[f].data_ptr()
            owner_fwd.append(pf)
            owner_bwd.append(pb)
        pull, push = base_plan(), base_plan()
        for f in range(F):
            pull.arr[f].table = None
            pull.arr[f].shard_mode = SHARD_POSITION | (me << 8)
            pull.arr[f].shard_tables = b["pull_ptrs"].data_ptr()
            push.arr[f].shard_mode = SHARD_POSITION | (me << 8)
            push.arr[f].shard_grads = b["push_ptrs"].data_ptr()
        b.update(owner_fwd=owner_fwd, owner_bwd=owner_bwd, pull=pull, push=push)
        dist.barrier()
        return b

    def _barrier(self):
        dist.all_reduce(self._tick)        # stream-ordered: completes when every rank reached it

    # ------------------------------------------------------------------ exchange phases
    def _gather_into(self, b, B, s):
        S, F, D = self.world_, self.F, self.D
        self._barrier()                                               # every rank's ids are staged
        for r in range(S):                                            # 1. owner side (local random reads)
            p = b["owner_fwd"][r]
            check(lib.krs_gather_fwd(p.arr, F, B, b["stage"][r].data_ptr(), D, 0, s))
        self._barrier()                                               # all owners staged their rows
        p = b["pull"]                                                 # 2. monotonic pull over NVLink
        check(lib.krs_gather_fwd(p.arr, F, B, ptr(b["xs"][0]), D, 0, s))

    def _scatter_from(self, b, B, cur, s):
        S, F, D = self.world_, self.F, self.D
        p = b["push"]                                                 # 3. monotonic push (plain stores)
        check(lib.krs_gather_bwd(p.arr, F, B, ptr(cur), D, s))
        dist.all_reduce(self.dense_grad_flat)                         # real collective; also orders 3 before 4
        for r in range(S):                                            # 4. owner side (local atomics)
            pb = b["owner_bwd"][r]
            check(lib.krs_gather_bwd(pb.arr, F, B, b["gstage"][r].data_ptr(), D, s))

    def _end_of_step(self):
        self._barrier()                                               # peers are done with my ids / staging

    def _sync_gradients(self):
        """The dense all-reduce already happened inside _scatter_from."""

    def forward(self, ids, sparse_arena: bool = False):
        ids = self._as_ids(ids)
        B = ids.shape[0]
        b = self._step_buffers(B)
        b["ids"].copy_(ids)
        with torch.no_grad():
            self._gather_into(b, B, L.stream())
            self._barrier()
        x0 = b["xs"][0]
        xl = x0
        for i, c in enumerate(self.cross):
            xl = c(x0) if i == 0 else c(x0, xl)
        h = xl
        for d in self.mlp:
            h = d(h)
        return h

    def close(self):
        torch.cuda.synchronize()
        for p in self._opened:
            lib.krs_ipc_close(C.c_void_p(p))
        self._opened = []
