"""ShardedDCN — the DCN-v2 step with every embedding table row-sharded (MOD) across the GPUs of one
NVSwitch box (BASELINE configs[4] / SURVEY §8e), dense layers data-parallel.

The exchange is the reference's SparseCore protocol (ids routed to the owner, owner-side lookup,
activations returned, optimizer applied by the owner without a dense gradient:
jax/embedding_utils.py:144-217, jax/embedding_lookup.py:174-273) rebuilt on NVLink peer memory
(csrc/exchange.cu).  Measured fact that shapes it (benchmarks/p2p_probe.py): random 128-byte rows
out of a peer's HBM crawl at ~60 GB/s, monotonic remote access runs at NVLink speed, so random
traffic stays local and NVLink only carries position-ordered streams.  Per step and rank:

  route        local ids (B,F) -> per-owner request lists (owner arena row, position), stable sort
  barrier      flags in peer memory (no NCCL)
  gather_push  ONE launch: every requester's bucket is served from the local shard and the rows
               are written straight into the requester's activation (remote writes, increasing)
  barrier
  dense step   cross stack + MLP + loss + backward, local
  all-reduce   flat dense-gradient buffer (NCCL, asynchronous: overlaps the next three lines)
  barrier
  grad_pull    ONE launch: gradient rows pulled from the requesters' dL/dx0 (remote reads,
               increasing), duplicates combined, accumulated into COMPACT rows (one per distinct
               touched table row; slot = rank of the row's bit in the touched bitmap)
  optimizer    on the compact rows (row-sparse SGD / Adagrad / Adam / FTRL, or the dense-semantics
               AdamW sweep of the local shard reading its gradients from the compact rows)

Nothing table-sized exists besides the table and the optimizer's slot variables, so the stated C5
size (1e9 rows x 128 floats over 8 GPUs = 64 GB of table per GPU) fits 180 GB with Adagrad.
Request lists are double-buffered by step parity, which makes the three barriers sufficient
(a rank passing the first barrier of step t+1 proves every peer finished reading step t's lists,
activation gradients and activations).

`SimGroup` runs S such ranks inside ONE process on one GPU (regions are ordinary allocations,
barriers become stream order): that is how the protocol is checked against the oracle on the
driver's single-GPU test box at S = 2..8.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from . import _lib as L
from . import initializers
from ._lib import KrsXchg, XCHG_MAX_SHARDS, check, lib, ptr, stream
from .dcn import DCN
from .sharding import local_vocab, shard_row_offsets

_FULL_INIT_LIMIT = 1 << 28      # tables up to this many floats are initialised as the global table then sliced


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
    itemsize = torch.empty((), dtype=dtype).element_size()
    p = C.c_void_p()
    handle = (C.c_ubyte * 64)()
    check(lib.krs_ipc_alloc(C.byref(p), max(n * itemsize, 256), handle))
    typestr = {torch.float32: "<f4", torch.int32: "<i4", torch.uint8: "|u1"}[dtype]
    t = torch.as_tensor(_Raw(p.value, shape, typestr), device="cuda")
    t.zero_()
    return t, bytes(handle), p.value


def _align(n: int, a: int = 256) -> int:
    return (n + a - 1) // a * a


class RegionLayout:
    """Byte layout of one rank's exchange region (identical on every rank; include/krs_b200.h krs_xchg_t)."""

    def __init__(self, B: int, F: int, E: int):
        P = B * F
        off = 0
        self.off_flags = off; off += _align((XCHG_MAX_SHARDS + 1) * 4)
        self.off_hdr = off; off += _align(2 * (XCHG_MAX_SHARDS + 1) * 4)
        self.off_rows = off; off += _align(2 * P * 4)
        self.off_pos = off; off += _align(2 * P * 4)
        self.off_x0 = off; off += _align(P * E * 4)
        self.off_grad = off; off += _align(P * E * 4)
        self.nbytes = off
        self.P, self.E, self.B, self.F = P, E, B, F

    def view(self, region: torch.Tensor, off: int, shape, dtype):
        n = 1
        for s in shape:
            n *= int(s)
        return region[off:off + n * torch.empty((), dtype=dtype).element_size()].view(dtype).view(shape)


class CompactGrads:
    """Per-step compact gradient staging of one table shard (what replaces a table-sized gradient arena)."""

    def __init__(self, total_rows: int, E: int, cap_rows: int, device):
        self.nwords = (total_rows + 31) // 32
        self.touched = torch.zeros((self.nwords,), dtype=torch.int32, device=device)
        self.wordprefix = torch.zeros((self.nwords,), dtype=torch.int32, device=device)
        self.blockbase = torch.zeros((max(int(lib.krs_slot_scan_blocks(total_rows)), 1),), dtype=torch.int32, device=device)
        self.n_unique = torch.zeros((1,), dtype=torch.int32, device=device)
        self.E = E
        self.cap_rows = 0
        self.compact = self.uniq_rows = None
        self.reserve(cap_rows, device)

    def reserve(self, cap_rows: int, device):
        if cap_rows > self.cap_rows:
            self.cap_rows = int(cap_rows)
            self.compact = torch.zeros((self.cap_rows, self.E), dtype=torch.float32, device=device)   # persistently zero
            self.uniq_rows = torch.zeros((self.cap_rows,), dtype=torch.int32, device=device)

    def scan(self):
        check(lib.krs_slot_scan(ptr(self.touched), self.nwords, ptr(self.wordprefix), ptr(self.blockbase), ptr(self.n_unique),
                                stream()))


class ShardedDCN(DCN):
    """rank / world: this process's shard and the number of shards.  sim_peers=True defers the peer wiring to SimGroup
    (single-process simulation); otherwise regions are exchanged through torch.distributed (NCCL) + cudaIpc."""

    def __init__(self, vocab_sizes, rank: int, world: int, sim: bool = False, barrier_timeout_s: float = 20.0, **kw):
        self.rank_, self.world_ = int(rank), int(world)
        if not 1 <= self.world_ <= XCHG_MAX_SHARDS:
            raise ValueError(f"ShardedDCN supports 1..{XCHG_MAX_SHARDS} shards, got {world}")
        E = int(kw.get("embedding_dim", 32))
        if E < 4 or E % 4 != 0:
            raise ValueError(f"ShardedDCN needs embedding_dim to be a multiple of 4 (16-byte rows), got {E}")
        self._sim = bool(sim)
        self._timeout = float(barrier_timeout_s)
        self._opened = []
        self._owned = []
        super().__init__(vocab_sizes, **kw)
        self.rank, self.world = self.rank_, self.world_
        self._parity = 0
        self._ar = None
        self._sim_group = None

    # ------------------------------------------------------------------ tables (local shard only)
    def _init_tables(self, seed, embeddings_initializer):
        S, me = self.world_, self.rank_
        self.row_off, self.total_rows = shard_row_offsets(self.vocab_sizes, me, S)
        if self.total_rows >= 2 ** 31:
            raise ValueError("a shard's arena must stay below 2^31 rows")
        emb = torch.zeros((self.total_rows, self.E), dtype=torch.float32, device="cuda")
        g = torch.Generator(device="cuda").manual_seed(seed)
        init = initializers.get(embeddings_initializer)
        for f, v in enumerate(self.vocab_sizes):
            lv = local_vocab(v, me, S)
            dst = emb[self.row_off[f]:self.row_off[f] + lv]
            if v * self.E <= _FULL_INIT_LIMIT:
                # identical global tables for every world size (same seed): build the table, keep the MOD shard
                if isinstance(init, initializers.RandomUniform):
                    full = torch.rand((v, self.E), device="cuda", generator=g) * (init.maxval - init.minval) + init.minval
                else:
                    full = init((v, self.E)).cuda()
                dst.copy_(full[me::S])
                del full
            else:
                # large tables: the shard is drawn directly (per-rank stream), 64 M elements at a time
                gs = torch.Generator(device="cuda").manual_seed(seed * 1000003 + f * 101 + me)
                lo, hi = (init.minval, init.maxval) if isinstance(init, initializers.RandomUniform) else (-0.05, 0.05)
                step = max((1 << 26) // self.E, 1)
                for r0 in range(0, lv, step):
                    blk = dst[r0:r0 + step]
                    blk.uniform_(lo, hi, generator=gs)
        self.emb = torch.nn.Parameter(emb)
        self.emb_grad = None                      # no table-sized gradient buffer in the sharded model
        self.emb_touched = None
        self.cg = CompactGrads(self.total_rows, self.E, 1, "cuda")
        # rows that have ever received a gradient (AdamW sweeps the others with a decay-only update, krs_adamw_cold)
        self.emb_ever = torch.zeros_like(self.cg.touched)
        self.emb._krs_ever = self.emb_ever

    def tables(self):
        return [self.emb[self.row_off[f]:self.row_off[f] + local_vocab(v, self.rank_, self.world_)]
                for f, v in enumerate(self.vocab_sizes)]

    # ------------------------------------------------------------------ per-batch-size region + peer wiring
    def _make_plan(self, ids):
        return None

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
        lay = RegionLayout(B, F, E)
        if self._sim:
            region = torch.zeros((lay.nbytes,), dtype=torch.uint8, device="cuda")
            handle, base = None, region.data_ptr()
        else:
            region, handle, base = _ipc_tensor((lay.nbytes,), torch.uint8)
            self._owned.append(base)
        b["region"], b["layout"], b["region_base"], b["region_handle"] = region, lay, base, handle
        # the activation the owners write into and the gradient they read from live inside the region
        b["xs"][0] = lay.view(region, lay.off_x0, (B, self.D), torch.float32)
        final = "ga" if self.L % 2 == 0 else "gb"          # the buffer _dense_step returns (dcn.py)
        b[final] = lay.view(region, lay.off_grad, (B, self.D), torch.float32)
        b["grad_name"] = final
        b["flags"] = lay.view(region, lay.off_flags, (XCHG_MAX_SHARDS + 1,), torch.int32)
        b["hdr"] = lay.view(region, lay.off_hdr, (2, XCHG_MAX_SHARDS + 1), torch.int32)
        b["req_rows"] = lay.view(region, lay.off_rows, (2, B * F), torch.int32)
        b["req_pos"] = lay.view(region, lay.off_pos, (2, B * F), torch.int32)
        b["route_ws"] = torch.zeros((max(int(lib.krs_xchg_route_workspace_bytes(B, F, S)) // 4, 1),), dtype=torch.int32, device="cuda")
        b["vocab_dev"] = torch.tensor(self.vocab_sizes, dtype=torch.int64, device="cuda")
        offs = []
        for o in range(S):                                  # every rank can compute every peer's arena layout
            ro, tot = shard_row_offsets(self.vocab_sizes, o, S)
            if tot >= 2 ** 31:
                raise ValueError("a shard's arena must stay below 2^31 rows")
            offs.extend(ro)
        b["owner_row_off"] = torch.tensor(offs, dtype=torch.int32, device="cuda")
        b["epoch"] = 0
        self.cg.reserve(min(self.total_rows, S * B * F), "cuda")
        x = KrsXchg()
        x.S, x.me, x.F, x.E, x.B = S, me, F, E, B
        for name in ("off_flags", "off_hdr", "off_rows", "off_pos", "off_x0", "off_grad"):
            setattr(x, name, getattr(lay, name))
        b["xchg"] = x
        if not self._sim:
            if S > 1:
                allh = [None] * S
                dist.all_gather_object(allh, handle)
                for s in range(S):
                    x.peer_base[s] = base if s == me else self._open(allh[s])
                dist.barrier()
            else:
                x.peer_base[0] = base
        return b

    # ------------------------------------------------------------------ cross-rank ordering
    def _barrier(self, b, s):
        if self._sim:
            return                                   # one process, one stream: stream order is the barrier
        b["epoch"] += 1
        check(lib.krs_xchg_barrier(C.byref(b["xchg"]), b["epoch"], self._timeout, s))

    def check_exchange_errors(self, B: int | None = None) -> None:
        """Raises if a barrier timed out or the compact gradient buffer overflowed (synchronises the device)."""
        for bb, b in self._bufs.items():
            if B is not None and bb != B:
                continue
            word = int(b["flags"][XCHG_MAX_SHARDS].item())
            if word & 0x7fffffff:
                raise L.KrsError(f"row-sharded exchange: barrier timed out waiting for peers (mask {word & 0x7fffffff:#x})")
            if word & 0x80000000 or word < 0:
                raise L.KrsError("row-sharded exchange: more distinct touched rows than the compact gradient buffer holds")

    # ------------------------------------------------------------------ exchange phases
    def _route(self, b, B, s, fill_nan=True):
        ids = b["ids"]
        check(lib.krs_xchg_route(C.byref(b["xchg"]), self._parity, ptr(ids), 1 if ids.dtype == torch.int64 else 0, ids.stride(0),
                                 ptr(b["vocab_dev"]), ptr(b["owner_row_off"]), ptr(b["route_ws"]), 1 if fill_nan else 0, s))

    def _serve(self, b, s, train=True):
        check(lib.krs_xchg_gather_push(C.byref(b["xchg"]), self._parity, ptr(self.emb), ptr(self.cg.touched) if train else None, s))

    def _gather_into(self, b, B, s, train=True):
        self._route(b, B, s)
        self._barrier(b, s)                          # every rank's request lists are complete
        self._serve(b, s, train)
        self._barrier(b, s)                          # every owner delivered its rows into my activation

    def _pull_grads(self, b, s):
        cg = self.cg
        cg.scan()
        check(lib.krs_xchg_grad_pull(C.byref(b["xchg"]), self._parity, ptr(cg.touched), ptr(cg.wordprefix), ptr(cg.blockbase),
                                     ptr(cg.compact), ptr(cg.uniq_rows), cg.cap_rows, s))

    def _scatter_from(self, b, B, cur, s):
        assert cur.data_ptr() == b[b["grad_name"]].data_ptr(), "dL/dx0 must land in the exchange region"
        if self.world_ > 1 and not self._sim:
            self._ar = dist.all_reduce(self.dense_grad_flat, async_op=True)   # overlaps the owner-side backward
        self._barrier(b, s)                          # every rank's dL/dx0 is complete
        self._pull_grads(b, s)
        self._parity ^= 1

    def _update_tables(self, optimizer):
        optimizer._update_compact(self.emb, self.cg)

    def _sync_gradients(self):
        if self._ar is not None:
            self._ar.wait()
            self._ar = None

    def _end_of_step(self):
        """Nothing: the double-buffered request lists make the next step's first barrier sufficient."""

    def train_on_batch_graph(self, *a, **k):
        raise NotImplementedError("the row-sharded step is launched eagerly (its barriers spin on peer flags)")

    # ------------------------------------------------------------------ inference through the layer API
    def forward(self, ids, sparse_arena: bool = False):
        """Collective: every rank must call it with a batch of the same size."""
        if self._sim:
            raise RuntimeError("use SimGroup.predict in single-process simulation")
        ids = self._as_ids(ids)
        B = ids.shape[0]
        b = self._step_buffers(B)
        b["ids"].copy_(ids)
        with torch.no_grad():
            self._gather_into(b, B, stream(), train=False)
            self._parity ^= 1
        return self._dense_forward(b)

    def _dense_forward(self, b):
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
        for b in self._bufs.values():            # drop the tensors aliasing the regions before freeing them
            for k in ("region", "flags", "hdr", "req_rows", "req_pos"):
                b.pop(k, None)
        self._bufs = {}
        for p in self._owned:
            lib.krs_ipc_free(C.c_void_p(p))
        self._owned = []


class SimGroup:
    """S ShardedDCN ranks inside one process on one GPU (test / self-check harness).  Each phase of the protocol runs for
    every rank before the next phase starts, so stream order plays the role of the barriers; the dense-gradient
    all-reduce is a plain sum.  Kernels, regions, request lists and optimizers are exactly the multi-process ones."""

    def __init__(self, vocab_sizes, world: int, **kw):
        self.world = world
        self.ranks = [ShardedDCN(vocab_sizes, rank=r, world=world, sim=True, **kw) for r in range(world)]
        for m in self.ranks:
            m._sim_group = self

    def _wire(self, B):
        bs = [m._step_buffers(B) for m in self.ranks]
        for b in bs:
            for s in range(self.world):
                b["xchg"].peer_base[s] = bs[s]["region_base"]
        return bs

    def train_on_batch(self, ids_per_rank, labels_per_rank, optimizers, denom: int):
        B = ids_per_rank[0].shape[0]
        bs = self._wire(B)
        s = stream()
        for m, b, ids, y in zip(self.ranks, bs, ids_per_rank, labels_per_rank):
            b["ids"].copy_(ids)
            b["labels"].copy_(y.reshape(-1))
            m._route(b, B, s)
        for m, b in zip(self.ranks, bs):
            m._serve(b, s, train=True)
        curs = [m._dense_step(b, B, denom, s) for m, b in zip(self.ranks, bs)]
        total = torch.stack([m.dense_grad_flat for m in self.ranks]).sum(0)       # the all-reduce
        for m in self.ranks:
            m.dense_grad_flat.copy_(total)
        for m, b, cur in zip(self.ranks, bs, curs):
            m._scatter_from(b, B, cur, s)
        for m, opt in zip(self.ranks, optimizers):
            opt.iterations += 1
            with torch.no_grad():
                m._update_tables(opt)
                opt._update(m.dense_flat, m.dense_grad_flat, None)
        return [b["loss"] for b in bs]

    def predict(self, ids_per_rank):
        B = ids_per_rank[0].shape[0]
        bs = self._wire(B)
        s = stream()
        with torch.no_grad():
            for m, b, ids in zip(self.ranks, bs, ids_per_rank):
                b["ids"].copy_(ids)
                m._route(b, B, s)
            for m, b in zip(self.ranks, bs):
                m._serve(b, s, train=False)
            for m in self.ranks:
                m._parity ^= 1
            return [m._dense_forward(b) for m, b in zip(self.ranks, bs)]

    def check_errors(self):
        for m in self.ranks:
            m.check_exchange_errors()
