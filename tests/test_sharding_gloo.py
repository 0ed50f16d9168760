"""N>1 host logic on CPU: world_size-2 gloo processes agree on the MOD shard layout and on the
routing of every id (no GPU compute; the CUDA side is covered by tests/test_gpu_sharded.py)."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from keras_rs_b200.sharding import global_row, local_vocab, owner_and_local, shard_row_offsets
    from oracle import np_oracle as O
    vocab = [37, 64, 5, 1000]
    offs, total = shard_row_offsets(vocab, rank, world)
    # every rank can compute every peer's layout; check against what the peer itself computed
    mine = dict(offs=offs, total=total, lv=[local_vocab(v, rank, world) for v in vocab])
    allm = [None] * world
    dist.all_gather_object(allm, mine)
    ok = True
    for s in range(world):
        o, t = shard_row_offsets(vocab, s, world)
        ok &= (o == allm[s]["offs"] and t == allm[s]["total"])
        ok &= allm[s]["lv"] == [len(range(s, v, world)) for v in vocab]
    ok &= all(sum(allm[s]["lv"][i] for s in range(world)) == v for i, v in enumerate(vocab))
    # routing of a shared id stream is bit-identical to the oracle and invertible
    rng = np.random.default_rng(0)
    ids = rng.integers(0, 1000, size=500)
    owner, local = O.mod_route(ids, world)
    for i, o, l in zip(ids, owner, local):
        ok &= owner_and_local(int(i), world) == (int(o), int(l))
        ok &= global_row(int(o), int(l), world) == int(i)
    # emulate the exchange: each rank "serves" the rows it owns from its shard of a shared table
    table = np.arange(1000 * 4, dtype=np.float32).reshape(1000, 4)
    shard = O.mod_shard_table(table, world)[rank]
    served = {int(i): shard[int(l)] for i, o, l in zip(ids, owner, local) if o == rank}
    alls = [None] * world
    dist.all_gather_object(alls, served)
    merged = {}
    for d in alls:
        merged.update(d)
    ok &= all(np.array_equal(merged[int(i)], table[int(i)]) for i in ids)
    t = torch.tensor([1.0 if ok else 0.0])
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        q.put(float(t))
    dist.destroy_process_group()


def test_mod_sharding_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) == 1.0
