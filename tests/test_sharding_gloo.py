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
    # the routed protocol of csrc/exchange.cu, emulated with the oracle: every rank builds request lists for its own ids,
    # owners serve the rows of their shard, activations land at the requested positions; gradients flow back the same way
    vocab2 = [37, 64, 5]
    tabs = [np.arange(v * 2, dtype=np.float32).reshape(v, 2) + 1000 * f for f, v in enumerate(vocab2)]
    offs_all = [shard_row_offsets(vocab2, o, world)[0] for o in range(world)]
    my_ids = np.stack([np.random.default_rng(10 + rank).integers(-v, v, size=9) for v in vocab2], axis=1)
    rows, pos = O.route_requests(my_ids, vocab2, world, offs_all)
    reqs = [None] * world
    dist.all_gather_object(reqs, (rows, pos))
    arena = np.zeros((shard_row_offsets(vocab2, rank, world)[1], 2), np.float32)
    for f, t_ in enumerate(tabs):
        sh = O.mod_shard_table(t_, world)[rank]
        arena[offs_all[rank][f]:offs_all[rank][f] + len(sh)] = sh
    served = [(arena[reqs[r][0][rank]], reqs[r][1][rank]) for r in range(world)]        # what I send to requester r
    got = [None] * world
    dist.all_gather_object(got, served)
    x0 = np.full((9 * len(vocab2), 2), np.nan, np.float32)
    for o in range(world):
        vals, where = got[o][rank]
        x0[where] = vals
    ref = np.concatenate([O.embedding_lookup(t_, my_ids[:, f]) for f, t_ in enumerate(tabs)], axis=1).reshape(9 * len(vocab2), 2)
    ok &= bool(np.array_equal(x0, ref))
    from keras_rs_b200.sharded import RegionLayout
    lay = RegionLayout(9, len(vocab2), 4)
    lays = [None] * world
    dist.all_gather_object(lays, [lay.off_flags, lay.off_hdr, lay.off_rows, lay.off_pos, lay.off_x0, lay.off_grad, lay.nbytes])
    ok &= all(l == lays[0] for l in lays) and lay.off_x0 % 256 == 0 and lay.off_grad % 256 == 0
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
