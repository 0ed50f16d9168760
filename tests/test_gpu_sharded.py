"""Row-sharded (MOD) multi-GPU step vs the oracle on the GLOBAL batch.  Needs >= 2 GPUs on one box
(run under `gpurun --gpus 2`); skipped otherwise."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    try:
        sys.path.insert(0, ROOT)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
        import torch.distributed as dist
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
        import keras_rs_b200 as K
        from keras_rs_b200.dcn import DCN
        from keras_rs_b200.sharded import ShardedDCN
        from oracle import np_oracle as O
        from util import assert_close, npy
        vocab, E, Bl = [50, 33, 64, 7], 32, 64
        m = ShardedDCN(vocab, rank=rank, world=world, embedding_dim=E, num_cross_layers=2, dense_units=(16,), seed=5)
        # oracle parameters: the unsharded tables are re-assembled from every rank's shard
        shards = [None] * world
        dist.all_gather_object(shards, [npy(t) for t in m.tables()])
        tables = [O.mod_unshard_table([shards[s][f] for s in range(world)]) for f in range(len(vocab))]
        params = dict(tables=tables, cross=[dict(V=npy(c.kernel), b=npy(c.bias)) for c in m.cross],
                      mlp=[(npy(d.kernel), npy(d.bias), "relu" if d._act_id else None) for d in m.mlp])
        flat = lambda P: P["tables"] + [a for c in P["cross"] for a in (c["V"], c["b"])] + [a for W, b, _ in P["mlp"] for a in (W, b)]
        st = [dict(m=np.zeros_like(a), v=np.zeros_like(a)) for a in flat(params)]
        opt = K.optimizers.AdamW(0.01)
        for step in range(1, 3):
            rng = np.random.default_rng(100 + step)
            gids = np.stack([rng.integers(0, v, size=Bl * world) for v in vocab], axis=1).astype(np.int32)
            gy = rng.uniform(size=Bl * world).astype(np.float32)
            cache = {}
            pred = O.dcn_forward(params, gids, cache)
            loss_ref, dpred = O.mse_loss(pred, gy)
            g = O.dcn_backward(params, gids, dpred, cache)
            gl = g["tables"] + [a for c in g["cross"] for a in (c["V"], c["b"])] + [a for dW, db in g["mlp"] for a in (dW, db)]
            new = []
            for a, ga, s in zip(flat(params), gl, st):
                p2, s["m"], s["v"] = O.adamw_step(a, s["m"], s["v"], ga, step, lr=0.01)
                new.append(p2)
            nt = len(vocab)
            params["tables"] = new[:nt]
            k = nt
            for c in params["cross"]:
                c["V"], c["b"] = new[k], new[k + 1]
                k += 2
            params["mlp"] = [(new[k + 2 * i], new[k + 2 * i + 1], params["mlp"][i][2]) for i in range(len(params["mlp"]))]
            lids = torch.from_numpy(gids[rank * Bl:(rank + 1) * Bl]).cuda()
            ly = torch.from_numpy(gy[rank * Bl:(rank + 1) * Bl]).cuda()
            loss = m.train_on_batch(lids, ly, opt, denom=Bl * world)
            tot = loss.clone()
            dist.all_reduce(tot)
            np.testing.assert_allclose(float(tot), float(loss_ref), rtol=2e-4)
        for f, t in enumerate(m.tables()):
            assert_close(npy(t), params["tables"][f][rank::world], rel=2e-4, what=f"rank {rank} table {f}")
        for c, pc in zip(m.cross, params["cross"]):
            assert_close(npy(c.kernel), pc["V"], rel=2e-4, what="V")
        # forward through the layer API reads peer shards too
        pred = m.predict(torch.from_numpy(gids[:8]).cuda())
        assert_close(npy(pred), O.dcn_forward(params, gids[:8]), rel=1e-4, what="sharded predict")
        dist.barrier()
        m.close()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, "FAIL: " + traceback.format_exc()))


def test_sharded_dcn_world2():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(60)
    for r, msg in res:
        assert msg == "ok", f"rank {r}: {msg}"
