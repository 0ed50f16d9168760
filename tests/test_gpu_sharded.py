"""Row-sharded (MOD) multi-GPU step, one process per GPU over NVLink peer memory, vs the oracle on the GLOBAL batch.
Needs >= 2 GPUs on one box (`gpurun --gpus N`); skipped otherwise — the same protocol is covered on one GPU by
tests/test_gpu_exchange.py (simulated ranks).  Checks per-step loss at 1e-5, every table shard and dense weight after
3 steps, and the collective predict()."""
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


def _worker(rank, world, port, q, E, i64, opt_name, rel_params):
    try:
        sys.path.insert(0, ROOT)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
        import torch.distributed as dist
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
        import keras_rs_b200 as K
        from keras_rs_b200.sharded import ShardedDCN
        from oracle import np_oracle as O
        from oracle import parity as PAR
        from util import assert_close, npy
        K.set_gemm_engine("ffma")
        vocab, Bl, steps = [50, 33, 64, 7, 3], 96, 3
        m = ShardedDCN(vocab, rank=rank, world=world, embedding_dim=E, num_cross_layers=2, dense_units=(16,), seed=5,
                       barrier_timeout_s=30.0, dense_activation="tanh")     # smooth: see oracle/parity.py
        shards = [None] * world
        dist.all_gather_object(shards, [npy(t) for t in m.tables()])
        tables = [O.mod_unshard_table([shards[s][f] for s in range(world)]) for f in range(len(vocab))]
        tr = PAR.OracleTrainer(PAR.params_of(tables, m.cross, m.mlp), opt_name, lr=0.01)
        opt = {"adamw": K.optimizers.AdamW, "adagrad": K.optimizers.Adagrad, "sgd": K.optimizers.SGD}[opt_name](0.01)
        idt = torch.int64 if i64 else torch.int32
        for gids, gy in PAR.make_batches(vocab, Bl, world, steps, bad_ids=True):
            ref_loss = tr.train(gids, gy)
            lids = torch.from_numpy(gids[rank * Bl:(rank + 1) * Bl]).to(idt).cuda()
            ly = torch.from_numpy(gy[rank * Bl:(rank + 1) * Bl]).cuda()
            loss = m.train_on_batch(lids, ly, opt, denom=Bl * world)
            tot = loss.clone()
            dist.all_reduce(tot)
            assert abs(float(tot) - ref_loss) <= 1e-5 * max(abs(ref_loss), 1e-6), (float(tot), ref_loss)
        m.check_exchange_errors()
        P = tr.P
        for f, t in enumerate(m.tables()):
            assert_close(npy(t), P["tables"][f][rank::world], rel=rel_params, what=f"rank {rank} table {f}")
        for c, pc in zip(m.cross, P["cross"]):
            assert_close(npy(c.kernel), pc["V"], rel=rel_params, what="V")
        for d, (W, b, _) in zip(m.mlp, P["mlp"]):
            assert_close(npy(d.kernel), W, rel=rel_params, what="mlp W")
        assert float(m.cg.compact.abs().max()) == 0.0 and int(m.cg.touched.abs().max()) == 0
        gids = PAR.make_batches(vocab, 8, world, 1, seed=999)[0][0]
        pred = m.predict(torch.from_numpy(gids[rank * 8:(rank + 1) * 8]).cuda())
        ref = O.dcn_forward(P, gids)
        assert_close(npy(pred), ref[rank * 8:(rank + 1) * 8], rel=1e-5, what="sharded predict", scale=max(float(np.abs(ref).max()), 1e-3))
        m.check_exchange_errors()
        dist.barrier()
        m.close()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception:  # pragma: no cover
        import traceback
        q.put((rank, "FAIL: " + traceback.format_exc()))


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("E,i64,opt_name,rel_params", [(32, False, "adamw", 5e-5), (128, True, "adamw", 5e-5)])
def test_sharded_dcn_multi_process(world, E, i64, opt_name, rel_params):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, E, i64, opt_name, rel_params)) for r in range(world)]
    for p in procs:
        p.start()
    import queue
    import time
    res, deadline = [], time.monotonic() + 180
    while len(res) < world and time.monotonic() < deadline:
        try:
            res.append(q.get(timeout=5))
        except queue.Empty:
            pass
        if any(m != "ok" for _, m in res):            # a failed rank leaves its peers waiting in a collective
            break
    for p in procs:
        p.join(10 if len(res) == world else 0.1)
        if p.is_alive():
            p.terminate()
    for r, msg in res:
        assert msg == "ok", f"rank {r}: {msg}"
    assert len(res) == world, f"only {len(res)} of {world} ranks reported within the time limit"
