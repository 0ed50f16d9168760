#!/usr/bin/env python
"""Single-GPU reproduction of bench.py's N>1 parity self-check with simulated ranks (SimGroup): which table row deviates?"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
import keras_rs_b200 as K
from keras_rs_b200.sharded import SimGroup
from oracle import np_oracle as O
from oracle import parity as PAR
npy = lambda t: t.detach().float().cpu().numpy()
K.set_gemm_engine("ffma")
vocab, E, Bl, steps = [1000, 777, 1000, 50], 32, 256, 3
for world in (4, 8):
    for use_ever in (True, False):
        g = SimGroup(vocab, world, embedding_dim=E, num_cross_layers=2, dense_units=(32,), seed=11)
        if not use_ever:
            for m in g.ranks:
                m.emb._krs_ever = None
        m0 = g.ranks[0]
        tables = [O.mod_unshard_table([npy(m.tables()[f]) for m in g.ranks]) for f in range(len(vocab))]
        tr = PAR.OracleTrainer(PAR.params_of(tables, m0.cross, m0.mlp), "adamw", lr=0.01)
        opts = [K.optimizers.AdamW(0.01) for _ in range(world)]
        seen = [np.zeros((steps, v), bool) for v in vocab]
        for si, (gids, gy) in enumerate(PAR.make_batches(vocab, Bl, world, steps, seed=4242, bad_ids=True)):
            tr.train(gids, gy)
            for f, v in enumerate(vocab):
                idx, ok = O.resolve_ids(gids[:, f], v); seen[f][si, idx[ok]] = True
            g.train_on_batch([torch.from_numpy(gids[r * Bl:(r + 1) * Bl]).cuda() for r in range(world)],
                             [torch.from_numpy(gy[r * Bl:(r + 1) * Bl]).cuda() for r in range(world)], opts, Bl * world)
        g.check_errors()
        worst = (0.0, None)
        for r, m in enumerate(g.ranks):
            for f, t in enumerate(m.tables()):
                ref = tr.P["tables"][f][r::world]; got = npy(t)
                err = np.abs(got - ref).max(axis=1) / max(np.abs(ref).max(), 1e-30)
                i = int(err.argmax())
                if err[i] > worst[0]:
                    worst = (float(err[i]), dict(rank=r, table=f, local_row=i, global_row=i * world + r,
                                                 touched_per_step=[bool(seen[f][s][i * world + r]) for s in range(steps)]))
        dense = max(PAR.max_rel(npy(c.kernel), pc["V"]) for c, pc in zip(m0.cross, tr.P["cross"]))
        print(json.dumps(dict(world=world, ever=use_ever, worst_table_rel=worst[0], where=worst[1], dense_rel=dense)))
