#!/usr/bin/env python
"""world-8 simulated AdamW parity, per step: first step at which a table row deviates, with its optimizer state."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
import keras_rs_b200 as K
from keras_rs_b200.sharded import SimGroup
from oracle import np_oracle as O
from oracle import parity as PAR
npy = lambda t: t.detach().float().cpu().numpy()
K.set_gemm_engine("ffma")
vocab, E, Bl, steps, world = [1000, 777, 1000, 50], 32, 256, 3, 8
g = SimGroup(vocab, world, embedding_dim=E, num_cross_layers=2, dense_units=(32,), seed=11)
m0 = g.ranks[0]
tables = [O.mod_unshard_table([npy(m.tables()[f]) for m in g.ranks]) for f in range(len(vocab))]
tr = PAR.OracleTrainer(PAR.params_of(tables, m0.cross, m0.mlp), "adamw", lr=0.01)
opts = [K.optimizers.AdamW(0.01) for _ in range(world)]
for si, (gids, gy) in enumerate(PAR.make_batches(vocab, Bl, world, steps, seed=4242, bad_ids=True)):
    tr.train(gids, gy)
    g.train_on_batch([torch.from_numpy(gids[r * Bl:(r + 1) * Bl]).cuda() for r in range(world)],
                     [torch.from_numpy(gy[r * Bl:(r + 1) * Bl]).cuda() for r in range(world)], opts, Bl * world)
    worst = (0.0, None)
    for r, m in enumerate(g.ranks):
        st = opts[r]._state[id(m.emb)]
        for f, t in enumerate(m.tables()):
            ref = tr.P["tables"][f][r::world]; got = npy(t)
            err = np.abs(got - ref); sc = max(np.abs(ref).max(), 1e-30)
            i, j = np.unravel_index(int(err.argmax()), err.shape)
            if err[i, j] / sc > worst[0]:
                a = m.row_off[f] + i
                worst = (float(err[i, j] / sc), dict(rank=r, table=f, local_row=int(i), col=int(j), got=float(got[i, j]), ref=float(ref[i, j]),
                          m_gpu=float(st["m"][a, j]), v_gpu=float(st["v"][a, j]), m_ref=float(tr.m[f][i * world + r, j]), v_ref=float(tr.v[f][i * world + r, j])))
    dW = [(PAR.max_rel(npy(c.kernel), pc["V"]), PAR.max_rel(npy(c.bias), pc["b"])) for c, pc in zip(m0.cross, tr.P["cross"])]
    print(json.dumps(dict(step=si + 1, worst_table_rel=worst[0], where=worst[1], cross_rel=dW)))
