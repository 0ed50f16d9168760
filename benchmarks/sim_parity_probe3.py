#!/usr/bin/env python
"""Simulated-rank AdamW parity with a POISONED caching allocator: every torch.empty buffer starts as NaN (or 1e30), so any
read of uninitialised memory shows up.  Usage: sim_parity_probe3.py WORLD [nan|big|none]"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
import keras_rs_b200 as K
from keras_rs_b200.sharded import SimGroup
from oracle import np_oracle as O
from oracle import parity as PAR
npy = lambda t: t.detach().float().cpu().numpy()
K.set_gemm_engine("ffma")
world = int(sys.argv[1]); mode = sys.argv[2] if len(sys.argv) > 2 else "nan"
if mode != "none":
    junk = [torch.full((1 << 26,), float("nan") if mode == "nan" else 1e30, device="cuda") for _ in range(8)]   # 2 GB
    small = [torch.full((n,), float("nan") if mode == "nan" else 1e30, device="cuda") for n in (64, 256, 1024, 4096, 65536, 1 << 20) for _ in range(32)]
    del junk, small
vocab, E, Bl, steps = [1000, 777, 1000, 50], 32, 256, 3
g = SimGroup(vocab, world, embedding_dim=E, num_cross_layers=2, dense_units=(32,), seed=11)
m0 = g.ranks[0]
tables = [O.mod_unshard_table([npy(m.tables()[f]) for m in g.ranks]) for f in range(len(vocab))]
tr = PAR.OracleTrainer(PAR.params_of(tables, m0.cross, m0.mlp), "adamw", lr=0.01)
opts = [K.optimizers.AdamW(0.01) for _ in range(world)]
rel_loss = 0.0
for si, (gids, gy) in enumerate(PAR.make_batches(vocab, Bl, world, steps, seed=4242, bad_ids=True)):
    ref = tr.train(gids, gy)
    losses = g.train_on_batch([torch.from_numpy(gids[r * Bl:(r + 1) * Bl]).cuda() for r in range(world)],
                              [torch.from_numpy(gy[r * Bl:(r + 1) * Bl]).cuda() for r in range(world)], opts, Bl * world)
    rel_loss = max(rel_loss, abs(sum(float(l) for l in losses) - ref) / abs(ref))
worst = 0.0
for r, m in enumerate(g.ranks):
    for f, t in enumerate(m.tables()):
        worst = max(worst, PAR.max_rel(npy(t), tr.P["tables"][f][r::world]))
dense = max(PAR.max_rel(npy(c.kernel), pc["V"]) for c, pc in zip(m0.cross, tr.P["cross"]))
print(json.dumps(dict(world=world, poison=mode, rel_loss=rel_loss, worst_table_rel=worst, dense_rel=dense)))
