#!/usr/bin/env python
"""Single-GPU probe (simulated ranks): compact gradient rows after grad_pull vs the oracle's table gradients, step 1."""
import os, sys, json, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
import keras_rs_b200 as K
from keras_rs_b200._lib import stream
from keras_rs_b200.sharded import SimGroup
from keras_rs_b200.sharding import local_vocab
from oracle import np_oracle as O
from oracle import parity as PAR
npy = lambda t: t.detach().float().cpu().numpy()
K.set_gemm_engine("ffma")
vocab, E, Bl = [1000, 777, 1000, 50], 32, 256
for world in (4, 8):
    g = SimGroup(vocab, world, embedding_dim=E, num_cross_layers=2, dense_units=(32,), seed=11)
    m0 = g.ranks[0]
    tables = [O.mod_unshard_table([npy(m.tables()[f]) for m in g.ranks]) for f in range(len(vocab))]
    P = PAR.params_of(tables, m0.cross, m0.mlp)
    gids, gy = PAR.make_batches(vocab, Bl, world, 1, seed=4242, bad_ids=True)[0]
    cache = {}
    pred = O.dcn_forward(P, gids, cache); loss, dpred = O.mse_loss(pred, gy); og = O.dcn_backward(P, gids, dpred, cache)
    B = Bl; bs = g._wire(B); s = stream()
    for r, (m, b) in enumerate(zip(g.ranks, bs)):
        b["ids"].copy_(torch.from_numpy(gids[r * Bl:(r + 1) * Bl]).cuda()); b["labels"].copy_(torch.from_numpy(gy[r * Bl:(r + 1) * Bl]).cuda())
        m._route(b, B, s)
    for m, b in zip(g.ranks, bs): m._serve(b, s, train=True)
    curs = [m._dense_step(b, B, Bl * world, s) for m, b in zip(g.ranks, bs)]
    # activation gradient vs oracle
    gx0 = np.concatenate([npy(c) for c in curs], axis=0)
    for m, b, cur in zip(g.ranks, bs, curs): m._scatter_from(b, B, cur, s)
    torch.cuda.synchronize()
    worst = []
    for r, m in enumerate(g.ranks):
        nu = int(m.cg.n_unique.item()); rows = m.cg.uniq_rows[:nu].cpu().numpy(); comp = npy(m.cg.compact[:nu])
        dense = np.zeros((m.total_rows, E), np.float32); dense[rows] = comp
        for f, v in enumerate(vocab):
            lv = local_vocab(v, r, world)
            got = dense[m.row_off[f]:m.row_off[f] + lv]; ref = og["tables"][f][r::world]
            err = np.abs(got - ref).max(axis=1); sc = max(np.abs(og["tables"][f]).max(), 1e-30)
            i = int(err.argmax())
            grow = i * world + r
            idx, ok = O.resolve_ids(gids[:, f], v)
            occ = [int(((idx[q * Bl:(q + 1) * Bl] == grow) & ok[q * Bl:(q + 1) * Bl]).sum()) for q in range(world)]
            worst.append((float(err[i] / sc), dict(rank=r, table=f, local_row=i, global_row=grow, lookups_per_requester=occ,
                                                   got_norm=float(np.abs(got[i]).sum()), ref_norm=float(np.abs(ref[i]).sum()))))
    worst.sort(key=lambda t: -t[0])
    print(json.dumps(dict(world=world, top=worst[:4])))
