"""Row-sharded exchange (csrc/exchange.cu) on ONE GPU: S ranks simulated inside one process (keras_rs_b200.sharded.SimGroup —
same kernels, regions, request lists and optimizers as the multi-process path; stream order replaces the flag barriers).
Integer work (routing, request lists, slot numbering) is checked bit-exact against the oracle; the training step is
checked against np_oracle on the GLOBAL batch at the north star's 1e-5 (optimizer amplification stated per case)."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import np_oracle as O
from oracle import parity as PAR
from util import assert_close, dev, npy

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def K():
    import keras_rs_b200 as K
    return K


def _region(K, B, F, E, S, me):
    from keras_rs_b200._lib import KrsXchg
    from keras_rs_b200.sharded import RegionLayout
    lay = RegionLayout(B, F, E)
    region = torch.zeros((lay.nbytes,), dtype=torch.uint8, device="cuda")
    x = KrsXchg()
    x.S, x.me, x.F, x.E, x.B = S, me, F, E, B
    for name in ("off_flags", "off_hdr", "off_rows", "off_pos", "off_x0", "off_grad"):
        setattr(x, name, getattr(lay, name))
    return lay, region, x


@pytest.mark.parametrize("S", [1, 2, 3, 8])
@pytest.mark.parametrize("idt", [torch.int32, torch.int64])
def test_route_request_lists_bit_exact(K, S, idt):
    """krs_xchg_route vs np_oracle.route_requests: same buckets, same (row, position) pairs in the same (position) order,
    including wrapped negative ids and ids that address no row (left out, NaN row in the activation)."""
    from keras_rs_b200._lib import XCHG_MAX_SHARDS, check, lib, ptr, stream
    from keras_rs_b200.sharding import shard_row_offsets
    rng = np.random.default_rng(S)
    vocab = [37, 64, 5, 1000, 129]
    B, F, E = 1501, len(vocab), 8          # B*F spans several 2048-position chunks, ragged tail
    ids = np.stack([rng.integers(0, v, size=B) for v in vocab], axis=1)
    ids[0, 0], ids[1, 1], ids[2, 2], ids[3, 3], ids[B - 1, 4] = -1, -64, 5, -1001, 10 ** 6
    lay, region, x = _region(K, B, F, E, S, 0)
    for s in range(S):
        x.peer_base[s] = region.data_ptr()
    offs = [shard_row_offsets(vocab, o, S)[0] for o in range(S)]
    d_ids = dev(ids, idt)
    ws = torch.zeros((int(lib.krs_xchg_route_workspace_bytes(B, F, S)) // 4 + 1,), dtype=torch.int32, device="cuda")
    x0 = lay.view(region, lay.off_x0, (B * F, E), torch.float32)
    d_vocab, d_offs = dev(np.array(vocab), torch.int64), dev(np.array(offs).reshape(-1), torch.int32)
    for parity in (0, 1):
        x0.fill_(1.0)
        check(lib.krs_xchg_route(C.byref(x), parity, ptr(d_ids), int(idt == torch.int64), F, ptr(d_vocab), ptr(d_offs), ptr(ws), 1,
                                 stream()))
        hdr = npy(lay.view(region, lay.off_hdr, (2, XCHG_MAX_SHARDS + 1), torch.int32))[parity]
        rows = npy(lay.view(region, lay.off_rows, (2, B * F), torch.int32))[parity]
        pos = npy(lay.view(region, lay.off_pos, (2, B * F), torch.int32))[parity]
        erows, epos = O.route_requests(ids, vocab, S, offs)
        assert hdr[0] == 0
        for o in range(S):
            np.testing.assert_array_equal(rows[hdr[o]:hdr[o + 1]], erows[o])
            np.testing.assert_array_equal(pos[hdr[o]:hdr[o + 1]], epos[o])
        assert hdr[S] == sum(len(r) for r in erows) == B * F - 3          # three ids address no row
        got = npy(x0)
        bad = [2 * F + 2, 3 * F + 3, (B - 1) * F + 4]
        assert np.isnan(got[bad]).all()
        keep = np.ones(B * F, bool)
        keep[bad] = False
        assert (got[keep] == 1.0).all()


def test_slot_scan_bit_exact(K):
    """slot(row) = number of set bits before it: prefix popcount over 1024-word blocks vs numpy cumsum."""
    from keras_rs_b200._lib import check, lib, ptr, stream
    rng = np.random.default_rng(3)
    nrows = 32 * (3 * 1024 + 77)
    bits = (rng.random(nrows) < 0.03)
    bits[[0, 31, 32, nrows - 1]] = True
    words = np.packbits(bits.reshape(-1, 32)[:, ::-1], axis=1).view(">u4").astype(np.uint32).reshape(-1)
    t = dev(words.view(np.int32))
    wp = torch.zeros_like(t)
    bb = torch.zeros((int(lib.krs_slot_scan_blocks(nrows)),), dtype=torch.int32, device="cuda")
    nu = torch.zeros((1,), dtype=torch.int32, device="cuda")
    check(lib.krs_slot_scan(ptr(t), t.numel(), ptr(wp), ptr(bb), ptr(nu), stream()))
    pc = np.array([bin(int(w)).count("1") for w in words])
    excl = np.concatenate([[0], np.cumsum(pc)[:-1]])
    got = npy(bb).astype(np.int64)[np.arange(len(words)) // 1024] + npy(wp).astype(np.int64)
    np.testing.assert_array_equal(got, excl)
    assert int(nu.item()) == int(bits.sum())


def _global_tables(group, vocab):
    return [O.mod_unshard_table([npy(m.tables()[f]) for m in group.ranks]) for f in range(len(vocab))]


def _mk_opt(K, name):
    return {"adamw": lambda: K.optimizers.AdamW(0.01), "adagrad": lambda: K.optimizers.Adagrad(0.01),
            "sgd": lambda: K.optimizers.SGD(0.01), "lazy_adam": lambda: K.optimizers.Adam(0.01, sparse_rows=True),
            "ftrl": lambda: K.optimizers.Ftrl(0.01, l1_regularization_strength=0.001, l2_regularization_strength=0.01)}[name]()


@pytest.mark.parametrize("world,E,idt,opt_name,rel_params", [
    (2, 32, torch.int32, "adamw", 5e-5), (2, 32, torch.int64, "adagrad", 1e-5), (3, 8, torch.int32, "sgd", 1e-5),
    (4, 128, torch.int64, "adagrad", 1e-5), (8, 32, torch.int32, "adamw", 5e-5), (8, 128, torch.int32, "sgd", 1e-5),
    (4, 48, torch.int32, "adagrad", 1e-5), (2, 16, torch.int32, "lazy_adam", 5e-5), (2, 16, torch.int32, "ftrl", 5e-5),
    (1, 32, torch.int32, "adagrad", 1e-5),
])
def test_sharded_step_matches_oracle_on_the_global_batch(K, world, E, idt, opt_name, rel_params):
    """3 training steps of the row-sharded model (S simulated ranks) == 3 oracle steps on the global batch: per-step loss at
    1e-5, every table shard and dense weight afterwards.  Duplicate ids inside and across ranks, wrapped negative ids,
    vocabularies smaller than the shard count."""
    from keras_rs_b200.sharded import SimGroup
    vocab, Bl, steps = [50, 33, 64, 7, 3], 96, 3
    g = SimGroup(vocab, world, embedding_dim=E, num_cross_layers=2, dense_units=(16,), seed=5, dense_activation="tanh")
    for m in g.ranks[1:]:                                   # data-parallel replicas start from identical dense weights
        m.dense_flat.copy_(g.ranks[0].dense_flat)
    m0 = g.ranks[0]
    tr = PAR.OracleTrainer(PAR.params_of(_global_tables(g, vocab), m0.cross, m0.mlp), opt_name, lr=0.01,
                           l1=0.001, l2=0.01)
    opts = [_mk_opt(K, opt_name) for _ in range(world)]
    dense_opt_is_table_opt = opt_name in ("adamw", "adagrad", "sgd", "lazy_adam")
    for gids, gy in PAR.make_batches(vocab, Bl, world, steps, bad_ids=True):
        ref_loss = tr.train(gids, gy)
        ids_r = [dev(gids[r * Bl:(r + 1) * Bl], idt) for r in range(world)]
        y_r = [dev(gy[r * Bl:(r + 1) * Bl]) for r in range(world)]
        if not dense_opt_is_table_opt:                      # FTRL: the oracle trainer updates dense weights with Adagrad
            dense_opts = getattr(g, "_dense_opts", None) or [K.optimizers.Adagrad(0.01) for _ in range(world)]
            g._dense_opts = dense_opts
            losses = _train_split(g, ids_r, y_r, opts, dense_opts, Bl * world)
        else:
            losses = g.train_on_batch(ids_r, y_r, opts, Bl * world)
        total = float(sum(float(l) for l in losses))
        assert abs(total - ref_loss) <= 1e-5 * max(abs(ref_loss), 1e-6), (total, ref_loss)
    g.check_errors()
    P = tr.P
    for r, m in enumerate(g.ranks):
        for f, t in enumerate(m.tables()):
            assert_close(npy(t), P["tables"][f][r::world], rel=rel_params, what=f"rank {r} table {f}")
        for c, pc in zip(m.cross, P["cross"]):
            assert_close(npy(c.kernel), pc["V"], rel=rel_params, what="cross V")
            assert_close(npy(c.bias), pc["b"], rel=rel_params, what="cross b", scale=max(float(np.abs(pc["b"]).max()), 1e-3))
        for d, (W, b, _) in zip(m.mlp, P["mlp"]):
            assert_close(npy(d.kernel), W, rel=rel_params, what="mlp W")
        assert float(m.cg.compact.abs().max()) == 0.0 and int(m.cg.touched.abs().max()) == 0   # staging re-zeroed
    # forward through the protocol (no training side effects)
    gids = PAR.make_batches(vocab, 8, world, 1, seed=999)[0][0]
    preds = g.predict([dev(gids[r * 8:(r + 1) * 8], idt) for r in range(world)])
    ref = O.dcn_forward(P, gids)
    for r in range(world):
        assert_close(npy(preds[r]), ref[r * 8:(r + 1) * 8], rel=1e-5, what="sharded predict",
                     scale=max(float(np.abs(ref).max()), 1e-3))
        assert int(g.ranks[r].cg.touched.abs().max()) == 0


@pytest.mark.parametrize("world,E", [(2, 32), (4, 128), (8, 32), (8, 8), (3, 48)])
def test_compact_gradient_rows_match_oracle(K, world, E):
    """The owner-side compact gradient rows (after grad_pull, before any optimizer) against the oracle's dense table
    gradients of the GLOBAL batch, relative to max|g| — the direct check of the exchange's backward: SGD / Adagrad updates
    at lr = 0.01 are too small against the parameter scale to expose a wrong gradient row."""
    from keras_rs_b200._lib import stream
    from keras_rs_b200.sharded import SimGroup
    from keras_rs_b200.sharding import local_vocab
    vocab, Bl = [1000, 777, 64, 7], 256
    g = SimGroup(vocab, world, embedding_dim=E, num_cross_layers=2, dense_units=(16,), seed=7, dense_activation="tanh")
    m0 = g.ranks[0]
    P = PAR.params_of(_global_tables(g, vocab), m0.cross, m0.mlp)
    gids, gy = PAR.make_batches(vocab, Bl, world, 1, seed=77, bad_ids=True)[0]
    cache = {}
    pred = O.dcn_forward(P, gids, cache)
    _, dpred = O.mse_loss(pred, gy)
    og = O.dcn_backward(P, gids, dpred, cache)
    bs, s = g._wire(Bl), stream()
    for r, (m, b) in enumerate(zip(g.ranks, bs)):
        b["ids"].copy_(dev(gids[r * Bl:(r + 1) * Bl]))
        b["labels"].copy_(dev(gy[r * Bl:(r + 1) * Bl]))
        m._route(b, Bl, s)
    for m, b in zip(g.ranks, bs):
        m._serve(b, s, train=True)
    curs = [m._dense_step(b, Bl, Bl * world, s) for m, b in zip(g.ranks, bs)]
    for r, cur in enumerate(curs):         # activation gradient of every rank's slice first (names the failing side)
        ref = np.concatenate([og["x0"][r * Bl:(r + 1) * Bl]], axis=0) if "x0" in og else None
        if ref is not None:
            assert_close(npy(cur), ref, rel=1e-5, what=f"rank {r} dL/dx0")
    for m, b, cur in zip(g.ranks, bs, curs):
        m._scatter_from(b, Bl, cur, s)
    g.check_errors()
    for r, m in enumerate(g.ranks):
        nu = int(m.cg.n_unique.item())
        rows = m.cg.uniq_rows[:nu].cpu().numpy()
        assert (np.diff(rows) > 0).all()                       # slots are numbered in arena-row order
        dense = np.zeros((m.total_rows, E), np.float32)
        dense[rows] = npy(m.cg.compact[:nu])
        looked_up = 0
        for f, v in enumerate(vocab):
            lv = local_vocab(v, r, world)
            ref = og["tables"][f][r::world]
            assert_close(dense[m.row_off[f]:m.row_off[f] + lv], ref, rel=1e-5, what=f"rank {r} table {f} gradient rows",
                         scale=max(float(np.abs(og["tables"][f]).max()), 1e-30))
            idx, ok = O.resolve_ids(gids[:, f], v)
            looked_up += len(set(int(i) for i in idx[ok] if i % world == r))
        assert nu == looked_up                                  # one compact row per distinct looked-up row


def _train_split(g, ids_r, y_r, table_opts, dense_opts, denom):
    """SimGroup.train_on_batch with different optimizers for tables and dense weights."""
    from keras_rs_b200._lib import stream
    B = ids_r[0].shape[0]
    bs = g._wire(B)
    s = stream()
    for m, b, ids, y in zip(g.ranks, bs, ids_r, y_r):
        b["ids"].copy_(ids)
        b["labels"].copy_(y.reshape(-1))
        m._route(b, B, s)
    for m, b in zip(g.ranks, bs):
        m._serve(b, s, train=True)
    curs = [m._dense_step(b, B, denom, s) for m, b in zip(g.ranks, bs)]
    total = torch.stack([m.dense_grad_flat for m in g.ranks]).sum(0)
    for m in g.ranks:
        m.dense_grad_flat.copy_(total)
    for m, b, cur in zip(g.ranks, bs, curs):
        m._scatter_from(b, B, cur, s)
    for m, to, do in zip(g.ranks, table_opts, dense_opts):
        to.iterations += 1
        do.iterations += 1
        with torch.no_grad():
            m._update_tables(to)
            do._update(m.dense_flat, m.dense_grad_flat, None)
    return [b["loss"] for b in bs]


def test_compact_overflow_is_reported(K):
    """More distinct touched rows than the compact buffer holds must raise, not corrupt memory."""
    from keras_rs_b200._lib import KrsError
    from keras_rs_b200.sharded import SimGroup
    vocab, Bl = [4000], 512
    g = SimGroup(vocab, 2, embedding_dim=8, num_cross_layers=1, dense_units=(8,), seed=1)
    rng = np.random.default_rng(0)
    ids = [dev(rng.permutation(4000)[:Bl].reshape(Bl, 1), torch.int32) for _ in range(2)]
    y = [dev(rng.uniform(size=Bl).astype(np.float32)) for _ in range(2)]
    g._wire(Bl)
    for m in g.ranks:                      # shrink the staging below what the batch touches
        m.cg.cap_rows = 16
    opts = [K.optimizers.SGD(0.01) for _ in range(2)]
    g.train_on_batch(ids, y, opts, 2 * Bl)
    with pytest.raises(KrsError, match="compact gradient buffer"):
        g.check_errors()


def test_barrier_single_rank_and_timeout(K):
    """krs_xchg_barrier: with S = 1 the rank signals itself; a peer that never arrives sets the error word after the
    timeout instead of hanging the GPU."""
    from keras_rs_b200._lib import XCHG_MAX_SHARDS, check, lib, stream
    lay, region, x = _region(K, 4, 2, 8, 1, 0)
    x.peer_base[0] = region.data_ptr()
    flags = lay.view(region, lay.off_flags, (XCHG_MAX_SHARDS + 1,), torch.int32)
    for epoch in (1, 2, 3):
        check(lib.krs_xchg_barrier(C.byref(x), epoch, 5.0, stream()))
    torch.cuda.synchronize()
    assert int(flags[0].item()) == 3 and int(flags[XCHG_MAX_SHARDS].item()) == 0
    lay2, region2, x2 = _region(K, 4, 2, 8, 2, 0)            # "peer" 1 is a second region nobody drives
    other = torch.zeros_like(region2)
    x2.peer_base[0], x2.peer_base[1] = region2.data_ptr(), other.data_ptr()
    check(lib.krs_xchg_barrier(C.byref(x2), 1, 0.05, stream()))
    torch.cuda.synchronize()
    flags2 = lay2.view(region2, lay2.off_flags, (XCHG_MAX_SHARDS + 1,), torch.int32)
    assert int(flags2[XCHG_MAX_SHARDS].item()) == 0b10       # peer 1 missing
    assert int(lay2.view(other, lay2.off_flags, (XCHG_MAX_SHARDS + 1,), torch.int32)[0].item()) == 1   # my signal landed
