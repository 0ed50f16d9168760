"""Host-side logic that needs no GPU: ragged -> dense preprocessing of the DistributedEmbedding front end
(base_distributed_embedding.py:31-92), exchange-region layout, MOD shard bookkeeping, the oracle's routed request lists and
the row-sparse optimizer rules (lazy Adam / FTRL) against their dense counterparts."""
import numpy as np
import pytest
import torch

from oracle import np_oracle as O


def test_ragged_to_dense_matches_the_reference_loop():
    """Same result as the reference's per-row Python loop (base_distributed_embedding.py:73-85) for every input form."""
    from keras_rs_b200.layers.distributed_embedding import ragged_to_dense_inputs
    rng = np.random.default_rng(0)
    lens = [3, 0, 5, 1, 5, 2]
    rows = [rng.integers(0, 100, size=n) for n in lens]
    w = [rng.uniform(size=n).astype(np.float32) for n in lens]
    for L in (None, 5, 8):
        width = L or max(lens)
        exp_ids = np.zeros((len(rows), width), np.int64)
        exp_w1 = np.zeros((len(rows), width), np.float32)
        exp_w = np.zeros((len(rows), width), np.float32)
        for i, r in enumerate(rows):
            exp_ids[i, :len(r)] = r
            exp_w1[i, :len(r)] = 1.0
            exp_w[i, :len(r)] = w[i]
        obj = np.empty((len(rows),), dtype=object)
        for i, r in enumerate(rows):
            obj[i] = r
        splits = np.concatenate([[0], np.cumsum(lens)])
        forms = [([list(r) for r in rows], [list(x) for x in w]), (obj, w), ((np.concatenate(rows), splits), (np.concatenate(w), splits))]
        for x, xw in forms:
            ids, ones = ragged_to_dense_inputs(x, None, L, device="cpu")
            np.testing.assert_array_equal(ids.numpy(), exp_ids)
            np.testing.assert_array_equal(ones.numpy(), exp_w1)
            ids2, ww = ragged_to_dense_inputs(x, xw, L, device="cpu")
            np.testing.assert_array_equal(ids2.numpy(), exp_ids)
            np.testing.assert_array_equal(ww.numpy(), exp_w)
    with pytest.raises(ValueError, match="exceeds the dense row length"):
        ragged_to_dense_inputs([list(r) for r in rows], None, 4, device="cpu")
    t = torch.zeros((4, 2), dtype=torch.int32)
    assert ragged_to_dense_inputs(t, None, 2)[0] is t          # dense inputs pass through untouched
    a = np.zeros((4, 2), np.int32)
    assert ragged_to_dense_inputs(a, None, 2)[0] is a


def test_region_layout_is_aligned_and_disjoint():
    from keras_rs_b200._lib import XCHG_MAX_SHARDS
    from keras_rs_b200.sharded import RegionLayout
    for B, F, E in ((1, 1, 4), (96, 5, 32), (65536, 26, 128)):
        lay = RegionLayout(B, F, E)
        P = B * F
        spans = [(lay.off_flags, (XCHG_MAX_SHARDS + 1) * 4), (lay.off_hdr, 2 * (XCHG_MAX_SHARDS + 1) * 4), (lay.off_rows, 2 * P * 4),
                 (lay.off_pos, 2 * P * 4), (lay.off_x0, P * E * 4), (lay.off_grad, P * E * 4)]
        end = 0
        for off, n in spans:
            assert off % 256 == 0 and off >= end
            end = off + n
        assert lay.nbytes >= end


def test_route_requests_partition_and_order():
    """Every valid id appears exactly once, in the bucket of its owner, with increasing positions; wrapped negatives resolve
    to their row, out-of-range ids are left out."""
    from keras_rs_b200.sharding import shard_row_offsets
    rng = np.random.default_rng(1)
    vocab, S, B = [37, 64, 5], 4, 200
    ids = np.stack([rng.integers(-v, v + 3, size=B) for v in vocab], axis=1)
    offs = [shard_row_offsets(vocab, o, S)[0] for o in range(S)]
    rows, pos = O.route_requests(ids, vocab, S, offs)
    seen = {}
    for o in range(S):
        assert (np.diff(pos[o]) > 0).all()
        for r, p in zip(rows[o], pos[o]):
            b, f = divmod(int(p), len(vocab))
            i = int(ids[b, f])
            i = i + vocab[f] if i < 0 else i
            assert 0 <= i < vocab[f] and i % S == o and r == offs[o][f] + i // S
            seen[int(p)] = True
    valid = sum(1 for b in range(B) for f in range(len(vocab)) if 0 <= (ids[b, f] + vocab[f] if ids[b, f] < 0 else ids[b, f]) < vocab[f])
    assert len(seen) == valid


def test_lazy_rules_equal_dense_rules_on_touched_rows_and_leave_the_rest():
    rng = np.random.default_rng(2)
    p = rng.normal(size=(10, 4)).astype(np.float32)
    g = np.zeros_like(p)
    rows = np.zeros((10,), bool)
    rows[[1, 4, 7]] = True
    g[rows] = rng.normal(size=(3, 4)).astype(np.float32)
    m, v = np.zeros_like(p), np.zeros_like(p)
    p2, m2, v2 = O.lazy_adam_step(p, m, v, g, 1, lr=0.01, rows=rows)
    pd, md, vd = O.adamw_step(p, m, v, g, 1, lr=0.01, wd=0.0)
    np.testing.assert_array_equal(p2[rows], pd[rows])
    np.testing.assert_array_equal(p2[~rows], p[~rows])
    np.testing.assert_array_equal(m2[~rows], 0 * m[~rows])
    acc, lin = np.full_like(p, 0.1), np.zeros_like(p)
    p3, a3, l3 = O.ftrl_step(p, acc, lin, g, lr=0.05, l1=0.001, l2=0.01, rows=rows)
    np.testing.assert_array_equal(p3[~rows], p[~rows])
    assert np.abs(p3[rows] - p[rows]).max() > 0 and np.isfinite(p3).all()
    np.testing.assert_allclose(a3[rows], 0.1 + g[rows] ** 2, rtol=1e-6)
