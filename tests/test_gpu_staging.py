"""keras_rs_b200.staging.prefetch: pinned double buffering on a copy stream delivers every host batch, in order and
intact, whether the host tensors are pinned or pageable, with more batches than slots and while the consumer's stream is busy."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("pinned", [True, False])
@pytest.mark.parametrize("depth", [1, 2, 3])
def test_prefetch_delivers_batches_in_order(pinned, depth):
    from keras_rs_b200.staging import prefetch
    g = torch.Generator().manual_seed(depth)
    n = 7
    ids = [torch.randint(0, 1000, (513, 5), generator=g, dtype=torch.int64) for _ in range(n)]
    ys = [torch.rand((513,), generator=g) for _ in range(n)]
    if pinned:
        ids, ys = [t.pin_memory() for t in ids], [t.pin_memory() for t in ys]
    busy = torch.randn((2048, 2048), device="cuda")
    seen = 0
    for k, (d_ids, d_y) in enumerate(prefetch(zip(ids, ys), depth=depth)):
        assert d_ids.is_cuda and d_ids.dtype == torch.int64 and tuple(d_ids.shape) == (513, 5)
        acc = d_ids.sum() + 0 * (busy @ busy).sum().long()          # consumer work enqueued behind the batch
        assert int(acc) == int(ids[k].sum())
        assert torch.equal(d_ids.cpu(), ids[k]) and torch.equal(d_y.cpu(), ys[k])
        seen += 1
    assert seen == n


def test_prefetch_feeds_the_training_step():
    """The staged device tensors drive DCN.train_on_batch exactly like directly supplied tensors."""
    import numpy as np
    import keras_rs_b200 as K
    from keras_rs_b200.dcn import DCN
    from keras_rs_b200.staging import prefetch
    rng = np.random.default_rng(0)
    vocab, B = [50, 33, 64], 128
    batches = [(torch.from_numpy(np.stack([rng.integers(0, v, size=B) for v in vocab], axis=1).astype(np.int32)).pin_memory(),
                torch.from_numpy(rng.uniform(size=B).astype(np.float32)).pin_memory()) for _ in range(5)]
    m1, m2 = DCN(vocab, embedding_dim=32, num_cross_layers=2, dense_units=(16,), seed=3), DCN(vocab, embedding_dim=32, num_cross_layers=2, dense_units=(16,), seed=3)
    o1, o2 = K.optimizers.Adagrad(0.05), K.optimizers.Adagrad(0.05)
    l1 = [float(m1.train_on_batch(i.cuda(), y.cuda(), o1)) for i, y in batches]
    l2 = [float(m2.train_on_batch(i, y, o2)) for i, y in prefetch(batches, depth=2)]
    np.testing.assert_allclose(l1, l2, rtol=1e-6)
    assert torch.allclose(m1.emb, m2.emb, rtol=1e-6, atol=1e-9)
