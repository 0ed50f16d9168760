"""keras_rs.layers.{HardNegativeMining, RemoveAccidentalHits, SamplingProbabilityCorrection} on the row kernels of
csrc/rowops.cu vs the numpy oracle and the properties the reference's own tests assert
(hard_negative_mining_test.py:58-85, remove_accidental_hits_test.py:62-131, sampling_probability_correction_test.py:53-84),
plus krs_row_topk itself and the gradient of the row selection."""
import numpy as np
import pytest
import torch

import keras_rs_b200 as K
from oracle import np_oracle as O

pytestmark = pytest.mark.gpu
SHAPE_3D = (15, 20, 10)


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def N(t):
    return t.detach().cpu().numpy()


def _inputs(rank, seed=42):
    rng = np.random.default_rng(seed)
    shape = SHAPE_3D[-rank:]
    logits = rng.uniform(size=shape).astype(np.float32)
    n = shape[-1]
    hot = rng.integers(0, n, size=shape[:-1])
    labels = np.eye(n, dtype=np.float32)[hot]
    return logits, labels


@pytest.mark.parametrize("rank", [1, 2, 3])
@pytest.mark.parametrize("num_hard_negatives", [3, 30])
def test_hard_negative_mining(rank, num_hard_negatives):
    logits, labels = _inputs(rank)
    layer = K.layers.HardNegativeMining(num_hard_negatives)
    out_logits, out_labels = layer(T(logits), T(labels))
    out_logits, out_labels = N(out_logits), N(out_labels)
    n = logits.shape[-1]
    assert out_logits.shape[-1] == min(num_hard_negatives + 1, n) == layer.compute_output_shape(logits.shape)[0][-1]
    # logits of the positives are always returned
    np.testing.assert_allclose((out_logits * out_labels).sum(-1), (logits * labels).sum(-1), rtol=1e-6)
    # with the label column lifted to the top, the highest k+1 logits are returned
    lifted = logits + labels * 1000.0
    out2, _ = layer(T(lifted), T(labels))
    np.testing.assert_allclose(np.sort(lifted, axis=-1)[..., -num_hard_negatives - 1:], np.sort(N(out2), axis=-1), rtol=1e-6)
    # and the oracle agrees element for element (same descending order)
    exp_logits, exp_labels = O.hard_negative_mining(logits, labels, num_hard_negatives)
    np.testing.assert_array_equal(out_logits, exp_logits)
    np.testing.assert_array_equal(out_labels, exp_labels)


@pytest.mark.parametrize("logits_rank,ids_rank", [(1, 1), (2, 1), (2, 2), (3, 1), (3, 2), (3, 3)])
def test_remove_accidental_hits(logits_rank, ids_rank):
    logits, labels = _inputs(logits_rank)
    rng = np.random.default_rng(7)
    ids = rng.integers(0, logits.shape[-1], size=SHAPE_3D[-ids_rank:]).astype(np.int32)
    out = N(K.layers.RemoveAccidentalHits()(T(logits), T(labels), T(ids)))
    assert out.shape == logits.shape and out.dtype == np.float32
    np.testing.assert_array_equal(out, O.remove_accidental_hits(logits, labels, ids))
    # logits of the labels are unchanged; every entry moves by at most SMALLEST_FLOAT (remove_accidental_hits_test.py:71-131)
    np.testing.assert_allclose((out * labels).sum(-1), (logits * labels).sum(-1), rtol=1e-6)
    assert np.abs(out - logits).max() <= O.SMALLEST_FLOAT * 1.01


def test_remove_accidental_hits_marks_duplicates_of_the_positive():
    logits = np.zeros((2, 4), np.float32)                 # zero logits make the subnormal increment visible
    labels = np.array([[0, 1, 0, 0], [1, 0, 0, 0]], np.float32)
    ids = np.array([5, 7, 7, 5], np.int32)
    out = N(K.layers.RemoveAccidentalHits()(T(logits), T(labels), T(ids)))
    np.testing.assert_array_equal(out > 0, np.array([[0, 0, 1, 0], [0, 0, 0, 1]], bool))


def test_remove_accidental_hits_errors():
    layer = K.layers.RemoveAccidentalHits()
    with pytest.raises(ValueError, match="`labels` and `logits` should have the same shape"):
        layer(torch.zeros(10, 20).cuda(), torch.zeros(10, 30).cuda(), torch.zeros(20).cuda())
    with pytest.raises(ValueError, match="`candidate_ids` should have the same shape as .* `labels`"):
        layer(torch.zeros(10, 20).cuda(), torch.zeros(10, 20).cuda(), torch.zeros(30).cuda())


@pytest.mark.parametrize("logits_rank,probs_rank", [(1, 1), (2, 1), (2, 2), (3, 1), (3, 2), (3, 3)])
def test_sampling_probability_correction(logits_rank, probs_rank):
    rng = np.random.default_rng(42)
    logits = rng.uniform(size=SHAPE_3D[-logits_rank:]).astype(np.float32)
    probs = rng.uniform(0.01, 0.99, size=SHAPE_3D[-probs_rank:]).astype(np.float32)
    layer = K.layers.SamplingProbabilityCorrection()
    out = N(layer(T(logits), T(probs)))
    assert (logits < out).all()                                          # log of a probability < 1 is negative
    np.testing.assert_allclose(out, O.sampling_probability_correction(logits, probs), rtol=1e-6, atol=1e-6)
    zeros = probs * (rng.uniform(size=probs.shape) >= 0.5)
    out0 = N(layer(T(logits), T(zeros.astype(np.float32))))
    assert (logits < out0).all() and np.isfinite(out0).all()             # epsilon keeps log(0) away


def test_helpers_serialization_round_trip():
    for layer in (K.layers.HardNegativeMining(num_hard_negatives=3), K.layers.RemoveAccidentalHits(),
                  K.layers.SamplingProbabilityCorrection(epsilon=1e-5)):
        restored = K.layers.deserialize(K.layers.serialize(layer))
        assert type(restored) is type(layer) and restored.get_config() == layer.get_config()


@pytest.mark.parametrize("rows,n,k", [(7, 10, 3), (33, 1000, 100), (4, 16384, 128), (5, 257, 257), (3, 1, 1)])
def test_row_topk_matches_oracle_order(rows, n, k):
    """krs_row_topk: values descending, ties -> lowest index first (jax.lax.top_k order), exact."""
    from keras_rs_b200._lib import check, lib, ptr, stream
    rng = np.random.default_rng(n)
    x = rng.normal(size=(rows, n)).astype(np.float32)
    x[:, ::7] = x[:, :1]                                   # plenty of exact ties
    if n > 3:
        x[0, 2] = np.inf; x[0, 3] = -np.inf
    tx = T(x)
    vals = torch.empty((rows, k), device="cuda"); idx = torch.empty((rows, k), device="cuda", dtype=torch.int32)
    ids = T(rng.integers(0, 10 ** 6, size=(rows, n)).astype(np.int32)); out_ids = torch.empty_like(idx)
    check(lib.krs_row_topk(ptr(tx), rows, n, n, None, 0, 0.0, k, ptr(vals), ptr(idx), None, 0, None, ptr(ids), n, ptr(out_ids), stream()))
    order = np.lexsort((np.arange(n)[None, :].repeat(rows, 0), -x), axis=-1)[:, :k]      # value desc, index asc
    np.testing.assert_array_equal(N(idx), order.astype(np.int32))
    np.testing.assert_array_equal(N(vals), np.take_along_axis(x, order, axis=1))
    np.testing.assert_array_equal(N(out_ids), np.take_along_axis(N(ids), order, axis=1))


def test_hard_negative_mining_gradient_scatters_to_selected_columns():
    rng = np.random.default_rng(3)
    logits, labels = _inputs(2)
    x = T(logits).requires_grad_(True)
    out_l, out_y = K.layers.HardNegativeMining(4)(x, T(labels))
    g = rng.normal(size=tuple(out_l.shape)).astype(np.float32)
    out_l.backward(T(g))
    exp_l, _ = O.hard_negative_mining(logits, labels, 4)
    boosted = logits + labels * O.MAX_FLOAT
    order = np.lexsort((np.arange(logits.shape[-1])[None, :].repeat(logits.shape[0], 0), -boosted), axis=-1)[:, :5]
    exp = np.zeros_like(logits)
    np.put_along_axis(exp, order, g, axis=1)
    np.testing.assert_array_equal(N(x.grad), exp)
    for layer, args in ((K.layers.RemoveAccidentalHits(), (T(labels), T(np.arange(logits.shape[-1], dtype=np.int32)))),
                        (K.layers.SamplingProbabilityCorrection(), (T(np.full(logits.shape[-1], 0.5, np.float32)),))):
        x2 = T(logits).requires_grad_(True)
        layer(x2, *args).backward(T(np.ones_like(logits)))
        np.testing.assert_array_equal(N(x2.grad), np.ones_like(logits))


def test_helpers_reject_cpu_tensors():
    from keras_rs_b200._lib import KrsError
    with pytest.raises(KrsError):
        K.layers.HardNegativeMining(2)(torch.zeros(3, 4), torch.zeros(3, 4))
