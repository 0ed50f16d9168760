"""Host-side mirrors of keras_rs.layers.{HardNegativeMining, RemoveAccidentalHits, SamplingProbabilityCorrection}
 vs the numpy oracle and the properties the reference's own tests assert
(hard_negative_mining_test.py:58-85, remove_accidental_hits_test.py:62-131, sampling_probability_correction_test.py:53-84).
They are torch tensor ops (no kernel of libkrs_b200.so), so they are checked on the CPU."""
import numpy as np
import pytest
import torch

import keras_rs_b200 as K
from oracle import np_oracle as O

SHAPE_3D = (15, 20, 10)


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
    out_logits, out_labels = layer(torch.from_numpy(logits), torch.from_numpy(labels))
    out_logits, out_labels = out_logits.numpy(), out_labels.numpy()
    n = logits.shape[-1]
    assert out_logits.shape[-1] == min(num_hard_negatives + 1, n) == layer.compute_output_shape(logits.shape)[0][-1]
    # logits of the positives are always returned
    np.testing.assert_allclose((out_logits * out_labels).sum(-1), (logits * labels).sum(-1), rtol=1e-6)
    # with the label column lifted to the top, the highest k+1 logits are returned
    lifted = logits + labels * 1000.0
    out2, _ = layer(torch.from_numpy(lifted), torch.from_numpy(labels))
    np.testing.assert_allclose(np.sort(lifted, axis=-1)[..., -num_hard_negatives - 1:], np.sort(out2.numpy(), axis=-1), rtol=1e-6)
    # and the oracle agrees element for element (same descending order)
    exp_logits, exp_labels = O.hard_negative_mining(logits, labels, num_hard_negatives)
    np.testing.assert_array_equal(out_logits, exp_logits)
    np.testing.assert_array_equal(out_labels, exp_labels)


@pytest.mark.parametrize("logits_rank,ids_rank", [(1, 1), (2, 1), (2, 2), (3, 1), (3, 2), (3, 3)])
def test_remove_accidental_hits(logits_rank, ids_rank):
    logits, labels = _inputs(logits_rank)
    rng = np.random.default_rng(7)
    ids = rng.integers(0, logits.shape[-1], size=SHAPE_3D[-ids_rank:]).astype(np.int32)
    out = K.layers.RemoveAccidentalHits()(torch.from_numpy(logits), torch.from_numpy(labels), torch.from_numpy(ids)).numpy()
    assert out.shape == logits.shape and out.dtype == np.float32
    np.testing.assert_array_equal(out, O.remove_accidental_hits(logits, labels, ids))
    # logits of the labels are unchanged; every entry moves by at most SMALLEST_FLOAT (remove_accidental_hits_test.py:71-131)
    np.testing.assert_allclose((out * labels).sum(-1), (logits * labels).sum(-1), rtol=1e-6)
    assert np.abs(out - logits).max() <= O.SMALLEST_FLOAT * 1.01


def test_remove_accidental_hits_marks_duplicates_of_the_positive():
    logits = np.zeros((2, 4), np.float32)                 # zero logits make the subnormal increment visible
    labels = np.array([[0, 1, 0, 0], [1, 0, 0, 0]], np.float32)
    ids = np.array([5, 7, 7, 5], np.int32)
    out = K.layers.RemoveAccidentalHits()(torch.from_numpy(logits), torch.from_numpy(labels), torch.from_numpy(ids)).numpy()
    np.testing.assert_array_equal(out > 0, np.array([[0, 0, 1, 0], [0, 0, 0, 1]], bool))


def test_remove_accidental_hits_errors():
    layer = K.layers.RemoveAccidentalHits()
    with pytest.raises(ValueError, match="`labels` and `logits` should have the same shape"):
        layer(torch.zeros(10, 20), torch.zeros(10, 30), torch.zeros(20))
    with pytest.raises(ValueError, match="`candidate_ids` should have the same shape as .* `labels`"):
        layer(torch.zeros(10, 20), torch.zeros(10, 20), torch.zeros(30))


@pytest.mark.parametrize("logits_rank,probs_rank", [(1, 1), (2, 1), (2, 2), (3, 1), (3, 2), (3, 3)])
def test_sampling_probability_correction(logits_rank, probs_rank):
    rng = np.random.default_rng(42)
    logits = rng.uniform(size=SHAPE_3D[-logits_rank:]).astype(np.float32)
    probs = rng.uniform(0.01, 0.99, size=SHAPE_3D[-probs_rank:]).astype(np.float32)
    layer = K.layers.SamplingProbabilityCorrection()
    out = layer(torch.from_numpy(logits), torch.from_numpy(probs)).numpy()
    assert (logits < out).all()                                          # log of a probability < 1 is negative
    np.testing.assert_allclose(out, O.sampling_probability_correction(logits, probs), rtol=1e-6, atol=1e-6)
    zeros = probs * (rng.uniform(size=probs.shape) >= 0.5)
    out0 = layer(torch.from_numpy(logits), torch.from_numpy(zeros.astype(np.float32))).numpy()
    assert (logits < out0).all() and np.isfinite(out0).all()             # epsilon keeps log(0) away


def test_helpers_serialization_round_trip():
    for layer in (K.layers.HardNegativeMining(num_hard_negatives=3), K.layers.RemoveAccidentalHits(),
                  K.layers.SamplingProbabilityCorrection(epsilon=1e-5)):
        restored = K.layers.deserialize(K.layers.serialize(layer))
        assert type(restored) is type(layer) and restored.get_config() == layer.get_config()
