"""Writes tests/golden/*.json — the known-answer vectors the REFERENCE's own tests hold for the
hot path, transcribed by hand (the reference cannot be imported here: `import keras` fails,
SURVEY.md F3).  Every expected value below is either a literal copied from the cited reference
test or computed by the closed-form expression that test uses, with plain Python floats —
deliberately NOT through oracle/ so the fixtures pin the oracle rather than echo it.

Run:  python tests/golden/make_golden.py      (idempotent; output is committed)
"""
import json
import math
import os

HERE = os.path.dirname(os.path.abspath(__file__))


def dot(a, b):
    return sum(x * y for x, y in zip(a, b))


def feature_cross():
    # keras_rs/src/layers/feature_interaction/feature_cross_test.py:15-19
    x0 = [[0.1, 0.2, 0.3]]
    x = [[0.4, 0.5, 0.6]]
    cases = [
        # :21-26 full rank, kernel_initializer="ones", bias zeros
        dict(name="full_ones", x0=x0, x=x, projection_dim=None, diag_scale=0.0,
             expected=[[0.55, 0.8, 1.05]], weight_shapes=[[3, 3], [3]]),
        # :34-39 low rank P=1, ones
        dict(name="low_rank_ones", x0=x0, x=x, projection_dim=1, diag_scale=0.0,
             expected=[[0.55, 0.8, 1.05]], weight_shapes=[[3, 1], [1, 3], [3]]),
        # :49-52 one input
        dict(name="one_input", x0=x0, x=None, projection_dim=None, diag_scale=0.0,
             expected=[[0.16, 0.32, 0.48]], weight_shapes=[[3, 3], [3]]),
        # :67-73 diag_scale=1.0
        dict(name="diag_scale_1", x0=x0, x=x, projection_dim=None, diag_scale=1.0,
             expected=[[0.59, 0.9, 1.23]], weight_shapes=[[3, 3], [3]]),
        # :75-79 pre_activation=zeros_like => output == x
        dict(name="pre_activation_zeros", x0=x0, x=x, projection_dim=None, diag_scale=0.0,
             pre_activation="zeros_like", expected=x, weight_shapes=[[3, 3], [3]]),
    ]
    return dict(source="feature_cross_test.py:15-79", atol=1e-6, rtol=1e-6, cases=cases)


def dot_interaction():
    # keras_rs/src/layers/feature_interaction/dot_interaction_test.py:17-53
    f1 = [0.1, -4.3, 0.2, 1.1, 0.3]
    f2 = [2.0, 3.2, -1.0, 0.0, 1.0]
    f3 = [0.0, 1.0, -3.0, -2.2, -0.2]
    f11, f12, f13 = dot(f1, f1), dot(f1, f2), dot(f1, f3)
    f22, f23, f33 = dot(f2, f2), dot(f2, f3), dot(f3, f3)
    cases = [
        dict(self_interaction=False, skip_gather=False, expected=[[f12, f13, f23]]),
        dict(self_interaction=False, skip_gather=True,
             expected=[[0, 0, 0, f12, 0, 0, f13, f23, 0]]),
        dict(self_interaction=True, skip_gather=False,
             expected=[[f11, f12, f22, f13, f23, f33]]),
        dict(self_interaction=True, skip_gather=True,
             expected=[[f11, 0, 0, f12, f22, 0, f13, f23, f33]]),
    ]
    return dict(source="dot_interaction_test.py:17-91", atol=1e-6, rtol=1e-6,
                inputs=[[f1], [f2], [f3]], cases=cases)


def embed_reduce():
    # keras_rs/src/layers/embedding/embed_reduce_test.py:45-119 — EmbedReduce(10, 20); the table
    # is random in the reference, expected values are linear combinations of its rows: we store
    # the coefficient lists [(row, coeff)...] per output row and a divisor.
    cases = []
    for combiner in ("sum", "mean", "sqrtn"):
        for use_weights in (False, True):
            # dense 1-D: inputs [1, 2], weights [1.0, 2.0]   (:45-47, :91-95)
            if combiner == "sum" and use_weights:
                exp = [[(1, 1.0)], [(2, 2.0)]]
            else:
                exp = [[(1, 1.0)], [(2, 1.0)]]
            cases.append(dict(combiner=combiner, rank=1, use_weights=use_weights,
                              inputs=[1, 2], weights=[1.0, 2.0], expected_terms=exp,
                              divisors=[1.0, 1.0]))
            # dense 2-D: inputs [[1,2],[3,4]], weights [[1,2],[3,4]]   (:48-50, :96-107)
            if use_weights:
                exp = [[(1, 1.0), (2, 2.0)], [(3, 3.0), (4, 4.0)]]
            else:
                exp = [[(1, 1.0), (2, 1.0)], [(3, 1.0), (4, 1.0)]]
            div = [1.0, 1.0]
            if combiner == "mean":
                div = [3.0, 7.0] if use_weights else [2.0, 2.0]
            elif combiner == "sqrtn":
                div = [math.sqrt(5.0 if use_weights else 2.0), math.sqrt(25.0 if use_weights else 2.0)]
            cases.append(dict(combiner=combiner, rank=2, use_weights=use_weights,
                              inputs=[[1, 2], [3, 4]], weights=[[1.0, 2.0], [3.0, 4.0]],
                              expected_terms=exp, divisors=div))
    # ragged/sparse case of the same test (:51-84, :108-117) expressed densely with zero-weight
    # padding (the torch backend has dense inputs only, :37-43): row0=[1], row1=[2,3,4,5]
    for combiner in ("sum", "mean", "sqrtn"):
        for use_weights in (False, True):
            w1 = [1.0, 2.0, 3.0, 4.0] if use_weights else [1.0, 1.0, 1.0, 1.0]
            exp = [[(1, 1.0)], [(2, w1[0]), (3, w1[1]), (4, w1[2]), (5, w1[3])]]
            div = [1.0, 1.0]
            if combiner == "mean":
                div = [1.0, 10.0 if use_weights else 4.0]
            elif combiner == "sqrtn":
                div = [1.0, math.sqrt(30.0 if use_weights else 4.0)]
            cases.append(dict(combiner=combiner, rank=2, use_weights=True, padded=True,
                              inputs=[[1, 0, 0, 0], [2, 3, 4, 5]],
                              weights=[[1.0, 0.0, 0.0, 0.0], w1],
                              expected_terms=exp, divisors=div))
    return dict(source="embed_reduce_test.py:45-119", atol=1e-6, rtol=1e-6, vocab=10, dim=20,
                cases=cases)


def distributed_embedding():
    # keras_rs/src/layers/embedding/distributed_embedding_test.py:448-454,574-599: inputs are
    # [2, 3] repeated; expected rows emb[2], emb[3] (x2.0 when combiner == "sum" and weights given,
    # weights being [1.0, 2.0]-style per-sample weights for the dense 1-D case).
    return dict(source="distributed_embedding_test.py:448-454,574-599", ids=[2, 3],
                note="expected = [emb[2], emb[3]] ; with sum+weights w: [w0*emb[2], w1*emb[3]]")


def retrieval():
    # keras_rs/src/layers/retrieval/retrieval_test.py:21-40 (error regexes) and
    # brute_force_retrieval_test.py:13-64 (num_candidates=100, dim 4, 16 queries, k=20, ids = arange+3;
    # expected = argsort(-scores)[:, :k], scores atol 1e-4, indices exact)
    return dict(
        source="retrieval_test.py:21-48; brute_force_retrieval_test.py:13-64",
        errors=[
            dict(emb_shape=None, ids_shape=None, k=5, regex="`candidate_embeddings` is required."),
            dict(emb_shape=[10], ids_shape=None, k=5,
                 regex="`candidate_embeddings` must be a tensor of rank 2"),
            dict(emb_shape=[3, 10], ids_shape=None, k=5,
                 regex="The number of candidates provided \\(3\\) is less than"),
            dict(emb_shape=[6, 10], ids_shape=[4], k=5,
                 regex="The `candidate_embeddings` and `candidate_is` tensors must have "
                       "the same number of rows"),
        ],
        brute_force=dict(num_candidates=100, dim=4, num_queries=16, k=20, id_offset=3,
                         score_atol=1e-4))


def tril():
    # dot_interaction.py:118-132 evaluated by hand for N=3 and N=4
    return dict(source="dot_interaction.py:118-132",
                cases=[dict(n=3, self_interaction=False, idx=[3, 6, 7]),
                       dict(n=3, self_interaction=True, idx=[0, 3, 4, 6, 7, 8]),
                       dict(n=4, self_interaction=False, idx=[4, 8, 9, 12, 13, 14])])


def main():
    out = dict(feature_cross=feature_cross(), dot_interaction=dot_interaction(),
               embed_reduce=embed_reduce(), distributed_embedding=distributed_embedding(),
               retrieval=retrieval(), tril=tril())
    for name, obj in out.items():
        with open(os.path.join(HERE, name + ".json"), "w") as f:
            json.dump(obj, f, indent=1)
    print("wrote", sorted(out))


if __name__ == "__main__":
    main()
