"""Full-size (BASELINE.json C3 / C4) property checks.  The oracle cannot score these sizes in seconds, so the kernels
are checked through size-independent properties against float64 products computed independently on the device for a
sample of rows.  The file name sorts last on purpose: these are the largest allocations of the suite."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_c4_full_size_topk_properties():
    """BruteForceRetrieval at C4 (Q 4096x64, C 1e7x64, k=100) on the tensor-pipe kernel: rows sorted, the returned
    indices carry the returned scores, and a sample of queries agrees with an exact float64 top-k."""
    import keras_rs_b200 as K
    nq, nc, d, k = 4096, 10_000_000, 64, 100
    g = torch.Generator(device="cuda").manual_seed(42)
    C = torch.randn((nc, d), device="cuda", generator=g)
    Q = torch.randn((nq, d), device="cuda", generator=g)
    before = K._lib.lib.krs_topk_tc_launch_count()
    s, i = K.ops.top_k_scores(Q, C, None, k)
    torch.cuda.synchronize()
    assert K._lib.lib.krs_topk_tc_launch_count() == before + 1, "the tensor-pipe top-k kernel did not run"
    assert tuple(s.shape) == (nq, k) and tuple(i.shape) == (nq, k) and i.dtype == torch.int32
    assert bool((s[:, 1:] <= s[:, :-1]).all()), "scores must be sorted descending"
    assert int(i.min()) >= 0 and int(i.max()) < nc
    # no candidate twice in a row of results
    srt = torch.sort(i, dim=1).values
    assert bool((srt[:, 1:] != srt[:, :-1]).all())
    # every returned index carries its returned score (float64 recomputation, 512 sampled queries) at the reference
    # test's own bar (brute_force_retrieval_test.py:59-64: scores 1e-4, indices exact)
    rows = torch.randperm(nq, device="cuda", generator=g)[:512]
    cand = C[i[rows].long()].double()                                        # (512, k, d)
    re = torch.einsum("qkd,qd->qk", cand, Q[rows].double())
    assert float((re - s[rows].double()).abs().max()) <= 1e-4
    # exact float64 top-k of 256 sampled queries: scores within 1e-4, indices EXACTLY equal wherever the reference order is
    # not a near-tie (neighbouring reference scores further apart than 2e-4; fp32-level scores cannot order closer pairs)
    Cd = C.double()
    checked = 0
    for c0 in range(0, 256, 64):
        qr = rows[c0:c0 + 64]
        ref = Q[qr].double() @ Cd.T                                          # (64, 1e7) float64
        rs, ri = torch.topk(ref, k + 1, dim=1)
        assert float((s[qr].double() - rs[:, :k]).abs().max()) <= 1e-4
        gap_prev = torch.cat([torch.full((64, 1), 1.0, device="cuda", dtype=torch.float64), rs[:, :k - 1] - rs[:, 1:k]], dim=1)
        gap_next = rs[:, :k] - rs[:, 1:k + 1]
        clear = (gap_prev > 2e-4) & (gap_next > 2e-4)
        assert bool((i[qr].long()[clear] == ri[:, :k][clear]).all()), "top-k indices differ from the exact float64 ranking"
        checked += int(clear.sum())
        del ref
    assert checked > 0.95 * 256 * k                                          # near-ties are rare: almost every slot was checked


def test_c3_full_size_dot_interaction_properties():
    """DotInteraction at C3 (27 features x 128 dims, batch 65536): a sample of rows against float64 Gram matrices, for
    the gathered lower triangle and for the masked full matrix."""
    import keras_rs_b200 as K
    B, N, E = 65536, 27, 128
    g = torch.Generator(device="cuda").manual_seed(3)
    buf = torch.randn((B, N * E), device="cuda", generator=g)
    feats = [buf[:, j * E:(j + 1) * E] for j in range(N)]
    rows = torch.randperm(B, device="cuda", generator=g)[:1024]
    F = buf[rows].reshape(-1, N, E).double()
    P = F @ F.transpose(1, 2)                                                 # (1024, N, N)
    scale = float(P.abs().max())
    out = K.layers.DotInteraction()(feats)
    assert tuple(out.shape) == (B, N * (N - 1) // 2)
    ii, jj = torch.tril_indices(N, N, offset=-1, device="cuda")              # row-major lower triangle: (1,0),(2,0),(2,1)...
    assert float((out[rows].double() - P[:, ii, jj]).abs().max()) <= 1e-5 * scale
    full = K.layers.DotInteraction(self_interaction=True, skip_gather=True)(feats)
    assert tuple(full.shape) == (B, N * N)
    mask = torch.tril(torch.ones((N, N), device="cuda", dtype=torch.float64))
    assert float((full[rows].double().reshape(-1, N, N) - P * mask).abs().max()) <= 1e-5 * scale
