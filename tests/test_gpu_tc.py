"""tcgen05 (3xTF32) GEMM engine vs float64 references and vs the oracle, through the same C ABI.
Runs last (alphabetical order) so that a fault in the tensor-core kernel cannot mask other results."""
import os

import numpy as np
import pytest
import torch

from oracle import np_oracle as O
from util import assert_close, dev, npy

pytestmark = pytest.mark.gpu


# engines under test: "tcgen05" (SS operands) and "tcgen05_ts" (A operand in tensor memory)
ENGINES = [e for e in os.environ.get("KRS_TEST_TC_ENGINES", "tcgen05,tcgen05_ts").split(",") if e]


@pytest.fixture(params=ENGINES)
def tc(request):
    import keras_rs_b200 as K
    K.set_gemm_engine(request.param)
    yield K
    K.set_gemm_engine("ffma")


def _count(K):
    return K._lib.lib.krs_gemm_tc_launch_count()


@pytest.mark.parametrize("tA,tB", [(False, True), (False, False), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,K_", [(128, 256, 64), (256, 208, 128), (384, 832, 832), (1000, 192, 840), (130, 48, 20), (832, 832, 4096)])
def test_tc_sgemm(tc, tA, tB, M, N, K_):
    rng = np.random.default_rng(M + N + K_)
    A = rng.normal(size=(K_, M) if tA else (M, K_)).astype(np.float32)
    B = rng.normal(size=(N, K_) if tB else (K_, N)).astype(np.float32)
    if tA and M % 4:   # stored (K,M): leading dimension must be 16-byte aligned for TMA
        pytest.skip("lda not a multiple of 4 -> FFMA fallback (covered elsewhere)")
    ref = (A.T if tA else A).astype(np.float64) @ (B.T if tB else B).astype(np.float64)
    before = _count(tc)
    got = tc.ops.sgemm(dev(A), dev(B), tA, tB)
    torch.cuda.synchronize()
    if (not tA and K_ % 4) or (tB and K_ % 4) or (not tB and N % 4):
        return
    assert _count(tc) == before + 1, "tcgen05 kernel did not run (fell back to FFMA)"
    assert_close(npy(got), ref, what=f"tc sgemm tA={tA} tB={tB} {M}x{N}x{K_}")


def test_tc_split_k_weight_gradient(tc):
    rng = np.random.default_rng(5)
    Kb, M, N = 8192, 832, 192        # dW = x^T dz with the batch as the reduction dimension
    x = rng.normal(size=(Kb, M)).astype(np.float32)
    dz = rng.normal(size=(Kb, N)).astype(np.float32)
    before = _count(tc)
    got = tc.ops.sgemm(dev(x), dev(dz), True, False)
    assert _count(tc) == before + 1
    assert_close(npy(got), x.astype(np.float64).T @ dz.astype(np.float64), what="split-k dW")


@pytest.mark.parametrize("P", [None, 64])
@pytest.mark.parametrize("act", [None, "relu"])
def test_tc_feature_cross_vs_oracle(tc, P, act):
    rng = np.random.default_rng(11)
    B, D = 640, 832
    x0 = rng.normal(size=(B, D)).astype(np.float32)
    x = rng.normal(size=(B, D)).astype(np.float32)
    layer = tc.layers.FeatureCross(projection_dim=P, diag_scale=0.25, pre_activation=act,
                                   bias_initializer=tc.initializers.RandomUniform(-0.5, 0.5, seed=3))
    tx0, tx = dev(x0).requires_grad_(True), dev(x).requires_grad_(True)
    before = _count(tc)
    y = layer(tx0, tx)
    assert _count(tc) > before
    U = npy(layer.down_proj_kernel) if P is not None else None
    V, b = npy(layer.kernel), npy(layer.bias)
    f64 = lambda a: None if a is None else a.astype(np.float64)
    ref = O.feature_cross(f64(x0), f64(x), f64(V), f64(b), f64(U), 0.25, act)
    assert_close(npy(y), ref, what="tc cross fwd")
    gy = rng.normal(size=(B, D)).astype(np.float32)
    y.backward(dev(gy))
    hz = dev(x) if P is None else tc.ops.linear_no_bias(dev(x), layer.down_proj_kernel.detach())
    z_gpu = npy(tc.ops.dense(hz.contiguous(), layer.kernel.detach(), layer.bias.detach(), 0))
    r = O.feature_cross_bwd(f64(gy), f64(x0), f64(x), f64(V), f64(b), f64(U), 0.25, act, z_for_grad=z_gpu)
    assert_close(npy(tx0.grad), r["dx0"], what="tc dx0")
    assert_close(npy(tx.grad), r["dx"], what="tc dx")
    assert_close(npy(layer.kernel.grad), r["dV"], what="tc dV")
    assert_close(npy(layer.bias.grad), r["db"], what="tc db", scale=np.abs(gy * x0).sum(axis=0).max())
    if P is not None:
        assert_close(npy(layer.down_proj_kernel.grad), r["dU"], what="tc dU")


def test_tc_dcn_step_matches_ffma_engine(tc):
    from keras_rs_b200.dcn import DCN
    rng = np.random.default_rng(3)
    vocab, E, B = [1000] * 26, 32, 1024
    ids = np.stack([rng.integers(0, v, size=B) for v in vocab], axis=1).astype(np.int32)
    y = rng.uniform(size=B).astype(np.float32)
    m = DCN(vocab, embedding_dim=E, num_cross_layers=3, dense_units=(192, 192), seed=0)
    before = _count(tc)
    loss_tc = float(m.forward_backward(dev(ids), dev(y)))
    assert _count(tc) >= before + 12
    g_tc, e_tc = m.dense_grad_flat.clone(), m.emb_grad.clone()
    m.emb_grad.zero_(); m.emb_touched.zero_()
    tc.set_gemm_engine("ffma")
    loss_ff = float(m.forward_backward(dev(ids), dev(y)))
    np.testing.assert_allclose(loss_tc, loss_ff, rtol=1e-5)
    assert_close(npy(g_tc), npy(m.dense_grad_flat), what="dense grads tc vs ffma")
    assert_close(npy(e_tc), npy(m.emb_grad), what="emb grads tc vs ffma")


def test_tc_full_size_accuracy_c2_shapes(tc):
    """BASELINE C2 shapes (B=65536, D=832): forward x@V and the batch-reduction weight gradient x^T dz
    against float64 products computed independently on the device."""
    g = torch.Generator(device="cuda").manual_seed(7)
    B, D = 65536, 832
    x = torch.randn((B, D), device="cuda", generator=g)
    V = (torch.rand((D, D), device="cuda", generator=g) * 2 - 1) * 0.06
    dz = torch.randn((B, D), device="cuda", generator=g)
    before = _count(tc)
    y = tc.ops.sgemm(x, V)                       # NN, K = 832
    dV = tc.ops.sgemm(x, dz, True, False)        # TN, K = 65536 (split + RED)
    dx = tc.ops.sgemm(dz, V, False, True)        # NT
    assert _count(tc) == before + 3
    for got, ref, what in ((y, x.double() @ V.double(), "fwd"), (dV, x.double().T @ dz.double(), "dV"),
                           (dx, dz.double() @ V.double().T, "dx")):
        err = float((got.double() - ref).abs().max())
        scale = float(ref.abs().max())
        assert err <= 1e-5 * scale, f"{what}: max abs err {err:.3e} > 1e-5 * {scale:.3e}"
        print(f"tc accuracy {what}: max abs err / max|ref| = {err / scale:.2e}")


def test_tc_mlperf_dcnv2_shape_low_rank(tc):
    """examples/ml_perf shape (DLRM-DCNv2): D = 3456, projection_dim = 512 (ml_perf/model.py:317-325).  The down
    projection reduces over K = 3456 > 1024, i.e. the K-split + fp32 RED path; compared with float64 on device."""
    g = torch.Generator(device="cuda").manual_seed(3)
    B, D, P = 1024, 3456, 512
    x0 = torch.randn((B, D), device="cuda", generator=g)
    x = torch.randn((B, D), device="cuda", generator=g)
    layer = tc.layers.FeatureCross(projection_dim=P)
    before = _count(tc)
    y = layer(x0, x)
    assert _count(tc) >= before + 2
    U, V, b = layer.down_proj_kernel.detach().double(), layer.kernel.detach().double(), layer.bias.detach().double()
    ref = x0.double() * ((x.double() @ U) @ V + b) + x.double()
    err = float((y.double() - ref).abs().max() / ref.abs().max())
    assert err <= 1e-5, err
