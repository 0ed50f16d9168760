import numpy as np
import torch

REL = 1e-5   # north_star: outputs within 1e-5 relative fp32 of the reference semantics


def dev(a, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def npy(t):
    return t.detach().float().cpu().numpy() if t.dtype.is_floating_point else t.detach().cpu().numpy()


def assert_close(got, ref, rel=REL, what="", scale=None):
    """|got - ref| <= rel * max|ref| elementwise (matrix-scale relative error, the meaningful measure
    for accumulated fp32 dot products), plus exact shape equality."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, f"{what}: shape {got.shape} vs {ref.shape}"
    if scale is None:
        scale = float(np.max(np.abs(ref))) if ref.size else 0.0
    scale = max(float(scale), 1e-30)
    err = float(np.max(np.abs(got - ref))) if ref.size else 0.0
    assert np.isfinite(got).all(), f"{what}: non-finite values"
    assert err <= rel * scale + 1e-12, f"{what}: max abs err {err:.3e} > {rel:.0e} * {scale:.3e}"
