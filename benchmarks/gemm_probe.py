#!/usr/bin/env python
"""Times the dense contractions of one C2 cross layer (and the first Dense layer) on each GEMM engine:
CUDA events on the launch stream, 218 MB operands (>> L2), results compared with the exact-fp32 FFMA engine.

  python benchmarks/gemm_probe.py [--engines tcgen05,tcgen05_ts] [--reps 10] [--batch 65536] [--dim 832]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import keras_rs_b200 as K  # noqa: E402
from keras_rs_b200._lib import check, lib, ptr, stream  # noqa: E402


def timed(fn, reps, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--engines", default="tcgen05,tcgen05_ts")
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--dim", type=int, default=832)
    ap.add_argument("--units", type=int, default=192)
    a = ap.parse_args()
    B, D, U = a.batch, a.dim, a.units
    g = torch.Generator(device="cuda").manual_seed(0)
    x0 = torch.randn((B, D), device="cuda", generator=g)
    x1 = torch.randn((B, D), device="cuda", generator=g)
    dz = torch.randn((B, D), device="cuda", generator=g)
    V = (torch.rand((D, D), device="cuda", generator=g) * 2 - 1) * 0.06
    W1 = (torch.rand((D, U), device="cuda", generator=g) * 2 - 1) * 0.06
    bias = torch.zeros(D, device="cuda")
    y, h2 = torch.empty_like(x0), torch.empty_like(x0)
    dV, dx, yd = torch.empty_like(V), torch.empty_like(x0), torch.empty((B, U), device="cuda")
    s = stream()

    ops = {
        "cross_fwd NN (B,D)x(D,D) fused cross epilogue": (lambda: check(lib.krs_cross_fwd(
            ptr(x0), ptr(x1), None, ptr(V), ptr(bias), 0.0, 0, ptr(y), ptr(h2), None, None, B, D, 0, s)), 2.0 * B * D * D, lambda: y),
        "dW TN (D,B)x(B,D) split-K": (lambda: K.ops.sgemm(x1, dz, True, False, out=dV), 2.0 * B * D * D, lambda: dV),
        "dx NT (B,D)x(D,D)^T": (lambda: K.ops.sgemm(dz, V, False, True, out=dx), 2.0 * B * D * D, lambda: dx),
        "dense_fwd NN (B,D)x(D,U) bias+relu": (lambda: check(lib.krs_dense_fwd(
            ptr(x0), ptr(W1), ptr(bias), 1, ptr(yd), B, D, U, s)), 2.0 * B * D * U, lambda: yd),
    }
    K.set_gemm_engine("ffma")
    refs = {}
    for name, (fn, _, out) in ops.items():
        fn()
        torch.cuda.synchronize()
        refs[name] = out().clone()
    for eng in a.engines.split(","):
        K.set_gemm_engine(eng)
        for name, (fn, flops, out) in ops.items():
            before = lib.krs_gemm_tc_launch_count()
            ms = timed(fn, a.reps)
            launched = lib.krs_gemm_tc_launch_count() - before
            err = float((out().double() - refs[name].double()).abs().max() / refs[name].double().abs().max())
            print(json.dumps(dict(engine=eng, fuse_n=os.environ.get("KRS_TC_FUSE_N", "0"), op=name, ms=round(ms, 4),
                                  TFLOPs=round(flops / ms * 1e-9, 1), tc_launches=launched, max_err_vs_ffma=err)), flush=True)


if __name__ == "__main__":
    main()
