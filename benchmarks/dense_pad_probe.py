#!/usr/bin/env python
"""Dense GEMM with an input width that is not a multiple of 4 floats (DLRM top MLP: 128 + 351 = 479) vs the padded width."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
import keras_rs_b200 as K
from keras_rs_b200._lib import check, lib, ptr, stream
K.set_gemm_engine("tcgen05_ts")
B, N = 65536, 1024
def timed(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
for Kd in (479, 480):
    x = torch.randn((B, Kd), device="cuda"); W = torch.randn((Kd, N), device="cuda") * 0.05; b = torch.zeros((N,), device="cuda")
    y = torch.empty((B, N), device="cuda"); gy = torch.randn((B, N), device="cuda")
    dx, dW, db, dz = torch.empty_like(x), torch.empty_like(W), torch.empty_like(b), torch.empty_like(gy)
    n0 = lib.krs_gemm_tc_launch_count()
    f = timed(lambda: check(lib.krs_dense_fwd(ptr(x), ptr(W), ptr(b), 1, ptr(y), B, Kd, N, stream())))
    bw = timed(lambda: check(lib.krs_dense_bwd(ptr(gy), ptr(x), ptr(W), ptr(y), 1, ptr(dx), ptr(dW), ptr(db), ptr(dz), B, Kd, N, stream())))
    print(json.dumps(dict(K=Kd, fwd_ms=round(f, 3), bwd_ms=round(bw, 3), fwd_TFLOPs=round(2.0 * B * Kd * N / f * 1e-9, 1),
                          tc_launches=int(lib.krs_gemm_tc_launch_count() - n0))))
