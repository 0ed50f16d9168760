#!/usr/bin/env python
"""Per-kernel measurements at the BASELINE.json shapes other than the headline step (C3 DotInteraction,
C3 gather E=128, C4 BruteForceRetrieval), CUDA events on the launch stream, inputs >> L2.

  python benchmarks/bench_kernels.py [--what dot,gather128,topk] [--reps 10]
Prints one JSON line per kernel: algorithmic bytes/flops, ms, GB/s or TFLOP/s, roofline frac.
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
    for i in range(warmup):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(warmup + i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}


def bench_dot(reps):
    B, N, E = 65536, 27, 128
    g = torch.Generator(device="cuda").manual_seed(0)
    bufs = [torch.randn((B, N * E), device="cuda", generator=g) for _ in range(3)]       # 3 x 906 MB
    layer = K.layers.DotInteraction()
    outs = []

    def f(i):
        buf = bufs[i % 3]
        outs[:] = [layer([buf[:, j * E:(j + 1) * E] for j in range(N)])]

    ms = timed(f, reps)
    by = B * N * E * 4 + B * 351 * 4
    gout = torch.randn((B, 351), device="cuda", generator=g)
    feats = [bufs[0][:, j * E:(j + 1) * E] for j in range(N)]
    x = [t.detach().requires_grad_(True) for t in feats]

    def fb(i):
        o = layer(x)
        o.backward(gout)

    ms_fb = timed(fb, max(3, reps // 2))
    # the backward kernel alone, through the C ABI (no autograd / allocation in the timed region)
    import ctypes as C
    dbufs = [torch.empty((B, N * E), device="cuda") for _ in range(2)]
    fp = (C.c_void_p * N)(*[t.data_ptr() for t in feats])
    fs = (C.c_int64 * N)(*[N * E] * N)
    dps = [(C.c_void_p * N)(*[d[:, j * E:(j + 1) * E].data_ptr() for j in range(N)]) for d in dbufs]
    ms_b = timed(lambda i: check(lib.krs_dot_bwd(fp, fs, ptr(gout), dps[i % 2], fs, N, E, B, 0, 0, stream())), reps)
    by_b = 2 * B * N * E * 4 + B * 351 * 4                 # F read once, dF written once, G read
    return [dict(kernel="dot_fwd_mma_kernel", config="C3 N=27 E=128 B=65536 -> 351", ms=ms, GBps=by / ms * 1e-6, bytes=by,
                 frac_of_measured_hbm=by / ms * 1e-6 / peaks()["hbm_gbs"]),
            dict(kernel="dot_bwd_mma_kernel", config="C3 (C ABI call)", ms=ms_b, GBps=by_b / ms_b * 1e-6, bytes=by_b,
                 frac_of_measured_hbm=by_b / ms_b * 1e-6 / peaks()["hbm_gbs"]),
            dict(kernel="dot fwd+bwd (autograd path, incl. host overhead of 27 strided views)", config="C3", ms=ms_fb)]


def bench_gather128(reps):
    F, V, E, B = 26, 1_000_000, 128, 65536
    g = torch.Generator(device="cuda").manual_seed(1)
    arena = torch.rand((F * V, E), device="cuda", generator=g)       # 13.3 GB
    out = torch.empty((B, F * E), device="cuda")
    res = []
    for dt in (torch.int32, torch.int64):
        ids = [torch.randint(0, V, (B, F), device="cuda", generator=g).to(dt) for _ in range(3)]
        plans = [K.ops.GatherPlan([dict(table=arena[f * V:(f + 1) * V], ids=i[:, f], combiner="sum") for f in range(F)]) for i in ids]
        ms = timed(lambda i: plans[i % 3].forward(out), reps)
        by = B * F * E * 4 * 2 + B * F * (8 if dt == torch.int64 else 4)
        res.append(dict(kernel="gather_fast_kernel", config=f"C3 gather F=26 E=128 B=65536 ids={dt}", ms=ms, GBps=by / ms * 1e-6,
                        bytes=by, frac_of_measured_hbm=by / ms * 1e-6 / peaks()["hbm_gbs"], frac_of_8TBps=by / ms * 1e-6 / 8000))
    return res


# examples/ml_perf/configs/v6e_8.py:15-172: vocabulary sizes and feature_list_length (hotness) of the 26 sparse features
MLPERF_VOCAB = [40000000, 39060, 17295, 7424, 20265, 3, 7122, 1543, 63, 40000000, 3067956, 405282, 10, 2209, 11938, 155, 4, 976, 14,
                40000000, 40000000, 40000000, 590152, 12973, 108, 36]
MLPERF_HOT = [3, 2, 1, 2, 6, 1, 1, 1, 1, 7, 3, 8, 1, 6, 9, 5, 1, 1, 1, 12, 100, 27, 10, 3, 1, 1]


def bench_multihot(reps):
    """Fused multi-table MULTI-HOT gather on the ml_perf feature list (sum of hotness 214, combiner sum, int64 ids as
    examples/ml_perf/dataloader.py:93-98 produces them), E = 128, batch 65536.  Vocabularies above 2e6 rows are capped at
    2e6 (the 40M-row tables would need 20 GB each); rows stay far larger than L2 either way."""
    E, B = 128, 65536
    g = torch.Generator(device="cuda").manual_seed(2)
    vocab = [min(v, 2_000_000) for v in MLPERF_VOCAB]
    tables = [torch.rand((v, E), device="cuda", generator=g) for v in vocab]
    res = []
    feats = []
    for f, (v, h) in enumerate(zip(vocab, MLPERF_HOT)):
        ids = torch.randint(0, v, (B, h), device="cuda", generator=g)            # int64
        feats.append(dict(table=tables[f], ids=ids if h > 1 else ids[:, 0], combiner="sum"))
    plan = K.ops.GatherPlan(feats)
    out = torch.empty((B, len(vocab) * E), device="cuda")
    ms = timed(lambda i: plan.forward(out), reps)
    n_lookups = B * sum(MLPERF_HOT)
    by = n_lookups * E * 4 + B * len(vocab) * E * 4 + n_lookups * 8
    res.append(dict(kernel="gather_sample_kernel (multi-hot)", config=f"ml_perf 26 features, sum(H)=214, E=128, B={B}, int64 ids, sum",
                    ms=ms, GBps=by / ms * 1e-6, bytes=by, frac_of_measured_hbm=by / ms * 1e-6 / peaks()["hbm_gbs"]))
    # backward (scatter-add of the (B, F*E) gradient into the 214 rows of every sample)
    grads = [torch.zeros_like(t) for t in tables]
    touched = [torch.zeros(((t.shape[0] + 31) // 32,), dtype=torch.int32, device="cuda") for t in tables]
    gout = torch.randn((B, len(vocab) * E), device="cuda", generator=g)
    ms_b = timed(lambda i: plan.backward(gout, grads, touched), max(3, reps // 2))
    by_b = B * len(vocab) * E * 4 + 2 * n_lookups * E * 4 + n_lookups * 8
    res.append(dict(kernel="scatter_sample_kernel (multi-hot)", config="same", ms=ms_b, GBps=by_b / ms_b * 1e-6, bytes=by_b,
                    frac_of_measured_hbm=by_b / ms_b * 1e-6 / peaks()["hbm_gbs"]))
    return res


def bench_topk(reps):
    nq, nc, d, k = 4096, 10_000_000, 64, 100
    g = torch.Generator(device="cuda").manual_seed(42)
    C = torch.randn((nc, d), device="cuda", generator=g)
    Q = torch.randn((nq, d), device="cuda", generator=g)
    layer = K.layers.BruteForceRetrieval(candidate_embeddings=C, k=k)
    ms = timed(lambda i: layer(Q), max(2, reps // 3), warmup=1)
    fl = 2.0 * nq * nc * d
    s, i = layer(Q)
    # spot check 8 queries against a float64 reference on the device
    ref = (Q[:8].double() @ C.double().T)
    rs, ri = torch.topk(ref, k, dim=1)
    ok = bool(torch.allclose(s[:8].double(), rs, atol=1e-3)) and bool((torch.gather(ref, 1, i[:8].long()) - rs).abs().max() < 1e-3)
    return [dict(kernel="topk_partial+merge", config="C4 Q=4096x64 C=1e7x64 k=100", ms=ms, TFLOPs=fl / ms * 1e-9, flops=fl,
                 queries_per_s=nq / ms * 1e3, spot_check_ok=ok)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--what", default="dot,gather128,topk")
    ap.add_argument("--reps", type=int, default=10)
    a = ap.parse_args()
    fns = {"dot": bench_dot, "gather128": bench_gather128, "topk": bench_topk, "multihot": bench_multihot}
    for w in a.what.split(","):
        for r in fns[w](a.reps):
            print(json.dumps(r), flush=True)
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
