#!/usr/bin/env python
"""C4 BruteForceRetrieval timing only (no float64 spot check): python benchmarks/topk_probe.py [--engines tcgen05,ffma]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import keras_rs_b200 as K

ap = argparse.ArgumentParser()
ap.add_argument("--engines", default="tcgen05")
ap.add_argument("--nc", type=int, default=10_000_000)
ap.add_argument("--nq", type=int, default=4096)
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
g = torch.Generator(device="cuda").manual_seed(42)
C = torch.randn((a.nc, 64), device="cuda", generator=g)
Q = torch.randn((a.nq, 64), device="cuda", generator=g)
for eng in a.engines.split(","):
    K.ops.set_topk_engine(eng)
    K.ops.top_k_scores(Q, C, None, 100)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.reps):
        s, i = K.ops.top_k_scores(Q, C, None, 100)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.reps
    print(json.dumps(dict(engine=eng, nq=a.nq, nc=a.nc, ms=round(ms, 3), TFLOPs=round(2.0 * a.nq * a.nc * 64 / ms * 1e-9, 1),
                          queries_per_s=round(a.nq / ms * 1e3))), flush=True)
