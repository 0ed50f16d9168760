"""Debug harness (not a pytest): repeat the individual GEMMs of the low-rank cross backward on the
tcgen05 engine and report the worst error per GEMM, to localise intermittent mismatches."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import keras_rs_b200 as K

K.set_gemm_engine("tcgen05")
g = torch.Generator(device="cuda").manual_seed(0)
B, D, P = 640, 832, 64
R = int(os.environ.get("REPS", "30"))

def rel(got, ref):
    return float((got.double() - ref).abs().max() / ref.abs().max())

cases = {
  "dh=dz@V^T  (M640 N64 K832 NT)": lambda dz, V, U, dh, x: (K.ops.sgemm(dz, V, False, True), dz.double() @ V.double().T),
  "dx=dh@U^T  (M640 N832 K64 NT)": lambda dz, V, U, dh, x: (K.ops.sgemm(dh, U, False, True), dh.double() @ U.double().T),
  "dU=x^T@dh  (M832 N64 K640 TN)": lambda dz, V, U, dh, x: (K.ops.sgemm(x, dh, True, False), x.double().T @ dh.double()),
  "dV=h^T@dz  (M64 N832 K640 TN)": lambda dz, V, U, dh, x: (K.ops.sgemm(dh, dz, True, False), dh.double().T @ dz.double()),
  "h=x@U      (M640 N64 K832 NN)": lambda dz, V, U, dh, x: (K.ops.sgemm(x, U), x.double() @ U.double()),
  "y=h@V      (M640 N832 K64 NN)": lambda dz, V, U, dh, x: (K.ops.sgemm(dh, V), dh.double() @ V.double()),
}
worst = {k: 0.0 for k in cases}
bad = {k: 0 for k in cases}
for r in range(R):
    dz = torch.randn((B, D), device="cuda", generator=g)
    V = torch.randn((P, D), device="cuda", generator=g) * 0.05
    U = torch.randn((D, P), device="cuda", generator=g) * 0.05
    dh = torch.randn((B, P), device="cuda", generator=g)
    x = torch.randn((B, D), device="cuda", generator=g)
    for name, fn in cases.items():
        got, ref = fn(dz, V, U, dh, x)
        e = rel(got, ref)
        worst[name] = max(worst[name], e)
        bad[name] += e > 1e-5
for k in cases:
    print(f"{k}: worst rel err {worst[k]:.2e}  bad {bad[k]}/{R}")
