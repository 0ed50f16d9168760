"""Debug harness: per-role clock64 timeline of CTA 0 of one tcgen05 GEMM (cross forward, C2 shape)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import keras_rs_b200 as K
from keras_rs_b200._lib import lib, check, ptr, stream
K.set_gemm_engine(os.environ.get("ENGINE", "tcgen05"))
B, D = 65536, 832
g = torch.Generator(device="cuda").manual_seed(0)
x0 = torch.randn((B, D), device="cuda", generator=g); V = torch.randn((D, D), device="cuda", generator=g) * 0.03
b = torch.zeros(D, device="cuda"); y = torch.empty_like(x0); h2 = torch.empty_like(x0)
MODE = os.environ.get("MODE", "cross")
x1 = torch.randn((B, D), device="cuda", generator=g)
def run():
    if MODE == "cross":
        check(lib.krs_cross_fwd(ptr(x0), ptr(x0), None, ptr(V), ptr(b), 0.0, 0, ptr(y), ptr(h2), None, None, B, D, 0, stream()))
    elif MODE == "cross_noh2":
        check(lib.krs_cross_fwd(ptr(x0), ptr(x0), None, ptr(V), ptr(b), 0.0, 0, ptr(y), None, None, None, B, D, 0, stream()))
    elif MODE == "cross_x1":   # x != x0: the epilogue's second stream is NOT what TMA just read
        check(lib.krs_cross_fwd(ptr(x0), ptr(x1), None, ptr(V), ptr(b), 0.0, 0, ptr(y), ptr(h2), None, None, B, D, 0, stream()))
    elif MODE == "dense":
        check(lib.krs_dense_fwd(ptr(x0), ptr(V), ptr(b), 0, ptr(y), B, D, D, stream()))
    else:
        K.ops.sgemm(x0, V, out=y)
for _ in range(2): run()
torch.cuda.synchronize()
tr = torch.zeros(16001 + 16, dtype=torch.int64, device="cuda")
lib.krs_gemm_tc_set_trace(tr.data_ptr())
run(); torch.cuda.synchronize()
lib.krs_gemm_tc_set_trace(None)
t = tr.cpu().numpy()
ev = []
for role in range(4):
    for i in range(2000):
        a, c = int(t[1 + role * 4000 + 2 * i]), int(t[2 + role * 4000 + 2 * i])
        if c: ev.append((a >> 32, a & 0xffffffff, c))
n = len(ev)
ev.sort(key=lambda e: e[2]); t0 = ev[0][2]
names = {8: "conv stores issued", 9: "conv fence done", 1: "TMA issued", 2: "conv saw full", 3: "conv done", 4: "mma saw conv", 5: "mma committed", 6: "epi start", 7: "epi done"}
print("MODE", MODE, "entries", n)
by = {k: [(i, c - t0) for tag, i, c in ev if tag == k] for k in names}
for k in (6, 7):
    print(names[k], by[k][:6])
# per k-block deltas for the first tile (52 k-blocks) and the second
for tile in (0, 1, 2):
    lo, hi = tile * 52, tile * 52 + 52
    def seq(k): return [c for (i, c) in by[k][lo:hi]]
    s1, s2, s3, s4, s5 = seq(1), seq(2), seq(3), seq(4), seq(5)
    s8, s9 = seq(8), seq(9)
    if len(s8) == 52:
        print(f"   conv breakdown: full->stores issued {np.median(np.array(s8)-np.array(s2)):.0f}; stores->fence done {np.median(np.array(s9)-np.array(s8)):.0f}; fence->arrive {np.median(np.array(s3)-np.array(s9)):.0f}")
    if len(s5) < 52: break
    d = lambda a: np.diff(np.array(a))
    print(f"tile {tile}: mainloop {s5[-1]-s4[0]} clks; per-kb period mma-committed median {np.median(d(s5)):.0f}; TMA-issue period {np.median(d(s1)):.0f}; "
          f"TMA issue->full seen median {np.median(np.array(s2)-np.array(s1)):.0f}; conv time median {np.median(np.array(s3)-np.array(s2)):.0f}; "
          f"conv done->mma saw {np.median(np.array(s4)-np.array(s3)):.0f}; mma issue {np.median(np.array(s5)-np.array(s4)):.0f}")
    print("   first 8 kb: TMA", [c - s1[0] for c in s1[:8]], " full", [c - s1[0] for c in s2[:8]], " convdone", [c - s1[0] for c in s3[:8]], " commit", [c - s1[0] for c in s5[:8]])

# epilogue chunk breakdown (first tile)
e10 = [c for tag, i, c in ev if tag == 10][:8]; e11 = [c for tag, i, c in ev if tag == 11][:8]; e12 = [c for tag, i, c in ev if tag == 12][:8]
if len(e12) >= 4:
    print("epilogue chunks (tile 0): begin->tmem loaded", [b - a for a, b in zip(e10, e11)])
    print("                          loaded->chunk done ", [b - a for a, b in zip(e11, e12)])
    print("                          chunk period       ", list(np.diff(np.array(e10))))
