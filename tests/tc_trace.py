"""Debug harness: per-role clock64 timeline of CTA 0 of one tcgen05 GEMM (cross forward, C2 shape).
Needs a trace build of the library: touch keras_rs_b200/csrc/gemm_tc.cu && KRS_EXTRA_FLAGS=-DKRS_TC_TRACE=1 bash keras_rs_b200/csrc/build.sh
ENGINE=tcgen05|tcgen05_ts  MODE=cross|cross_noh2|cross_x1|dense|sgemm"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import keras_rs_b200 as K
from keras_rs_b200._lib import lib, check, ptr, stream
ENGINE = os.environ.get("ENGINE", "tcgen05")
K.set_gemm_engine(ENGINE)
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
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record(); torch.cuda.synchronize()
print("ENGINE", ENGINE, "MODE", MODE, "FUSE_N", os.environ.get("KRS_TC_FUSE_N", "0"), "kernel ms", round(e0.elapsed_time(e1), 4))
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
ev.sort(key=lambda e: e[2]); t0 = ev[0][2]
names = {1: "TMA issued", 2: "conv saw full", 9: "conv fence done / (ts) A stage free", 8: "conv stores issued", 3: "conv done",
         4: "mma saw conv", 5: "mma committed", 6: "epi start", 7: "epi done", 15: "mma loop top", 16: "mma wait done"}
print("entries", len(ev))
by = {k: [(i, c - t0) for tag, i, c in ev if tag == k] for k in names}
for k in (6, 7):
    print(names[k], by[k][:6])
KB = 26 if ENGINE == "tcgen05_ts" else 52   # ring stages per tile at K = 832 (tcgen05_ts moves 32 k per stage)
def per_tile(tag, tile):
    """{kb: clock} of this tag's events inside the tile-th visit (events are in time order; kb restarts at 0 per tile)."""
    out, seen_tile, prev = {}, -1, 10 ** 9
    for i, c in by[tag]:
        if i < prev: seen_tile += 1
        prev = i
        if seen_tile == tile: out[i] = c
    return out
def med(a, b2, keys=None):
    keys = sorted(set(a) & set(b2)) if keys is None else keys
    d = [b2[k] - a[k] for k in keys if k in a and k in b2]
    return float(np.median(d)) if d else float("nan")
for tile in (0, 1, 2):
    T = {k: per_tile(k, tile) for k in (1, 2, 9, 8, 3, 4, 5, 15, 16)}
    if len(T[5]) < KB: break
    c5 = [T[5][k] for k in sorted(T[5])]
    print(f"tile {tile}: mainloop {c5[-1] - T[4][0]} clks; k-block period (mma committed) median {np.median(np.diff(c5)):.0f}; "
          f"TMA issue->conv saw full {med(T[1], T[2]):.0f}; conv: full->(9) {med(T[2], T[9]):.0f}, (9)->stores issued {med(T[9], T[8]):.0f}, "
          f"full->stores issued {med(T[2], T[8]):.0f}, stores->done {med(T[8], T[3]):.0f}, full->done {med(T[2], T[3]):.0f}; "
          f"conv done->mma saw {med(T[3], T[4]):.0f}; mma saw->committed {med(T[4], T[5]):.0f}")
    nxt = {k: T[15].get(k + 1) for k in T[5]}
    print(f"   mma warp: loop top->wait done {med(T[15], T[16]):.0f}; wait done->(fence, elect) tag 4 {med(T[16], T[4]):.0f}; issue (4->5) {med(T[4], T[5]):.0f}; "
          f"5->next loop top {np.median([nxt[k] - T[5][k] for k in T[5] if nxt.get(k)]):.0f}")
    ks = sorted(T[2])[:8]
    print("   traced conv k-blocks", ks, " TMA", [T[1].get(k) for k in ks], " full", [T[2].get(k) for k in ks], " convdone", [T[3].get(k) for k in ks],
          " mma saw", [T[4].get(k) for k in ks], " commit", [T[5].get(k) for k in ks])
e10 = [c for tag, i, c in ev if tag == 10][:8]; e11 = [c for tag, i, c in ev if tag == 11][:8]; e12 = [c for tag, i, c in ev if tag == 12][:8]
e13 = [c for tag, i, c in ev if tag == 13][:8]; e14 = [c for tag, i, c in ev if tag == 14][:8]
if len(e13) >= 3:
    print("epilogue chunks (tile 0): loaded->row group 0 done", [b - a for a, b in zip(e11, e13)], " group 0->3", [b - a for a, b in zip(e13, e14)], " group 3->7 (+syncwarp)", [b - a for a, b in zip(e14, e12)])
if len(e12) >= 3:
    print("epilogue chunks (tile 0): begin->tmem loaded", [b - a for a, b in zip(e10, e11)])
    print("                          loaded->chunk done ", [b - a for a, b in zip(e11, e12)])
    print("                          chunk period       ", [int(x) for x in np.diff(np.array(e10))])
