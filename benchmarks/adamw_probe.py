#!/usr/bin/env python
"""AdamW table sweep: effect of the touched-row fraction (gradient arena reads) on the streaming kernel."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
import keras_rs_b200 as K
V, E = 26_000_000, 32
p = torch.rand((V, E), device="cuda"); g = torch.zeros_like(p)
tb = torch.zeros((V // 32,), dtype=torch.int32, device="cuda")
p._krs_arena, p._krs_touched = g, tb
opt = K.optimizers.AdamW(0.01); opt.iterations = 1
gen = torch.Generator(device="cuda").manual_seed(0)
def run(frac, reps=5):
    ts = []
    for r in range(reps + 1):
        tb.zero_()
        if frac > 0:
            rows = torch.randint(0, V, (int(V * frac),), device="cuda", generator=gen)
            words = torch.zeros((V // 32,), dtype=torch.int64, device="cuda")
            words.scatter_reduce_(0, rows // 32, (1 << (rows % 32)).to(torch.int64), reduce="sum", include_self=True)  # approx bitmap (dups may carry; fine for a probe)
            tb.copy_((words & 0xFFFFFFFF).to(torch.int32) if False else torch.zeros_like(tb))
            # exact bitmap via unique rows
            ur = torch.unique(rows)
            w = torch.zeros((V // 32,), dtype=torch.int64, device="cuda")
            w.index_put_((ur // 32,), (torch.ones_like(ur) << (ur % 32)), accumulate=True)
            tb.copy_(torch.where(w >= 2**31, w - 2**32, w).to(torch.int32))
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); opt._update(p, g, tb); b.record(); torch.cuda.synchronize()
        if r: ts.append(a.elapsed_time(b))
    return sum(ts) / len(ts)
res = {f"touched {f:.3f}": round(run(f), 3) for f in (0.0, 0.01, 0.065, 0.25, 1.0)}
print(json.dumps(res))
