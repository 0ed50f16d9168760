#!/usr/bin/env python
"""Probe (torchrun, 2 GPUs): IPC-exported vs torch-allocated arenas, local vs remote, random vs sequential rows."""
import os, sys, json, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
world = int(os.environ["WORLD_SIZE"]); rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import keras_rs_b200 as K
from keras_rs_b200 import _lib as L
from keras_rs_b200._lib import check, lib, ptr, stream
from keras_rs_b200.sharded import _ipc_tensor
V, E, B, F = 13_000_000, 32, 65536, 26
def timed(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
res = {}
t_arena = torch.rand((V, E), device="cuda")
i_arena, h, p = _ipc_tensor((V, E), torch.float32)
i_arena.copy_(t_arena)
opt = K.optimizers.AdamW(0.01); opt.iterations = 1
for name, a in (("torch", t_arena), ("ipc", i_arena)):
    gr = torch.zeros_like(a); tb = torch.zeros((V // 32 + 1,), dtype=torch.int32, device="cuda")
    a._krs_arena, a._krs_touched = gr, tb
    res[f"adamw_{name}_ms"] = timed(lambda: opt._update(a, gr, tb))
    del gr, tb
    opt._state.clear()
# exchange handles
allh = [None] * world
dist.all_gather_object(allh, h)
peer = (rank + 1) % world
q = C.c_void_p(); hb = (C.c_ubyte * 64).from_buffer_copy(allh[peer]); check(lib.krs_ipc_open(hb, C.byref(q)))
out = torch.empty((B, F * E), device="cuda")
g = torch.Generator(device="cuda").manual_seed(rank)
def make_plan(ids, base_ptr):
    plan = K.ops.GatherPlan([dict(table=t_arena[:1], ids=ids[:, f], combiner="sum") for f in range(F)])
    for f in range(F):
        d = plan.arr[f]; d.table = base_ptr; d.vocab = V; d.dim = E; d.out_offset = f * E
    plan.out_dim = F * E
    return plan
rnd = torch.randint(0, V, (B, F), device="cuda", generator=g, dtype=torch.int32)
seq = (torch.arange(B * F, device="cuda", dtype=torch.int32).reshape(B, F)) % V
by = B * F * E * 4 * 2 + B * F * 4
for name, base in (("local_torch", t_arena.data_ptr()), ("local_ipc", i_arena.data_ptr()), ("remote_ipc", q.value)):
    for idn, ids in (("rand", rnd), ("seq", seq)):
        plan = make_plan(ids, base)
        ms = timed(lambda: plan.forward(out))
        res[f"gather_{name}_{idn}"] = f"{ms:.3f} ms ({by / ms * 1e-6:.0f} GB/s)"
    dist.barrier()
# remote streaming copy for reference
dst = torch.empty((V // 4, E), device="cuda")
src = torch.as_tensor(type("R", (), {"__cuda_array_interface__": {"shape": (V // 4, E), "typestr": "<f4", "data": (q.value, False), "version": 3, "strides": None}})(), device="cuda")
ms = timed(lambda: dst.copy_(src)); res["remote_stream_copy"] = f"{ms:.3f} ms ({dst.numel() * 4 / ms * 1e-6:.0f} GB/s read)"
dist.barrier()
if rank == 0: print(json.dumps(res, indent=1))
lib.krs_ipc_close(q); dist.destroy_process_group()
