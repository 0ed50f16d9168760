#!/usr/bin/env python
"""Phase timing of the row-sharded DCN step (torchrun, N GPUs): where does the multi-GPU step spend its time?"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
world = int(os.environ["WORLD_SIZE"]); rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import keras_rs_b200 as K
from keras_rs_b200._lib import check, lib, ptr, stream
from keras_rs_b200.sharded import ShardedDCN
K.set_gemm_engine(os.environ.get("ENGINE", "tcgen05"))
B, F, V, E = 65536, 26, 1_000_000, 32
m = ShardedDCN([V] * F, rank=rank, world=world, embedding_dim=E, num_cross_layers=3, dense_units=(192, 192), seed=1)
opt = K.optimizers.AdamW(0.01)
g = torch.Generator().manual_seed(rank)
ids = torch.randint(0, V, (B, F), generator=g, dtype=torch.int32).cuda(); y = torch.rand((B,), generator=g).cuda()
for _ in range(3): m.train_on_batch(ids, y, opt, B * world)
torch.cuda.synchronize(); dist.barrier()
def ev(): e = torch.cuda.Event(enable_timing=True); e.record(); return e
b = m._step_buffers(B); s = stream()
R = 10
acc = {}
def add(k, a, c): acc[k] = acc.get(k, 0.0) + a.elapsed_time(c)
for it in range(R):
    e0 = ev(); m._gather_into(b, B, s); e1 = ev()
    m._scatter_from(b, B, b["ga"], s); e2 = ev()
    opt._update(m.emb, m.emb_grad, m.emb_touched); e3 = ev()
    loss = m.forward_backward(ids, y, B * world); e4 = ev()
    torch.cuda.synchronize()
    for k, a, c in (("gather: owner-stage + pull (2 barriers)", e0, e1), ("scatter: push + allreduce + owner scatter + barrier", e1, e2),
                    ("adamw local shard", e2, e3), ("forward_backward total", e3, e4)):
        add(k, a, c)
    dist.barrier()
if rank == 0:
    print(json.dumps({k: round(v / R, 3) for k, v in acc.items()}))
m.close(); dist.destroy_process_group()
