"""One eager DLRM C3 training step between cudaProfilerStart/Stop, for an ncu launch list:
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c3_launches.csv \
      python benchmarks/c3_launches.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import keras_rs_b200 as K  # noqa: E402
from keras_rs_b200.dlrm import DLRM  # noqa: E402

B, F, V, E = 65536, 26, 1_000_000, 128
model = DLRM([V] * F, embedding_dim=E, num_dense=13, bottom_mlp_dims=(512, 256, E), interaction="dot", seed=1234)
opt = K.optimizers.Adagrad(0.01)
g = torch.Generator().manual_seed(1)
ids = torch.randint(0, V, (B, F), generator=g, dtype=torch.int32).cuda()
dense = torch.rand((B, 13), generator=g).cuda()
y = torch.randint(0, 2, (B,), generator=g).float().cuda()
for _ in range(3):
    model.train_on_batch(dense, ids, y, opt)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
model.train_on_batch(dense, ids, y, opt)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
