"""Kernel-time table of one training step (torch.profiler, CUDA activities) - where the hand-written step spends its time."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import vfr_b200
from vfr_b200 import models, main as vmain

dev = "cuda"
torch.manual_seed(123)
table = torch.randn(10000, 100) * 0.4; table[0] = 0
g = torch.Generator(device=dev).manual_seed(1)
R_, Bq = 123, 87
maskp = torch.sort(torch.randint(0, Bq, (R_,), device=dev, generator=g)).values
maskp[:Bq] = torch.arange(Bq, device=dev); maskp = torch.sort(maskp).values
model = models.CALModel(visual_input_dim=8194, pretrained_emb=table).to(dev)
opt = vmain.FusedAdam(filter(lambda q: q.requires_grad, model.parameters()), lr=5e-4, weight_decay=5e-3)
batch = {"posit": torch.rand(R_, 8194, device=dev, generator=g), "intra": torch.rand(R_, 8194, device=dev, generator=g),
         "inter": torch.rand(R_, 8194, device=dev, generator=g), "lang": torch.randint(1, 10000, (Bq, 20), device=dev, generator=g),
         "maskp": maskp, "maskn": maskp}
tr = vmain.Trainer(device=dev)
for _ in range(3):
    tr.train_epoch(model, [batch], opt)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        tr.train_epoch(model, [batch], opt)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
