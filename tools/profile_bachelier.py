"""Where a BachelierNet training step spends its GPU time (torch profiler, kernel totals): python tools/profile_bachelier.py"""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepdow_b200 import nn as dnn  # noqa: E402

dev = "cuda:0"
torch.manual_seed(0)
net = dnn.BachelierNet(1, 50, hidden_size=32).to(dev)
opt = torch.optim.Adam(net.parameters(), lr=1e-2)
X = torch.randn(256, 1, 40, 50, device=dev) * 0.01
y = torch.randn(256, 1, 10, 50, device=dev) * 0.01


def step():
    loss = dnn.sharpe_ratio_loss(net(X), y).mean()
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()


for _ in range(5):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(10):
        step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
