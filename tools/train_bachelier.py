"""BachelierNet training throughput (BASELINE.json config 2: n_assets=50, lookback=40, horizon=10, batch=256 per GPU,
SharpeRatio loss, Adam lr 1e-2 as deepdow/experiments.py:297), synthetic returns ~ N(0, 0.01^2) as tests/conftest.py:89-93.

    python tools/train_bachelier.py [--steps 30] [--batch 256] [--loop]            # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/train_bachelier.py                                                     # N GPUs, weak scaling

One process per GPU; every rank trains on its own 256 samples; the only exchange is the flat-bucket all-reduce of the
parameter gradients (deepdow_b200.distributed.GradAllReducer, NCCL).  --loop runs the reference's formulation of the
RNN / attention (a Python loop over the samples, transform.py:151-159, collapse.py:46-61) on the same CUDA layers, to show
what batching those callers is worth.  Prints one JSON line (rank 0): samples/s over all ranks, max-over-ranks device time.
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepdow_b200 import nn as dnn                          # noqa: E402
from deepdow_b200.distributed import GradAllReducer         # noqa: E402


class LoopRNN(dnn.RNN):
    def forward(self, x):
        xs = x.permute(0, 2, 3, 1)
        return torch.stack([self.cell(xs[i])[0].permute(2, 0, 1) for i in range(len(x))])


class LoopAttention(dnn.AttentionCollapse):
    def forward(self, x):
        out = []
        for i in range(len(x)):
            s = x[i].permute(2, 1, 0)
            wgt = torch.softmax(self.context_vector(self.affine(s)), dim=1)
            out.append((s * wgt).mean(dim=1).permute(1, 0))
        return torch.stack(out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--n-assets", type=int, default=50)
    ap.add_argument("--lookback", type=int, default=40)
    ap.add_argument("--horizon", type=int, default=10)
    ap.add_argument("--loop", action="store_true")
    ap.add_argument("--no-cudnn", action="store_true", help="run the recurrent cell on ATen's own kernels instead of cuDNN")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if a.no_cudnn:
        torch.backends.cudnn.enabled = False
    dev = f"cuda:{local}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    torch.manual_seed(0)                                      # identical initial parameters on every rank
    net = dnn.BachelierNet(1, a.n_assets, hidden_size=32, max_weight=1, shrinkage_strategy="diagonal", p=0.5)
    if a.loop:
        net.transform_layer.__class__ = LoopRNN
        net.time_collapse_layer.__class__ = LoopAttention
    net = net.to(dev)
    opt = torch.optim.Adam(net.parameters(), lr=1e-2)
    reduce_grads = GradAllReducer(net.parameters())
    gen = torch.Generator(device=dev).manual_seed(2000 + rank)
    X = torch.randn(a.batch, 1, a.lookback, a.n_assets, generator=gen, device=dev) * 0.01
    y = torch.randn(a.batch, 1, a.horizon, a.n_assets, generator=gen, device=dev) * 0.01

    def step():
        w = net(X)
        loss = dnn.sharpe_ratio_loss(w, y).mean()
        opt.zero_grad(set_to_none=True)
        loss.backward()
        reduce_grads()
        opt.step()
        return loss

    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        per = ms.item() / a.steps
        print(json.dumps({"metric": "BachelierNet training samples/s", "value": world * a.batch / (per * 1e-3), "unit": "samples/s",
                          "n_gpus": world, "ms_per_step": per, "steps": a.steps, "scaling": "weak",
                          "config": {"n_assets": a.n_assets, "lookback": a.lookback, "horizon": a.horizon, "batch_per_gpu": a.batch,
                                     "callers": "python loop over samples (reference formulation)" if a.loop else "batched RNN / attention",
                                     "cudnn": not a.no_cudnn},
                          "final_loss": float(loss.detach())}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
