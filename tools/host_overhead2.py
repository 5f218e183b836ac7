import sys, time
import torch
dev = "cuda:0"
x = torch.randn(8192, 8192, device=dev)
ws = torch.zeros(4096, dtype=torch.uint8, device=dev)
buf = torch.zeros(8, dtype=torch.int32).pin_memory()
print("pinned:", buf.is_pinned(), buf[2:3].is_pinned())
def busy():
    for _ in range(3): y = x @ x
torch.cuda.synchronize()
for name, fn in [
    ("slice+view", lambda: ws[8:12].view(torch.int32)),
    ("copy_ slice non_blocking", lambda: buf[2:3].copy_(ws[8:12].view(torch.int32), non_blocking=True)),
    ("copy_ whole non_blocking", lambda: buf.copy_(ws[0:32].view(torch.int32), non_blocking=True)),
    ("to cpu non_blocking", lambda: ws[8:12].view(torch.int32).to("cpu", non_blocking=True)),
]:
    busy(); t0 = time.perf_counter(); fn(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"{name}: call {1e6*(t1-t0):.0f} us (then drain {1e6*(t2-t1):.0f} us)")
