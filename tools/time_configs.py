"""Wall/GPU time of the other BASELINE.json configs through the layers (run on the GPU box)."""
import sys, time
import torch
sys.path.insert(0, ".")
from deepdow_b200 import CovarianceMatrix, NumericalMarkowitz, SparsemaxAllocator, SoftmaxAllocator
dev = "cuda:0"
def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
gen = torch.Generator(device=dev).manual_seed(1)
# C1: N=20, B=64, L=40
x = torch.randn(64, 40, 20, generator=gen, device=dev) * 0.01
cov = CovarianceMatrix(sqrt=True, shrinkage_strategy="diagonal")
C = cov(x); r = torch.randn(64, 20, generator=gen, device=dev) * 0.01; o = torch.ones(64, device=dev)
l = NumericalMarkowitz(20, 1.0)
ins = [t.requires_grad_() for t in (r, C.clone(), o.clone(), o.clone())]; G = torch.randn(64, 20, device=dev)
print(f"C1  N=20 B=64: covariance(sqrt) {timeit(lambda: cov(x)):.3f} ms, markowitz fwd+bwd {timeit(lambda: torch.autograd.grad(l(*ins), ins, G)):.3f} ms")
# C2 shapes: N=50, L=40, B=256 (covariance sqrt=False as BachelierNet + markowitz)
x = torch.randn(256, 40, 50, generator=gen, device=dev) * 0.01
cov2 = CovarianceMatrix(sqrt=False, shrinkage_strategy="diagonal")
C = cov2(x); r = torch.randn(256, 50, generator=gen, device=dev) * 0.01; o = torch.ones(256, device=dev)
l = NumericalMarkowitz(50, 1.0)
ins = [t.requires_grad_() for t in (r, C.clone(), o.clone(), o.clone())]; G = torch.randn(256, 50, device=dev)
print(f"C2  N=50 B=256: covariance {timeit(lambda: cov2(x)):.3f} ms, markowitz fwd+bwd {timeit(lambda: torch.autograd.grad(l(*ins), ins, G)):.3f} ms")
# C3: sparsemax / softmax N=100 B=1024
xs = (torch.rand(1024, 100, generator=gen, device=dev) - 0.5).requires_grad_(); Gs = torch.randn(1024, 100, device=dev)
for name, layer in (("sparsemax", SparsemaxAllocator(100, 1, 0.05)), ("softmax(variational)", SoftmaxAllocator(1, "variational", 100, 0.05))):
    print(f"C3  {name} N=100 B=1024: fwd+bwd {timeit(lambda: torch.autograd.grad(layer(xs), xs, Gs), 20):.4f} ms")
# C5 share: N=500, L=250, B=1024
x = torch.randn(1024, 250, 500, generator=gen, device=dev) * 0.01
cov5 = CovarianceMatrix(sqrt=True, shrinkage_strategy="diagonal")
t_cov = timeit(lambda: cov5(x), 2)
C = cov5(x); r = torch.randn(1024, 500, generator=gen, device=dev) * 0.01; o = torch.ones(1024, device=dev)
l = NumericalMarkowitz(500, 0.05)
ins = [t.requires_grad_() for t in (r, C.clone(), o.clone(), o.clone())]; G = torch.randn(1024, 500, device=dev)
t_f = timeit(lambda: l(*[t.detach() for t in ins]), 2)
t_fb = timeit(lambda: torch.autograd.grad(l(*ins), ins, G), 2)
w = l(*ins); nf = ((w > 0) & (w < 0.05)).sum(1).float().mean().item()
print(f"C5  N=500 L=250 B=1024: covariance(sqrt) {t_cov:.1f} ms, markowitz fwd {t_f:.1f} ms, fwd+bwd {t_fb:.1f} ms (mean free {nf:.1f})")
# C4 as BASELINE.json states it: ThorpNet(n_assets=100, max_weight=0.2), batch 4096 (identical problems; solved one by one, not deduplicated)
from deepdow_b200 import nn as dnn
net = dnn.ThorpNet(100, max_weight=0.2).to(dev)
with torch.no_grad():
    net.exp_returns.normal_(0, 0.1); net.matrix.add_(0.05 * torch.randn(100, 100, device=dev))
opt = torch.optim.Adam(net.parameters(), lr=1e-3)
X4 = torch.randn(4096, 1, 5, 100, device=dev); y4 = torch.randn(4096, 1, 10, 100, device=dev) * 0.01
def thorp_step():
    loss = dnn.sharpe_ratio_loss(net(X4), y4).mean()
    opt.zero_grad(set_to_none=True); loss.backward(); opt.step()
t_thorp = timeit(thorp_step, 10)
net.shared_problem = True            # opt-in: the 4096 identical problems solved once
t_thorp_shared = timeit(thorp_step, 10)
net.shared_problem = False
print(f"C4  ThorpNet N=100 B=4096 max_weight=0.2: training step {t_thorp:.3f} ms; shared_problem=True {t_thorp_shared:.3f} ms")
# C2 as BASELINE.json states it: BachelierNet training (tools/train_bachelier.py has the multi-GPU form)
bn = dnn.BachelierNet(1, 50, hidden_size=32).to(dev); ob = torch.optim.Adam(bn.parameters(), lr=1e-2)
X2 = torch.randn(256, 1, 40, 50, device=dev) * 0.01; y2 = torch.randn(256, 1, 10, 50, device=dev) * 0.01
def bach_step():
    loss = dnn.sharpe_ratio_loss(bn(X2), y2).mean()
    ob.zero_grad(set_to_none=True); loss.backward(); ob.step()
t_b = timeit(bach_step, 10)
bn.fused_covariance = True           # opt-in: covariance formed inside the solver kernel
t_bf = timeit(bach_step, 10)
bn.fused_covariance = False
# the allocation part alone (covariance -> markowitz, fwd + bwd) on the network's shapes, composed vs fused
from deepdow_b200.layers import covariance_markowitz
xs = (torch.randn(256, 40, 50, generator=gen, device=dev) * 0.01).requires_grad_(); rs = (torch.randn(256, 50, generator=gen, device=dev) * 0.01).requires_grad_()
o2 = torch.ones(256, device=dev); G2 = torch.randn(256, 50, device=dev)
cl, al = CovarianceMatrix(sqrt=False), NumericalMarkowitz(50, 1.0)
t_comp = timeit(lambda: torch.autograd.grad(al(rs, cl(xs), o2, o2), (xs, rs), G2), 20)
t_fus = timeit(lambda: torch.autograd.grad(covariance_markowitz(xs, rs, o2, o2, cl, al), (xs, rs), G2), 20)
print(f"C2  BachelierNet N=50 L=40 H=10 B=256: training step {t_b:.3f} ms; fused_covariance=True {t_bf:.3f} ms; "
      f"covariance->markowitz fwd+bwd alone: composed {t_comp:.3f} ms, fused {t_fus:.3f} ms")
