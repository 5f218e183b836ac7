#!/usr/bin/env python
"""Benchmark of the allocation hot path on B200: NumericalMarkowitz forward+backward solves/sec.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json metric, SURVEY.md section 8(d), "C4 distinct"): B = 4096 independent
portfolio problems per GPU, n_assets = 100, max_weight = 0.2, covmat_sqrt_b = R_b R_b'/N + 0.1 I,
rets ~ N(0, 0.5^2), gamma_sqrt = alpha = 1, upstream gradient ~ N(0, 1), float32, synthetic.
A step = one forward + one backward of the NumericalMarkowitz layer over the whole batch.

Prints ONE JSON line (rank 0).  `value` is solves/s with the inputs resident in HBM, `e2e` the
same through the layer API from pinned host buffers (H2D of every input and D2H of every output
inside the timed region).  `roofline` is for the dominant kernel (the forward solver) against the
measured HBM peak; `cpu_baseline` is the float64 oracle port on the host cores.

Extra keys of the line (all measured in the same run, on the same GPUs):
  f64            the same step in float64 (the reference solves in float64 internally)
  dense_regime   the same matrices with expected returns scaled down 25x (~86 of 100 assets free: the regime the networks
                 train in), CUDA-graph replays, with its own roofline fraction
  strong_scaling the SAME global batch of 4096 problems split over the N ranks (value = 4096 K / max-over-ranks time)
  bachelier_train  BASELINE config 2: BachelierNet training steps (256 samples per GPU, n_assets=50, lookback=40,
                 horizon=10, SharpeRatio loss, Adam) with the NCCL all-reduce of the parameter gradients INSIDE the
                 timed region (north_star: "near-linear 1->8 GPU scaling of BachelierNet training throughput")

`--impl reference` times the reference arm: the reference's cvxpylayers/SCS path is not installable
here, so it is the CPU oracle port (oracle/ddw_oracle.c, OpenMP, all host
threads) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

B_PER_GPU = 4096
N_ASSETS = 100
MAX_WEIGHT = 0.2
METRIC = "NumericalMarkowitz fwd+bwd solves/sec (B=4096,N=100)"
UNIT = "solves/s"
WORKLOAD = ("C4-distinct: NumericalMarkowitz fwd+bwd, B=4096 per GPU, n_assets=100, max_weight=0.2, "
            "covmat_sqrt=R R'/N+0.1I, rets~N(0,0.25), gamma_sqrt=alpha=1")


FWD_KERNELS = ["markowitz_gram_small_kernel", "markowitz_qwarp_kernel"]


def make_inputs_numpy(B, N, seed):
    import numpy as np
    rng = np.random.default_rng(seed)
    R = rng.standard_normal((B, N, N)).astype(np.float32)
    C = R @ R.transpose(0, 2, 1) / N + 0.1 * np.eye(N, dtype=np.float32)
    rets = (rng.standard_normal((B, N)) * 0.5).astype(np.float32)
    G = rng.standard_normal((B, N)).astype(np.float32)
    return rets, C.astype(np.float32), np.ones(B, np.float32), np.ones(B, np.float32), G


def host_threads():
    """Every core this process may run on (torchrun exports OMP_NUM_THREADS=1, which would pin the CPU arm to one)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def time_cpu_oracle(sample, repeats, seed=4000, nthreads=None, budget_s=None):
    """Oracle port (float64, OpenMP) forward+backward on `sample` problems; returns (solves/s, threads, seconds).
    With `budget_s` the number of repeats is chosen so that about that many seconds of CPU work are timed."""
    import numpy as np
    import oracle
    nthreads = host_threads() if nthreads is None else nthreads
    rets, C, g, a, G = make_inputs_numpy(sample, N_ASSETS, seed)
    rets, C, g, a, G = [x.astype(np.float64) for x in (rets, C, g, a, G)]
    oracle.markowitz_fwd(rets[:8], C[:8], g[:8], a[:8], MAX_WEIGHT, nthreads)  # load + warm
    times = []
    if budget_s is not None:
        t0 = time.perf_counter()
        w, st, kkt, status = oracle.markowitz_fwd(rets, C, g, a, MAX_WEIGHT, nthreads)
        oracle.markowitz_bwd(G, w, st, C, g, a, MAX_WEIGHT, nthreads)
        repeats = max(3, min(200, int(budget_s / max(time.perf_counter() - t0, 1e-3))))
    for _ in range(repeats):
        t0 = time.perf_counter()
        w, st, kkt, status = oracle.markowitz_fwd(rets, C, g, a, MAX_WEIGHT, nthreads)
        oracle.markowitz_bwd(G, w, st, C, g, a, MAX_WEIGHT, nthreads)
        times.append(time.perf_counter() - t0)
    return sample / statistics.median(times), nthreads, sum(times)


def run_reference(args, rank, world):
    """Reference arm: CPU oracle port on the host cores (rank 0 only)."""
    if rank != 0:
        return
    sample = B_PER_GPU
    oracle_warm = time_cpu_oracle(64, 1)
    del oracle_warm
    import numpy as np
    import oracle
    rets, C, g, a, G = [x.astype(np.float64) for x in make_inputs_numpy(sample, N_ASSETS, 4000)]
    cores = host_threads()
    per_step = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        w, st, kkt, status = oracle.markowitz_fwd(rets, C, g, a, MAX_WEIGHT, cores)
        oracle.markowitz_bwd(G, w, st, C, g, a, MAX_WEIGHT, cores)
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            per_step.append(dt)
    total = sum(per_step)
    value = sample * len(per_step) / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(per_step),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": f"all {sample} problems per step, fwd+bwd"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} problems fwd+bwd per step, float64 oracle port (cvxpylayers/SCS not installable)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "host_cpu_count": os.cpu_count(),
    }
    print(json.dumps(line), flush=True)


def bind_to_gpu_numa(index):
    """Pin this rank to the CPUs of its GPU's NUMA node before any pinned buffer is allocated, so that the staging buffers of
    the end-to-end path are first-touched (hence placed) next to the GPU's PCIe root instead of wherever the launcher ran.
    Eight ranks pushing 2 x 167 MB per step through one host are bound by the host memory system, not by PCIe."""
    try:
        import torch
        prop = torch.cuda.get_device_properties(index)
        bus = f"{prop.pci_domain_id:04x}:{prop.pci_bus_id:02x}:{prop.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bus}/local_cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        if not allowed:
            return f"GPU {index}: local cpulist {spec} not in this process's affinity mask; default placement"
        os.sched_setaffinity(0, allowed)
        return f"rank bound to the {len(allowed)} CPUs local to GPU {index} ({spec}) before allocating pinned memory"
    except Exception as exc:  # noqa: BLE001 - topology files missing in a container: keep the default placement
        return f"default placement ({type(exc).__name__}: {exc})"


class ClockSampler:
    """nvidia-smi sampled every 200 ms during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, smax, reasons = [], None, set()
        for ln in out.strip().splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax = float(f[2])
            except ValueError:
                continue
            for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--chunk", type=int, default=128, help="problems per stage of the host-buffer pipeline (e2e)")
    ap.add_argument("--status-mode", default="lazy", choices=["lazy", "sync", "off"],
                    help="NumericalMarkowitz.check_status: lazy (default of the layer), sync (read the failure "
                         "counter inside every forward: a host sync) or off (diagnostic)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from deepdow_b200 import NumericalMarkowitz, _C

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    _C.lib()  # fail loudly if the extension is not built
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa_note = bind_to_gpu_numa(local_rank) if world > 1 else "single process: default placement"
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner to stdout at the first collective; the contract is ONE JSON line there
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    B, N, u = B_PER_GPU, N_ASSETS, MAX_WEIGHT
    gen = torch.Generator(device=dev).manual_seed(1000 * 4 + rank)
    R = torch.randn(B, N, N, generator=gen, device=dev)
    C = (R @ R.transpose(1, 2) / N + 0.1 * torch.eye(N, device=dev)).contiguous()
    del R
    rets = torch.randn(B, N, generator=gen, device=dev) * 0.5
    gam = torch.ones(B, device=dev)
    alp = torch.ones(B, device=dev)
    G = torch.randn(B, N, generator=gen, device=dev)
    layer = NumericalMarkowitz(N, max_weight=u)
    layer.check_status = {"lazy": "lazy", "sync": True, "off": False}[args.status_mode]
    inputs = [t.requires_grad_() for t in (rets, C, gam, alp)]

    def step_device(ev=None):
        if ev: ev[0].record()
        w = layer(*inputs)
        if ev: ev[1].record()
        grads = torch.autograd.grad(w, inputs, G)
        if ev: ev[2].record()
        return w, grads

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_device()
    barrier()
    # active-set profile of the workload (work depends on it)
    with torch.no_grad():
        w0 = layer(*[t.detach() for t in inputs])
        nfree = float(((w0 > 0) & (w0 < u)).sum(1).float().mean())
        ncap = float((w0 >= u).sum(1).float().mean())
        nfallback = int((layer.last_status & 1).ne(0).sum())
        ndense = layer.solver_counters()[1]
    barrier()

    # ---- timed region 1: inputs resident in HBM ------------------------------------------------
    sampler = ClockSampler(local_rank)
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    cpu0 = time.perf_counter()
    t_start.record()
    for i in range(args.steps):
        step_device(evs[i])
    t_end.record()
    cpu_enqueue_ms = 1e3 * (time.perf_counter() - cpu0) / args.steps   # host time to enqueue one step
    barrier()
    elapsed_ms = t_start.elapsed_time(t_end)
    fwd_ms = statistics.mean(e[0].elapsed_time(e[1]) for e in evs)
    bwd_ms = statistics.mean(e[1].elapsed_time(e[2]) for e in evs)

    # the same step captured once into a CUDA graph and replayed: the layer's fwd+bwd is a fixed sequence of 6
    # kernels + 3 memsets, and at ~0.6 ms per step the Python side (autograd engine, ctypes, allocator: ~0.3 ms)
    # is of the same order as the kernels.  `value` is the replayed step; the eager loop is reported beside it.
    graph_ms, fwd_graph_ms, graph_note = None, None, "eager launches (graph capture unavailable)"
    rounds, timed_region_s = 1, None
    try:
        layer.check_status = False              # the status copy to pinned memory + event is not capturable
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                step_device()
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            w_g, grads_g = step_device()
        for _ in range(args.warmup):
            graph.replay()
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(args.steps):
            graph.replay()
        g1.record()
        barrier()
        pilot_ms = g0.elapsed_time(g1)
        # `value` = median over ROUNDS of exactly `steps` replays each; enough rounds that the timed region lasts >= 1 s, so
        # that the 200 ms clock sampler sees it (every rank runs the same number of rounds)
        pil = torch.tensor([pilot_ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(pil, op=dist.ReduceOp.MAX)
        rounds = int(min(400, max(3, -(-1000.0 // max(float(pil), 1e-3)))))
        rev = [torch.cuda.Event(enable_timing=True) for _ in range(rounds + 1)]
        barrier()
        rev[0].record()
        for r in range(rounds):
            for _ in range(args.steps):
                graph.replay()
            rev[r + 1].record()
        barrier()
        round_ms = torch.tensor([rev[r].elapsed_time(rev[r + 1]) for r in range(rounds)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(round_ms, op=dist.ReduceOp.MAX)      # every round: the slowest rank
        graph_ms = float(round_ms.median())
        timed_region_s = float(round_ms.sum()) * 1e-3
        # forward alone, replayed the same way: the per-kernel times behind `roofline` must not contain host gaps
        detached_g = [t.detach() for t in inputs]
        with torch.cuda.stream(side), torch.no_grad():
            side.wait_stream(torch.cuda.current_stream())
            layer(*detached_g)
        torch.cuda.current_stream().wait_stream(side)
        fgraph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(fgraph):
            w_f = layer(*detached_g)
        for _ in range(3):
            fgraph.replay()
        barrier()
        g0.record()
        for _ in range(args.steps):
            fgraph.replay()
        g1.record()
        barrier()
        fwd_graph_ms = g0.elapsed_time(g1) / args.steps
        w_e, grads_e = step_device()
        torch.cuda.synchronize()
        assert torch.equal(w_g, w_e) and torch.equal(grads_g[1], grads_e[1]), "graph replay and eager step disagree"
        assert int(layer.last_status.ne(0).sum()) == 0
        graph_note = "CUDA graph replay of one NumericalMarkowitz fwd+bwd step (torch.cuda.graph around the layer call)"
    except Exception as exc:  # noqa: BLE001 - report and keep the eager number
        graph_note = f"eager launches (graph capture failed: {type(exc).__name__}: {exc})"
    finally:
        layer.check_status = {"lazy": "lazy", "sync": True, "off": False}[args.status_mode]

    # forward only (validation / inference path)
    detached = [t.detach() for t in inputs]
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.no_grad():
        for _ in range(3):
            layer(*detached)
        barrier()
        f0.record()
        for _ in range(args.steps):
            layer(*detached)
        f1.record()
    barrier()
    fwd_only_ms = f0.elapsed_time(f1)

    # ---- side measurements (not part of `value`): CUDA-graph replays of the layer's fwd+bwd step on other inputs ---------
    def graph_step_ms(ins, upstream, fwd_only=False, steps=None):
        """ms per step (max over ranks) of `steps` replays of one captured fwd (+ bwd) step of a fresh layer on `ins`."""
        steps = steps or args.steps
        lay = NumericalMarkowitz(ins[0].shape[1], max_weight=u)
        lay.check_status = False

        def one():
            if fwd_only:
                with torch.no_grad():
                    return lay(*[t.detach() for t in ins]), None
            wq = lay(*ins)
            return wq, torch.autograd.grad(wq, ins, upstream)

        sd = torch.cuda.Stream()
        sd.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(sd):
            for _ in range(3):
                one()
        torch.cuda.current_stream().wait_stream(sd)
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            wq, _ = one()
        for _ in range(3):
            gr.replay()
        barrier()
        q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        q0.record()
        for _ in range(steps):
            gr.replay()
        q1.record()
        barrier()
        ms = torch.tensor([q0.elapsed_time(q1) / steps], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        assert int((lay.last_status & 12).ne(0).sum()) == 0
        return float(ms), wq

    # (1) the diversified regime: the same matrices with expected returns scaled down 25x, so that ~86 of the 100 assets end
    #     up free -- what BachelierNet / ThorpNet train in; it runs through the dense kernels
    dense_info = None
    try:
        d_in = [(rets.detach() * 0.04).requires_grad_()] + [t.detach().clone().requires_grad_() for t in (C, gam, alp)]
        d_fb, wd = graph_step_ms(d_in, G)
        d_f, _ = graph_step_ms(d_in, G, fwd_only=True)
        dense_info = {"workload": "same covmat_sqrt, rets~N(0,0.02^2): diversified optimum, CUDA-graph replays",
                      "mean_free": float(((wd > 0) & (wd < u)).sum(1).float().mean()),
                      "fwd_only": {"value": world * B / (d_f * 1e-3), "unit": UNIT, "ms_per_step": d_f},
                      "fwd_bwd": {"value": world * B / (d_fb * 1e-3), "unit": UNIT, "ms_per_step": d_fb}}
        del d_in, wd
    except Exception as exc:                      # never let a side measurement take the bench line down
        dense_info = {"error": f"{type(exc).__name__}: {exc}"}
    barrier()
    # (2) float64: the reference solves in float64 internally; the same kernels instantiated for double
    f64_info = None
    try:
        i64 = [t.detach().double().requires_grad_() for t in (rets, C, gam, alp)]
        ms64, w64 = graph_step_ms(i64, G.double(), steps=max(3, args.steps // 4))
        f64_info = {"value": world * B / (ms64 * 1e-3), "unit": UNIT, "ms_per_step": ms64,
                    "max_abs_diff_vs_f32": float((w64.float() - w_g).abs().max()) if graph_ms is not None else None}
        del i64, w64
    except Exception as exc:
        f64_info = {"error": f"{type(exc).__name__}: {exc}"}
    barrier()
    # (3) strong scaling: ONE global batch of 4096 problems (rank 0's) split over the ranks
    strong_info = None
    try:
        if world > 1:
            gl = [t.detach().clone() for t in (rets, C, gam, alp, G)]
            for t in gl:
                dist.broadcast(t, src=0)
            from deepdow_b200.distributed import shard_bounds
            lo, hi = shard_bounds(B, rank, world)
            # the gamma tiling indexes the whole batch: constant gamma here, so the shard is self-contained
            si = [t[lo:hi].contiguous().requires_grad_() for t in gl[:4]]
            sms, _ = graph_step_ms(si, gl[4][lo:hi].contiguous())
            del gl, si
        else:
            sms = graph_ms / args.steps if graph_ms is not None else None
        strong_info = {"value": B / (sms * 1e-3) if sms else None, "unit": UNIT, "ms_per_step": sms, "global_batch": B,
                       "per_rank": (B + world - 1) // world, "scaling": "strong"}
    except Exception as exc:
        strong_info = {"error": f"{type(exc).__name__}: {exc}"}
    barrier()
    # (4) BachelierNet training (BASELINE config 2), NCCL all-reduce of the parameter gradients inside the timed region
    bach_info = None
    try:
        from deepdow_b200 import nn as dnn
        from deepdow_b200.distributed import GradAllReducer
        torch.manual_seed(0)                                      # identical initial parameters on every rank
        net = dnn.BachelierNet(1, 50, hidden_size=32, max_weight=1, shrinkage_strategy="diagonal", p=0.5).to(dev)
        opt = torch.optim.Adam(net.parameters(), lr=1e-2)
        reduce_grads = GradAllReducer(net.parameters())
        bgen = torch.Generator(device=dev).manual_seed(2000 + rank)
        Xb = torch.randn(256, 1, 40, 50, generator=bgen, device=dev) * 0.01
        yb = torch.randn(256, 1, 10, 50, generator=bgen, device=dev) * 0.01

        def train_step():
            loss = dnn.sharpe_ratio_loss(net(Xb), yb).mean()
            opt.zero_grad(set_to_none=True)
            loss.backward()
            reduce_grads()
            opt.step()
            return loss

        for _ in range(5):
            train_step()
        barrier()
        n_train = max(10, min(50, args.steps))
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0.record()
        for _ in range(n_train):
            loss_b = train_step()
        b1.record()
        barrier()
        bms = torch.tensor([b0.elapsed_time(b1) / n_train], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(bms, op=dist.ReduceOp.MAX)
        net.portfolio_opt_layer.synchronize_status()
        bach_info = {"value": world * 256 / (float(bms) * 1e-3), "unit": "samples/s", "ms_per_step": float(bms), "steps": n_train,
                     "scaling": "weak", "collective": "NCCL all-reduce of 3524 parameter gradients (one flat bucket) per step, inside "
                                                      "the timed region" if world > 1 else "none (1 GPU)",
                     "config": "BachelierNet(1, 50, hidden 32), batch 256 per GPU, lookback 40, horizon 10, SharpeRatio, Adam lr 1e-2",
                     "final_loss": float(loss_b.detach())}
        del net, opt, Xb, yb
    except Exception as exc:
        bach_info = {"error": f"{type(exc).__name__}: {exc}"}
    barrier()

    # ---- timed region 2: end to end from pinned host buffers -----------------------------------
    # (a) the host-buffer C-ABI call (ddw_markowitz_host_step through MarkowitzHostPipeline): chunks flow through
    #     H2D / kernels / D2H streams; every input comes from and every output lands in pinned host memory
    # (b) the same through the nn.Module with explicit .to(device) / .copy_ around it (no overlap)
    from deepdow_b200.host import MarkowitzHostPipeline
    host_in = [t.detach().cpu().pin_memory() for t in (rets, C, gam, alp, G)]
    h2d_bytes = sum(t.numel() * 4 for t in host_in)
    pipe = MarkowitzHostPipeline(N, u, chunk=args.chunk, device=dev)
    out = pipe.step(*host_in)
    d2h_bytes = sum(t.numel() * t.element_size() for t in out)
    for _ in range(2):
        out = pipe.step(*host_in, out=out, sync=False)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = pipe.step(*host_in, out=out, sync=False)
    e1.record()
    barrier()
    pipe.check()
    e2e_ms = e0.elapsed_time(e1)
    # the same with the gradient of covmat_sqrt returned in factored form (ddw_markowitz_host_step_factored: two N-vectors
    # per portfolio instead of the N x N outer products; 6.6 MB come back instead of 167 MB)
    outf = pipe.step(*host_in, factored=True)
    d2h_bytes_f = sum(t.numel() * t.element_size() for t in outf)
    for _ in range(2):
        outf = pipe.step(*host_in, out=outf, sync=False, factored=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        outf = pipe.step(*host_in, out=outf, sync=False, factored=True)
    e1.record()
    barrier()
    pipe.check()
    e2e_fact_ms = e0.elapsed_time(e1)
    assert torch.equal(outf.w, out.w) and torch.equal(outf.d_rets, out.d_rets)
    clocks = sampler.stop()          # sampled over the timed regions (device-resident and end-to-end)
    n_chunks = (B + args.chunk - 1) // args.chunk
    # the pipelined host path must reproduce the device path: to rounding (its chunks of 128 portfolios take the one-CTA-per-portfolio
    # forward, the resident batch of 4096 the tensor-core Gram + warp-per-portfolio kernels; an asset whose weight is within
    # rounding of a bound may sit on either side of it)
    with torch.no_grad():
        w_dev = layer(*[t.detach() for t in inputs]).cpu()
    assert float((out.w - w_dev).abs().max()) < 2e-6, "host pipeline and device path disagree"

    host_out = [torch.empty(s, dtype=torch.float32).pin_memory() for s in ((B, N), (B, N), (B, N, N), (B,), (B,))]

    def step_e2e():
        d = [t.to(dev, non_blocking=True) for t in host_in]
        ins = [t.requires_grad_() for t in d[:4]]
        w = layer(*ins)
        grads = torch.autograd.grad(w, ins, d[4])
        for dst, src in zip(host_out, (w.detach(),) + tuple(grads)):
            dst.copy_(src, non_blocking=True)

    for _ in range(3):
        step_e2e()
    barrier()
    l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0.record()
    for _ in range(args.steps):
        step_e2e()
    l1.record()
    barrier()
    e2e_layer_ms = l0.elapsed_time(l1)

    times = torch.tensor([elapsed_ms, e2e_ms, fwd_only_ms, e2e_layer_ms, graph_ms if graph_ms is not None else -1.0, e2e_fact_ms],
                         device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
        ok = torch.tensor([1.0 if graph_ms is not None else 0.0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if float(ok) == 0.0:
            graph_ms = None
    elapsed_ms, e2e_ms, fwd_only_ms, e2e_layer_ms, graph_max, e2e_fact_ms = [float(x) for x in times.tolist()]
    eager_ms = elapsed_ms
    if graph_ms is not None:
        elapsed_ms = graph_max

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        s = 4
        fwd_bytes = (N * N + 2 * N + 2) * s * B       # read covmat_sqrt, rets, gamma, alpha; write w
        bwd_bytes = (2 * N * N + 3 * N + 2) * s * B
        if graph_ms is not None and fwd_graph_ms is not None:
            # device-only durations (graph replays): forward alone, and the step minus the forward
            fwd_ms, bwd_ms = fwd_graph_ms, graph_max / args.steps - fwd_graph_ms
        dom_fwd = fwd_ms >= bwd_ms
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", "r02", "ncu_traffic.json")
        if not os.path.exists(tpath):
            tpath = os.path.join(ROOT, "profiles", "r01_final", "ncu_traffic.json")
        if isinstance(dense_info, dict) and "fwd_bwd" in dense_info:
            # roofline of the diversified regime: algorithmic bytes of fwd + bwd over the step time (122 016 B / solve)
            for key, nbytes in (("fwd_only", fwd_bytes), ("fwd_bwd", fwd_bytes + bwd_bytes)):
                gbs = nbytes / (dense_info[key]["ms_per_step"] * 1e-3) / 1e9
                dense_info[key]["roofline"] = {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak}
        if os.path.exists(tpath):       # dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture
            tj = json.load(open(tpath))
            # the forward is two kernels since round 2 (tensor-core Gram + warp-per-portfolio solver): their DRAM bytes add up
            names = FWD_KERNELS if dom_fwd else ["markowitz_bwd_sparse_kernel"]
            if all(k in tj for k in names):
                traffic, traffic_src = sum(tj[k]["dram_bytes_read"] + tj[k]["dram_bytes_write"] for k in names), tj["source"]
        ach = (fwd_bytes / (fwd_ms * 1e-3) if dom_fwd else bwd_bytes / (bwd_ms * 1e-3)) / 1e9
        steps_total = args.steps
        value = world * B * steps_total / (elapsed_ms * 1e-3)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": elapsed_ms / steps_total, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": world * B, "parallelism": f"dp{world} (independent problems, no collective)",
                       "l2": "inputs (164 MB covmat_sqrt) + gradients (164 MB) per step exceed the 126 MB L2",
                       "status_check": ("off inside the timed graph replays (the status copy to pinned memory is not capturable); "
                                        f"'{args.status_mode}' in the eager and end-to-end measurements") if graph_ms is not None
                       else args.status_mode,
                       "launch": graph_note, "rounds": rounds, "timed_region_s": timed_region_s,
                       "active_set": {"mean_free": nfree, "mean_at_cap": ncap, "fallback_problems": nfallback,
                                      "dense_kernel_problems": ndense}},
            "e2e": {"value": world * B * steps_total / (e2e_ms * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                    "ms_per_step": e2e_ms / steps_total,
                    "api": f"ddw_markowitz_host_step via MarkowitzHostPipeline.step (pinned host buffers, {n_chunks} chunks of "
                           f"{args.chunk} problems on H2D / kernel / D2H streams)",
                    "layer_api_unpipelined": {"value": world * B * steps_total / (e2e_layer_ms * 1e-3),
                                              "ms_per_step": e2e_layer_ms / steps_total},
                    "factored_gradient": {"value": world * B * steps_total / (e2e_fact_ms * 1e-3), "unit": UNIT,
                                          "ms_per_step": e2e_fact_ms / steps_total, "h2d_bytes_per_step": h2d_bytes,
                                          "d2h_bytes_per_step": d2h_bytes_f,
                                          "api": "ddw_markowitz_host_step_factored: d_covmat_sqrt returned as its two factor "
                                                 "vectors per portfolio (rank 2), everything else identical"},
                    "pinned_buffers": numa_note},
            "eager": {"value": world * B * steps_total / (eager_ms * 1e-3), "unit": UNIT, "ms_per_step": eager_ms / steps_total,
                      "note": "same step launched from Python every time (autograd Function + ctypes)"},
            "gpu_launches": 10 * steps_total * rounds,
            "kernels_per_step": ["markowitz_qprobe_kernel (regime probe)", "markowitz_gram_small_kernel (tcgen05 3xTF32 Gram)",
                                 "markowitz_qwarp_kernel (one warp per portfolio)",
                                 "markowitz_fwd_sparse_kernel (list mode: sets beyond the warp solver's frame)",
                                 "markowitz_fwd_kernel (dense work list)",
                                 "markowitz_ipm_kernel (fallback list, empty here)", "markowitz_bwd_sparse_kernel",
                                 "markowitz_bwd_saved_kernel (factors kept by the dense forward)",
                                 "markowitz_bwd_kernel (what is left)", "markowitz_gamma_kernel"],
            "fwd_only": {"value": world * B * steps_total / (fwd_only_ms * 1e-3), "unit": UNIT},
            "dense_regime": dense_info,
            "f64": f64_info,
            "strong_scaling": strong_info,
            "bachelier_train": bach_info,
            "phase_ms": {"fwd": fwd_ms, "bwd": bwd_ms, "host_enqueue": cpu_enqueue_ms,
                         "source": "CUDA-graph replays (forward alone; step minus forward)" if graph_ms is not None
                         else "CUDA events around the eager layer calls"},
            "roofline": {"bound": "hbm", "kernel": " + ".join(FWD_KERNELS) + " (the whole forward, timed as one graph)" if dom_fwd
                         else "markowitz_bwd_sparse_kernel",
                         "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                         "traffic_source": traffic_src,
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": fwd_bytes if dom_fwd else bwd_bytes,
                         "note": "forward = probe + tcgen05 Gram kernel (hand-over loop of its role pipeline, ~7 k cycles per portfolio; "
                                 "tensor pipe 33% active) + one warp per portfolio (latency of the dependent shuffle / FMA chain); backward = "
                                 "sparse-support kernel + d_gamma pass; HBM is the iteration-free ceiling; when the forward leads, its Q "
                                 "scratch traffic (written and partly re-read) is in `traffic` but not in the algorithmic bytes"},
            "roofline_bwd": {"achieved": bwd_bytes / (bwd_ms * 1e-3) / 1e9, "frac": bwd_bytes / (bwd_ms * 1e-3) / 1e9 / peak,
                             "unit": "GB/s"},
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            v, cores, secs = time_cpu_oracle(B, 0, budget_s=12.0)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"the same {B} problems, fwd+bwd, repeated for {secs:.1f} s, float64 oracle port "
                                              "(OpenMP over problems; cvxpylayers/SCS is not installable)"}
            line["host_cpu_count"] = os.cpu_count()
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
