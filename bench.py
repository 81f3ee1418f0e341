#!/usr/bin/env python
"""Benchmark of the kernel-recombination hot path (BASELINE.json metric: candidate points
recombined per second, N -> n at d = 10).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One step = one full pass of the hot path over one batch of synthetic candidates: Nystrom basis of
the landmark Gram matrix + the Tchernychova-Lyons / Caratheodory loop (refined passes)
down to <= n weighted points.
Workload (config.workload): BASELINE config 3 - 10-D Gaussian-mixture setting, batch n = 1000,
N_rec = 1e7 candidates PER GPU (weak scaling: every rank owns 1e7 candidates, one 16 MB all-reduce
per Caratheodory level), M = 1e4 landmarks, RBF kernel, VBQ posterior-covariance kernel object with n_obs = 1002.
`value` times the device-resident path; `e2e` times the host-buffer C-ABI call including the
host<->device copies; `iteration` times one BASQ iteration with the candidates drawn on the device.  `--impl reference` times the CPU oracle port of the reference's algorithm
on a bounded sample (the reference is pure Python and does not ship to the GPU box).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=3)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--N", type=int, default=10_000_000, help="candidates per GPU")
    p.add_argument("--M", type=int, default=10_000, help="Nystrom landmarks")
    p.add_argument("--n", type=int, default=1000, help="batch size (points returned)")
    p.add_argument("--d", type=int, default=10)
    p.add_argument("--n-obs", type=int, default=1002)
    p.add_argument("--cpu-sample", type=int, default=100_000, help="candidates in the CPU-baseline sample")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-e2e", action="store_true")
    return p.parse_args()


# ------------------------------------------------------------------------------------------ workload
LENGTHSCALE = 2.5
NOISE = 1e-4


def make_observations(d, n_obs, seed=11):
    """Synthetic GP observations: prior N(0, 2 I_d) inputs, 3-component Gaussian-mixture likelihood
    (the shape of BASQ/experiment/gmm.py), fixed hyper-parameters."""
    g = torch.Generator().manual_seed(seed)
    X = math.sqrt(2.0) * torch.randn(n_obs, d, generator=g, dtype=torch.float64)
    centres = 1.5 * torch.randn(3, d, generator=g, dtype=torch.float64)
    y = sum(torch.exp(-0.25 * ((X - c) ** 2).sum(-1)) for c in centres) / 3.0
    return X, y


def workload_name(a):
    return (f"BASELINE config 3: N_rec={a.N:g}/GPU d={a.d} n={a.n} M={a.M} RBF l={LENGTHSCALE} "
            f"VBQ predictive-covariance kernel n_obs={a.n_obs}")


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.stop = index, [], threading.Event()
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([c.strip() for c in out.strip().split(",")])
            except Exception:
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *exc):
        self.stop.set()
        self.t.join(timeout=6)

    def summary(self):
        sm = sorted(float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 7:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------ reference arm
def oracle_step(a, n_sample, seed=0):
    """One pass of the reference's algorithm (oracle port, torch CPU, all host threads) over a
    bounded sample of the workload: same d, M, n, kernel object; N reduced to n_sample."""
    from oracle import gp_kernels as ogp
    from oracle import rchq as orchq

    Xo, yo = make_observations(a.d, a.n_obs)
    model = ogp.ExactGP(Xo, yo, ogp.ScaleKernel(ogp.RBFKernel(LENGTHSCALE), 1.0), noise=NOISE)
    kern = ogp.VanillaGP(model).predictive_kernel
    g = torch.Generator().manual_seed(seed)
    X = (math.sqrt(2.0) * torch.randn(n_sample, a.d, generator=g)).double()
    Z = X[: a.M].clone()
    t0 = time.perf_counter()
    torch.manual_seed(seed)
    idx, w = orchq.recombination(X, Z, a.n, kern, chunk=max(1, 200_000 // (2 * a.n)))
    dt = time.perf_counter() - t0
    assert len(idx) <= a.n and bool((w > 0).all())
    return dt


def run_reference(a, rank):
    if rank != 0:
        return
    cores = torch.get_num_threads()
    # bounded sample: about 40 s of CPU work per step at the default 1e5 candidates; shrink it when more
    # than three timed steps are requested so that the whole run stays within a few minutes
    n_sample = max(int(a.cpu_sample * min(1.0, 3.0 / max(a.steps, 1))), 4 * a.n, a.M)
    for _ in range(min(a.warmup, 1)):
        oracle_step(a, max(4 * a.n, a.M, n_sample // 8))
    times = [oracle_step(a, n_sample, seed=s) for s in range(max(1, a.steps))]
    t = sum(times) / len(times)
    v = n_sample / t
    sample = f"N={n_sample} candidates of the same workload (d={a.d}, M={a.M}, n={a.n}, n_obs={a.n_obs}), fp64 torch CPU"
    line = {
        "impl": "reference", "metric": "candidate points recombined/sec (N->n, d=10)", "value": v,
        "unit": "points/s", "n_gpus": a.gpus, "steps": len(times), "warmup": min(a.warmup, 1), "ms_per_step": t * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(a), "sample": sample},
        "cpu_baseline": {"value": v, "unit": "points/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ our arm
def run_ours(a, rank, world, local_rank):
    import torch.distributed as dist

    import basq_b200  # noqa: F401
    from basq_b200 import _lib, gp as bgp, ops, sharded

    assert torch.cuda.is_available(), "bench.py --impl ours needs a CUDA device (no CPU fallback)"
    dev = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(dev)
    ctx = _lib.context_for(dev)
    q = a.n - 1

    # GP model (replicated): fixed hyper-parameters, caches built once outside the timed region
    Xo, yo = make_observations(a.d, a.n_obs)
    model = bgp.FixedGP(Xo.to(dev, torch.float32), yo.to(dev), bgp.ScaleKernel(bgp.RBFKernel(LENGTHSCALE), 1.0),
                        noise=NOISE)
    from basq_b200.kernels import spec_from_model
    kern = spec_from_model(model, _lib.PRED_COV)

    # synthetic candidates: rank-local shard, pinned host copy for the end-to-end leg
    N_loc, N_glob = a.N, a.N * world
    g = torch.Generator(device=dev).manual_seed(1000 + rank)
    X = math.sqrt(2.0) * torch.randn(N_loc, a.d, generator=g, device=dev, dtype=torch.float32)
    if world > 1:
        Z = X[: a.M].clone()
        dist.broadcast(Z, 0)
    else:
        Z = X[: a.M].clone()
    gO = torch.Generator(device=dev).manual_seed(7)
    Omega = torch.randn(a.M, q, generator=gO, device=dev, dtype=torch.float64)

    def step_device():
        _, U = ops.nystrom_basis(kern, Z, q, omega=Omega, want_S=False)
        if world == 1:
            return ops.recombine(kern, X, Z, U)
        return sharded.recombination_sharded(X, Z, a.n, kern, N_glob, rank * N_loc, U)

    X_host = Z_host = Om_host = None
    side_stream = torch.cuda.Stream(dev)
    if not a.no_e2e:
        X_host = torch.empty(N_loc, a.d, dtype=torch.float32).pin_memory()
        X_host.copy_(X)
        Z_host = Z.cpu().pin_memory()
        Om_host = Omega.cpu().pin_memory()

    def step_e2e():
        if world == 1:
            return ops.recombine_host(kern, X_host, Z_host, q, omega_host=Om_host, device=dev)
        # landmarks and test matrix first (the copy engine is FIFO), candidates on a side stream underneath
        # the Nystrom phase - what basq_recombine_host does inside the C call at N = 1
        main = torch.cuda.current_stream(dev)
        Zd = Z_host.to(dev, non_blocking=True)
        Od = Om_host.to(dev, non_blocking=True)
        with torch.cuda.stream(side_stream):
            Xd = X_host.to(dev, non_blocking=True)
        x_ready = side_stream.record_event()
        _, U = ops.nystrom_basis(kern, Zd, q, omega=Od, want_S=False)
        main.wait_event(x_ready)
        Xd.record_stream(main)
        idx, w = sharded.recombination_sharded(Xd, Zd, a.n, kern, N_glob, rank * N_loc, U)
        return idx.cpu(), w.cpu()

    # BASELINE metric (2), "BASQ iteration time": candidates drawn on the device from the prior (every
    # rank its slice of one Philox stream), landmarks = leading rows, Nystrom basis, recombination
    # (GP hyper-parameter refit excluded, as in the reference's tutorial timing - SURVEY 8d)
    from basq_b200 import sampler as bsampler
    prior_mean = torch.zeros(a.d, dtype=torch.float64)
    prior_tril = math.sqrt(2.0) * torch.eye(a.d, dtype=torch.float64)
    it_count = [0]

    def step_iteration():
        it_count[0] += 1
        seed = 4242 + it_count[0]
        Xs = bsampler.sample_mvn(prior_mean, None, N_loc, seed=seed, offset=rank * N_loc, device=dev,
                                 scale_tril=prior_tril)
        Zs = Xs[: a.M] if world == 1 else bsampler.sample_mvn(prior_mean, None, a.M, seed=seed, offset=0, device=dev,
                                                              scale_tril=prior_tril)
        _, U = ops.nystrom_basis(kern, Zs, q, omega=Omega, want_S=False)
        if world == 1:
            return ops.recombine(kern, Xs, Zs, U)
        return sharded.recombination_sharded(Xs, Zs, a.n, kern, N_glob, rank * N_loc, U)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = None
        for _ in range(steps):
            out = fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t[0])
        return ms / steps, out

    for _ in range(max(a.warmup, 3)):
        out = step_device()
    # phase spans are recorded as CUDA-event pairs on the library's stream (= torch's current stream)
    # without synchronising, so they are taken live inside the timed region and read afterwards
    ctx.profile(True)
    ctx.profile_read(reset=True)
    launches0, pe0 = ctx.launches, ctx.pair_evals
    with ClockSampler(local_rank) as clk:
        ms_step, out = timed(step_device, a.steps)
    launches = ctx.launches - launches0
    pairs_step = (ctx.pair_evals - pe0) / a.steps
    prof = ctx.profile_read(reset=True)
    ctx.profile(False)
    idx, w = out
    assert 1 <= len(idx) <= a.n and bool((w > 0).all()), "invalid quadrature rule"
    assert abs(float(w.sum()) - 1.0) < 1e-9, float(w.sum())
    value = N_glob / (ms_step * 1e-3)

    e2e = None
    if not a.no_e2e:
        step_e2e()
        ms_e2e, out2 = timed(step_e2e, a.steps)
        h2d = X_host.numel() * 4 + Z_host.numel() * 4 + Om_host.numel() * 8
        d2h = len(out2[0]) * 16
        e2e = {"value": N_glob / (ms_e2e * 1e-3), "unit": "points/s", "ms_per_step": ms_e2e,
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)}

    step_iteration()
    ms_iter, out3 = timed(step_iteration, a.steps)
    assert 1 <= len(out3[0]) <= a.n and abs(float(out3[1].sum()) - 1.0) < 1e-9

    line = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        bf16_peak = peaks.get("bf16_tflops_sustained", 1400.0)
        peak_src = ("measured (MEASURED_PEAKS.json bf16_tflops_sustained: kernel timed inside a long step)"
                    if peaks else "fallback (B200_PROFILING.md: 1.4 PFLOP/s sustained)")
        ss_ms_tot, ss_calls = prof["set_sum"]
        ss_ms = ss_ms_tot / a.steps                    # set-sum kernel time per step (one launch per round)
        n_rounds = ss_calls // a.steps
        ss_launch_ms = ss_ms_tot / max(ss_calls, 1)    # average launch duration
        pairs_launch = pairs_step / max(n_rounds, 1)   # kernel evaluations k(z, x) per launch (average)
        flop_per_pair = 2 * a.d + 4                    # algorithmic: d-term distance contraction (2d), bias adds, exp, weighted add
        ach = pairs_launch * flop_per_pair / (ss_launch_ms * 1e-3) / 1e12 if ss_ms > 0 else None
        clocks = clk.summary()
        sm_mhz = clocks.get("sm_mhz") or 1965.0
        mufu_peak = 148 * 16 * sm_mhz * 1e6            # ex2.approx lanes/s: 16 per clock per SM
        pair_rate = pairs_launch / (ss_launch_ms * 1e-3) if ss_ms > 0 else None
        traffic = None
        try:   # per-launch DRAM bytes of this kernel from the committed ncu --set full capture
            traffic = json.load(open(os.path.join(ROOT, "profiles", "setsum_traffic.json")))["dram_bytes_per_launch_avg"]
        except Exception:
            pass
        line = {
            "metric": "candidate points recombined/sec (N->n, d=10)", "value": value, "unit": "points/s",
            "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 kernel evaluation (3xTF32 distance contraction), f64 accumulation/projection/Caratheodory",
            "data": "synthetic",
            "config": {"workload": workload_name(a), "N_total": N_glob, "parallelism": f"dp{world}",
                       "l2_policy": "inputs (640 MB of candidate records per GPU) exceed the 126 MB L2"},
            "clocks": clocks, "gpu_launches": int(launches), "e2e": e2e,
            "iteration": {"ms": ms_iter, "unit": "ms per BASQ iteration",
                          "what": "device prior sampling (Philox MVN) + Nystrom basis + recombination to n points; "
                                  "GP refit excluded (BASELINE metric part 2)"},
            "phases_ms": {k: round(v[0] / a.steps, 3) for k, v in prof.items()}, "rounds": n_rounds,
            "roofline": {
                "bound": "tensor", "kernel": "setsum_mma_kernel<RBF, 10> (tcgen05 kind::tf32 distance contraction + ex2 epilogue)",
                "achieved": ach, "peak": bf16_peak, "unit": "TFLOP/s", "frac": (ach / bf16_peak) if ach else None,
                "traffic": traffic, "peak_source": peak_src,
                "flop_per_pair": flop_per_pair, "pairs_per_launch_avg": pairs_launch, "launch_ms_avg": ss_launch_ms,
                "note": ("algorithmic flop = (2d+4) per kernel evaluation k(z,x); the kernel issues 80 tf32 MMA "
                         "flop per evaluation (K = 3d+6 padded to 40, 3xTF32) and is bound by the MUFU pipe "
                         "(one ex2 per evaluation), not by the tensor pipe - see sfu_roofline and DESIGN.md 4"),
            },
            "sfu_roofline": {
                "bound": "mufu", "achieved": pair_rate, "peak": mufu_peak,
                "unit": "kernel evaluations/s (one ex2.approx each; peak = 148 SMs x 16 lanes/clk x measured SM clock)",
                "frac": (pair_rate / mufu_peak) if pair_rate else None,
                "pairs_per_step": pairs_step, "set_sum_ms_per_step": ss_ms, "set_sum_launches_per_step": n_rounds,
            },
        }
    if rank == 0 and not a.no_cpu_baseline and world == 1:
        cores = torch.get_num_threads()
        n_sample = max(a.cpu_sample, 4 * a.n, a.M)
        t = oracle_step(a, n_sample)
        line["cpu_baseline"] = {
            "value": n_sample / t, "unit": "points/s", "cores": cores, "kind": "port", "seconds": t,
            "sample": f"N={n_sample} candidates of the same workload (d={a.d}, M={a.M}, n={a.n}, n_obs={a.n_obs}), "
                      "oracle port of BASQ/_rchq.py, fp64 torch CPU"}
    if rank == 0:
        print(json.dumps(line), flush=True)


def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if a.impl == "reference":
        run_reference(a, rank)
        return
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    try:
        run_ours(a, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
