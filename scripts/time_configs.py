"""Wall-clock of the hot path (Nystrom basis + recombination, device-resident inputs) at the full sizes
of BASELINE.json's configurations 2-5 (config 3 is bench.py's workload)."""
import math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from basq_b200 import _lib, gp, ops, sampler
from basq_b200.kernels import KernelSpec, spec_from_model
dev = torch.device("cuda:0")

def observations(d, n_obs, seed, log=False):
    g = torch.Generator().manual_seed(seed)
    X = math.sqrt(2.0) * torch.randn(n_obs, d, generator=g, dtype=torch.float64)
    c = 1.5 * torch.randn(3, d, generator=g, dtype=torch.float64)
    y = sum(torch.exp(-0.25 * ((X - ci) ** 2).sum(-1)) for ci in c) / 3.0
    return X, (torch.log(y + 1e-12) if log else y)

def run(name, kern, d, N, M, n, reps=3):
    X = sampler.sample_mvn(torch.zeros(d), 2.0 * torch.eye(d), N, seed=7, device=dev)
    Z = X[:M].clone()
    Om = torch.randn(M, n - 1, dtype=torch.float64, device=dev)
    def step():
        _, U = ops.nystrom_basis(kern, Z, n - 1, omega=Om, want_S=False)
        return ops.recombine(kern, X, Z, U)
    step(); torch.cuda.synchronize()
    ctx = _lib.context_for(dev)
    ctx.profile(True); ctx.profile_read(True)
    t0 = time.perf_counter()
    for _ in range(reps): idx, w = step()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / reps * 1e3
    prof = ctx.profile_read(True); ctx.profile(False)
    phases = ", ".join(f"{k} {v[0] / reps:.1f}" for k, v in prof.items() if v[0] > 0)
    print(f"{name:58s} N={N:.0e} M={M} n={n} d={d}: {ms:9.1f} ms  {N / ms * 1e3:.3g} points/s  ({len(idx)} points)  [{phases}]")

Xo, yo = observations(10, 102, 2)
m1 = gp.FixedGP(Xo.to(dev, torch.float32), yo.to(dev), gp.ScaleKernel(gp.RBFKernel(2.5), 1.0), noise=1e-4)
run("config 1: main.py default, 10-D, VBQ (batch selection)", spec_from_model(m1, _lib.PRED_COV), 10, 20_000, 200, 100, reps=10)
run("config 1: main.py default, 10-D, VBQ (quadrature)", spec_from_model(m1, _lib.PRED_COV), 10, 100_000, 200, 100, reps=10)
Xo, yo = observations(2, 102, 3)
m2 = gp.FixedGP(Xo.to(dev, torch.float32), yo.to(dev), gp.ScaleKernel(gp.RBFKernel(1.0), 1.0), noise=1e-4)
run("config 2: Tutorial 01, 2-D, VBQ posterior covariance", spec_from_model(m2, _lib.PRED_COV), 2, 1_000_000, 10_000, 100)
run("config 4: Tutorial 02, 20-D Matern-5/2", KernelSpec(_lib.MATERN25, _lib.PLAIN, torch.tensor([4.0]), 1.0), 20, 4_000_000, 5_000, 500)
Xo, yo = observations(10, 1002, 5)
m5 = gp.FixedGP(Xo.to(dev, torch.float32), torch.sqrt(2.0 * yo).to(dev), gp.ScaleKernel(gp.RBFKernel(2.5), 1.0), noise=1e-4)
run("config 5: Tutorial 03, WSABI-L, n_obs = 1002", spec_from_model(m5, _lib.WSABI_L), 10, 10_000_000, 10_000, 1000)
kern = spec_from_model(m5, _lib.PRED_COV)
X = sampler.sample_mvn(torch.zeros(10), 2.0 * torch.eye(10), 10_000_000, seed=9, device=dev)
torch.cuda.synchronize(); t0 = time.perf_counter()
mean, var = ops.gp_predict(kern, X, space=0, want_var=True)
torch.cuda.synchronize(); t1 = time.perf_counter()
w = sampler.calc_weights(kern, X, ratio=0.5)
torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"config 5: GP mean + variance over 1e7 candidates (n_obs = 1002): {1e3 * (t1 - t0):.1f} ms; calc_weights on top: {1e3 * (t2 - t1):.1f} ms")
