"""The fp64-INPUT path (SOBER's global dtype is torch.double, SOBER/_settings.py:4-11): wall clock and phases
of Nystrom basis + recombination at BASELINE configurations 2, 4 and 3 with float64 candidates, and the
evaluation rate of the CUDA-core set-sum kernel (setsum_kernel<double, ...>) they route to.
  python scripts/time_fp64.py [2|4|3 ...]      (default: all three)"""
import math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from basq_b200 import _lib, gp, ops, sampler
from basq_b200.kernels import KernelSpec, spec_from_model
def observations(d, n_obs, seed, log=False):
    g = torch.Generator().manual_seed(seed)
    X = math.sqrt(2.0) * torch.randn(n_obs, d, generator=g, dtype=torch.float64)
    c = 1.5 * torch.randn(3, d, generator=g, dtype=torch.float64)
    y = sum(torch.exp(-0.25 * ((X - ci) ** 2).sum(-1)) for ci in c) / 3.0
    return X, (torch.log(y + 1e-12) if log else y)
dev = torch.device("cuda:0")


def run(name, kern, d, N, M, n, reps=3):
    X = sampler.sample_mvn(torch.zeros(d), 2.0 * torch.eye(d), N, seed=7, device=dev, dtype=torch.float64)
    Z = X[:M].clone()
    ctx = _lib.context_for(dev)
    def step():
        _, U = ops.nystrom_basis(kern, Z, n - 1, want_S=False)
        return ops.recombine(kern, X, Z, U)
    step(); torch.cuda.synchronize()
    ctx.profile(True); ctx.profile_read(True)
    ev0 = ctx.pair_evals
    t0 = time.perf_counter()
    for _ in range(reps): idx, w = step()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / reps * 1e3
    prof = ctx.profile_read(True); ctx.profile(False)
    phases = {k: v[0] / reps for k, v in prof.items() if v[0] > 0}
    evals = (ctx.pair_evals - ev0) / reps
    ss = phases.get("set_sum", float("nan"))
    print(f"{name}: N={N:.0e} M={M} n={n} d={d} fp64 inputs: {ms:.1f} ms/step, {N / ms * 1e3:.3g} points/s; "
          f"phases {', '.join(f'{k} {v:.1f}' for k, v in phases.items())}; "
          f"set-sum {evals:.3g} kernel evaluations in {ss:.1f} ms = {evals / (ss * 1e-3) / 1e9:.3g} G evaluations/s", flush=True)


which = sys.argv[1:] or ["2", "4", "3"]
if "2" in which:
    Xo, yo = observations(2, 102, 3)
    m2 = gp.FixedGP(Xo.to(dev), yo.to(dev), gp.ScaleKernel(gp.RBFKernel(1.0), 1.0), noise=1e-4)
    run("config 2 (Tutorial 01, 2-D, VBQ posterior covariance)", spec_from_model(m2, _lib.PRED_COV), 2, 1_000_000, 10_000, 100)
if "4" in which:
    run("config 4 (Tutorial 02, 20-D Matern-5/2)", KernelSpec(_lib.MATERN25, _lib.PLAIN, torch.tensor([4.0]), 1.0),
        20, 4_000_000, 5_000, 500)
if "3" in which:
    Xo, yo = observations(10, 1002, 5)
    m3 = gp.FixedGP(Xo.to(dev), yo.to(dev), gp.ScaleKernel(gp.RBFKernel(2.5), 1.0), noise=1e-10)
    run("config 3 (N = 1e7, d = 10, n = 1000, M = 1e4, VBQ, n_obs = 1002)", spec_from_model(m3, _lib.PRED_COV),
        10, 10_000_000, 10_000, 1000, reps=2)
