"""Phase timings of BASELINE config 5 for the pairwise non-linear kernels (WSABI-M, MMLT) and of the
GP posterior variance over 1e7 candidates.  `--small` runs one sweep-sized problem for ncu."""
import math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from basq_b200 import _lib, gp, ops, sampler
from basq_b200.kernels import spec_from_model
dev = torch.device("cuda:0")
small = "--small" in sys.argv

def observations(d, n_obs, seed, log=False):
    g = torch.Generator().manual_seed(seed)
    X = math.sqrt(2.0) * torch.randn(n_obs, d, generator=g, dtype=torch.float64)
    c = 1.5 * torch.randn(3, d, generator=g, dtype=torch.float64)
    y = sum(torch.exp(-0.25 * ((X - ci) ** 2).sum(-1)) for ci in c) / 3.0
    return X, (torch.log(y + 1e-12) if log else y)

d, N, M, n = 10, (1_000_000 if small else 10_000_000), 10_000, 1000
ctx = _lib.context_for(dev)
X = sampler.sample_mvn(torch.zeros(d), 2.0 * torch.eye(d), N, seed=7, device=dev)
Z = X[:M].clone()
Om = torch.randn(M, n - 1, dtype=torch.float64, device=dev)
for name in (["wsabim"] if small else ["wsabim", "mmlt"]):
    Xo, yo = observations(d, 1002, 5 if name == "wsabim" else 6, log=(name == "mmlt"))
    y = torch.sqrt(2.0 * yo) if name == "wsabim" else yo - yo.max()
    model = gp.FixedGP(Xo.to(dev, torch.float32), y.to(dev), gp.ScaleKernel(gp.RBFKernel(2.5), 1.0), noise=1e-10)
    kern = spec_from_model(model, _lib.WSABI_M if name == "wsabim" else _lib.MMLT_G)
    _, U = ops.nystrom_basis(kern, Z, n - 1, omega=Om, want_S=False)
    reps = 1 if small else 2
    if not small:
        ops.recombine(kern, X, Z, U)
    torch.cuda.synchronize()
    ctx.profile(True); ctx.profile_read(True)
    t0 = time.perf_counter()
    for _ in range(reps):
        idx, w = ops.recombine(kern, X, Z, U)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / reps * 1e3
    prof = ctx.profile_read(True); ctx.profile(False)
    phases = ", ".join(f"{k} {v[0] / reps:.1f}" for k, v in prof.items() if v[0] > 0)
    pairs = M * N
    print(f"{name}: recombination of N={N:.0e} (M={M}, n={n}, n_obs=1002): {ms:.1f} ms  [{phases}]  "
          f"first-sweep algorithmic {2 * 1002 * pairs / 1e12:.0f} TFLOP", flush=True)
if not small:
    Xo, yo = observations(d, 1002, 5)
    model = gp.FixedGP(Xo.to(dev, torch.float32), yo.to(dev), gp.ScaleKernel(gp.RBFKernel(2.5), 1.0), noise=1e-10)
    kern = spec_from_model(model, _lib.PRED_COV)
    for env in ("1", "0"):
        os.environ["BASQ_GPVAR"] = env
        ops.gp_predict(kern, X[:200_000], space=0, want_var=True)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        mean, var = ops.gp_predict(kern, X, space=0, want_var=True)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        print(f"GP mean + variance over 1e7 candidates, BASQ_GPVAR={env}: {(t1 - t0) * 1e3:.1f} ms "
              f"(var min {float(var.min()):.3e} max {float(var.max()):.3e})", flush=True)
        if env == "1": v1 = var.clone()
        else: print(f"   max |fused - fp64 GEMM path| = {float((v1 - var).abs().max()):.3e}")
