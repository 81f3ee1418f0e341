"""Time recombination with the MMLT (pairwise non-linear) kernel: N candidates, M = 1e4, n = 1000, n_obs = 1002."""
import math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from basq_b200 import _lib, gp, ops, sampler
from basq_b200.kernels import spec_from_model
dev = torch.device("cuda:0")
N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 2_000_000
d, M, n, n_obs = 10, 10_000, 1000, 1002
g = torch.Generator().manual_seed(6)
Xo = math.sqrt(2.0) * torch.randn(n_obs, d, generator=g, dtype=torch.float64)
c = 1.5 * torch.randn(3, d, generator=g, dtype=torch.float64)
y = torch.log(sum(torch.exp(-0.25 * ((Xo - ci) ** 2).sum(-1)) for ci in c) / 3.0 + 1e-12)
model = gp.FixedGP(Xo.to(dev, torch.float32), (y - y.max()).to(dev), gp.ScaleKernel(gp.RBFKernel(2.5), 1.0), noise=1e-4)
kern = spec_from_model(model, _lib.MMLT_G)
X = sampler.sample_mvn(torch.zeros(d), 2.0 * torch.eye(d), N, seed=24, device=dev)
Z = X[:M].clone()
torch.manual_seed(0)
_, U = ops.nystrom_basis(kern, Z, n - 1, want_S=False)
torch.cuda.synchronize(); t0 = time.perf_counter()
idx, w = ops.recombine(kern, X, Z, U)
torch.cuda.synchronize(); t1 = time.perf_counter()
print(f"MMLT N={N}: recombine {t1 - t0:.3f} s, {len(idx)} points, sum w = {float(w.sum()):.12f}  BASQ_CELL_FACTOR={os.environ.get('BASQ_CELL_FACTOR', 'auto')}")
