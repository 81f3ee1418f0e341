"""Where does one bench step spend its wall time?  Per-call wall clock + library phase timers +
(BASQ_TRACE=1) the library's own trace points, on the bench workload at a reduced N."""
import math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from basq_b200 import _lib, gp as bgp, ops
from basq_b200.kernels import spec_from_model

dev = torch.device("cuda:0")
N = int(os.environ.get("N", 10_000_000))
M, n, d, n_obs = 10_000, 1000, 10, 1002
q = n - 1
Xo, yo = bench.make_observations(d, n_obs)
model = bgp.FixedGP(Xo.to(dev, torch.float32), yo.to(dev), bgp.ScaleKernel(bgp.RBFKernel(2.5), 1.0), noise=1e-10)
kern = spec_from_model(model, _lib.PRED_COV)
g = torch.Generator(device=dev).manual_seed(1000)
X = math.sqrt(2.0) * torch.randn(N, d, generator=g, device=dev, dtype=torch.float32)
Z = X[:M].clone()
Omega = torch.randn(M, q, generator=torch.Generator(device=dev).manual_seed(7), device=dev, dtype=torch.float64)
ctx = _lib.context_for(dev)


def wall(fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3, r


for it in range(4):
    prof = it >= 2
    ctx.profile(prof); ctx.profile_read(True)
    t_n, (_, U) = wall(lambda: ops.nystrom_basis(kern, Z, q, omega=Omega))
    t_r, (idx, w) = wall(lambda: ops.recombine(kern, X, Z, U))
    line = f"== call {it} profile={prof}: nystrom {t_n:.1f} ms  recombine {t_r:.1f} ms  kept {len(idx)}"
    if prof:
        line += "  " + str({k: (round(v[0], 1), v[1]) for k, v in ctx.profile_read(True).items() if v[1]})
    print(line, flush=True)
ctx.profile(False)

# Caratheodory on the first-round system of this workload vs a random system of the same shape
sess = ops.Session(kern, X[: min(N, 2_000_000)], Z, U, min(N, 2_000_000), 0)
A = torch.zeros(sess.n, sess.S, dtype=torch.float64, device=dev)
sess.partial(sess.count(), 0, 1, A)
Ar = torch.randn(n, 2 * n, generator=torch.Generator(device=dev).manual_seed(3), device=dev, dtype=torch.float64)
Ar[0] = Ar[0].abs() + 0.1
for name, mat in (("round-1 system", A), ("random system", Ar)):
    for it in range(3):
        t, om = wall(lambda: ops.caratheodory(mat))
        print(f"CAR {name}: {t:.2f} ms  kept {int((om > 0).sum())}  rank(float)~{int(torch.linalg.matrix_rank(mat))}" if it == 0 else f"CAR {name}: {t:.2f} ms", flush=True)
