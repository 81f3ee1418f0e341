"""Time the building blocks in isolation (CUDA events, after warm-up)."""
import math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from basq_b200 import _lib, ops
from basq_b200.kernels import KernelSpec
dev = torch.device("cuda:0")

def timeit(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

g = torch.Generator(device=dev).manual_seed(0)
for (n, S) in [(100, 200), (500, 1000), (1000, 2000)]:
    A = torch.randn(n, S, generator=g, device=dev, dtype=torch.float64)
    A[0] = torch.rand(S, generator=g, device=dev, dtype=torch.float64) + 0.1
    print(f"CAR n={n} S={S}: {timeit(lambda: ops.caratheodory(A)):.3f} ms")
for (m, n, k) in [(999, 2000, 11002), (10000, 999, 10000), (999, 999, 10000), (10000, 999, 999)]:
    a = torch.randn(m, k, generator=g, device=dev, dtype=torch.float64)
    b = torch.randn(k, n, generator=g, device=dev, dtype=torch.float64)
    t = timeit(lambda: ops.dgemm(a, b))
    print(f"DGEMM {m}x{n}x{k}: {t:.3f} ms  {2*m*n*k/t/1e9:.1f} TFLOP/s   (torch: {timeit(lambda: a@b):.3f} ms)")
spec = KernelSpec(_lib.RBF, _lib.PLAIN, torch.tensor([2.5]), 1.0)
for (M, q) in [(1000, 99), (10000, 999)]:
    Z = math.sqrt(2.0) * torch.randn(M, 10, generator=g, device=dev)
    om = torch.randn(M, q, generator=g, device=dev, dtype=torch.float64)
    ctx = _lib.context_for(dev)
    print(f"nystrom M={M} q={q}: {timeit(lambda: ops.nystrom_basis(spec, Z, q, omega=om), reps=3, warm=1):.2f} ms")
N, M, q = 4_000_000, 10000, 999
X = math.sqrt(2.0) * torch.randn(N, 10, generator=g, device=dev)
Z = X[:M].clone()
U = torch.linalg.qr(torch.randn(M, q, generator=g, device=dev, dtype=torch.float64)).Q.T.contiguous()
ctx = _lib.context_for(dev)
ctx.profile(True); ctx.profile_read(True)
t0 = time.perf_counter(); idx, w = ops.recombine(spec, X, Z, U); torch.cuda.synchronize(); t1 = time.perf_counter()
print(f"recombine N={N}: wall {1e3*(t1-t0):.1f} ms", {k: (round(v[0],2), v[1]) for k, v in ctx.profile_read(True).items()})
ctx.profile(False)
print(f"recombine N={N} (unprofiled): {timeit(lambda: ops.recombine(spec, X, Z, U), reps=3, warm=1):.1f} ms")
