import time, torch, ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
dev = torch.device("cuda:0")
N = 10_000_000
x = torch.empty(N, 10, dtype=torch.float32).pin_memory()
x.normal_()
print("pinned:", x.is_pinned())
torch.cuda.synchronize()
for _ in range(3):
    t0 = time.perf_counter(); y = x.to(dev, non_blocking=True); torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"torch pinned H2D 400MB: {(t1-t0)*1e3:.1f} ms  {0.4/(t1-t0):.1f} GB/s")
xp = torch.empty(N, 10, dtype=torch.float32).normal_()
t0 = time.perf_counter(); y = xp.to(dev); torch.cuda.synchronize(); t1 = time.perf_counter()
print(f"torch pageable H2D 400MB: {(t1-t0)*1e3:.1f} ms")
for sz in (16, 256, 1024, 2048):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    b = torch.empty(sz << 20, dtype=torch.uint8, device=dev); torch.cuda.synchronize(); t1 = time.perf_counter()
    del b; torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"torch alloc {sz} MB: {(t1-t0)*1e3:.2f} ms")
cudart = ctypes.CDLL("libcudart.so.12") if False else None
from basq_b200 import _lib, ops, gp as bgp
from basq_b200.kernels import KernelSpec
spec = KernelSpec(_lib.RBF, _lib.PLAIN, torch.tensor([2.5]), 1.0)
Z = x[:1000].clone().pin_memory()
om = torch.randn(1000, 99, dtype=torch.float64).pin_memory()
ctx = _lib.context_for(dev)
for it in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    idx, w = ops.recombine_host(spec, x, Z, 99, omega_host=om, device=dev)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    xd = x.to(dev, non_blocking=True); zd = Z.to(dev); od = om.to(dev)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    S, U = ops.nystrom_basis(spec, zd, 99, omega=od); i2, w2 = ops.recombine(spec, xd, zd, U)
    torch.cuda.synchronize(); t3 = time.perf_counter()
    print(f"recombine_host {1e3*(t1-t0):.1f} ms | torch H2D {1e3*(t2-t1):.1f} ms + device path {1e3*(t3-t2):.1f} ms")
ctx.profile(True); ctx.profile_read(True)
t0 = time.perf_counter(); ops.recombine_host(spec, x, Z, 99, omega_host=om, device=dev); t1 = time.perf_counter()
print("profiled host call", 1e3*(t1-t0), ctx.profile_read(True))
