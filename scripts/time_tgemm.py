"""tgemm_kernel alone on the Nystrom shape (K Y: 999 x 1e4 x 1e4) and the whole Nystrom basis; A/B builds through
BASQ_B200_LIB=<other .so>."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from basq_b200 import _lib, ops
from basq_b200.kernels import KernelSpec
dev = torch.device("cuda:0")


def timeit(fn, reps=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


g = torch.Generator(device=dev).manual_seed(0)
for (m, n, k) in [(999, 10000, 10000), (10000, 999, 10000), (1002, 4096, 1002)]:
    A = torch.randn(m, k, generator=g, device=dev, dtype=torch.float64)
    B = torch.randn(n, k, generator=g, device=dev, dtype=torch.float64)
    t = timeit(lambda: ops.tgemm(A, B))
    print(f"tgemm (conversion + product) {m} x {n} x {k}: {t:.3f} ms = {2 * m * n * k / t / 1e9:.0f} TFLOP/s (fp32-accurate product)")
spec = KernelSpec(_lib.RBF, _lib.PLAIN, torch.tensor([2.5]), 1.0)
Z = math.sqrt(2.0) * torch.randn(10000, 10, generator=g, device=dev)
om = torch.randn(10000, 999, generator=g, device=dev, dtype=torch.float64)
print(f"{os.environ.get('BASQ_B200_LIB', 'default build')}: nystrom M=10000 q=999: "
      f"{timeit(lambda: ops.nystrom_basis(spec, Z, 999, omega=om, want_S=False), reps=5, warm=2):.2f} ms")
