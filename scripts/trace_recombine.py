import math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from basq_b200 import _lib, ops
from basq_b200.kernels import KernelSpec
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
spec = KernelSpec(_lib.RBF, _lib.PLAIN, torch.tensor([2.5]), 1.0)
N, M, q = int(os.environ.get("N", 4_000_000)), 10000, 999
X = math.sqrt(2.0) * torch.randn(N, 10, generator=g, device=dev)
Z = X[:M].clone()
U = torch.linalg.qr(torch.randn(M, q, generator=g, device=dev, dtype=torch.float64)).Q.T.contiguous()
for it in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    idx, w = ops.recombine(spec, X, Z, U)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"== call {it}: wall {1e3*(t1-t0):.1f} ms", flush=True)
