"""The level-0 projection GEMM alone (999 x 11002 x 2000, fp64) for ncu captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from basq_b200 import ops
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
a = torch.randn(999, 11002, generator=g, device=dev, dtype=torch.float64)
b = torch.randn(11002, 2000, generator=g, device=dev, dtype=torch.float64)
for _ in range(5): c = ops.dgemm(a, b)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): c = ops.dgemm(a, b)
e1.record(); torch.cuda.synchronize()
t = e0.elapsed_time(e1) / 10
print(f"dgemm 999 x 11002 x 2000: {t:.3f} ms = {2 * 999 * 11002 * 2000 / t / 1e9:.1f} TFLOP/s")
