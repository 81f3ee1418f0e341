"""fp64 DMMA GEMM rates on the shapes of the hot path (projection, CholeskyQR Gram, apply)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from basq_b200 import ops
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
shapes = [("projection U' Gf", 999, 2000, 11002, False, False), ("projection (n cols)", 999, 1000, 11002, False, False),
          ("Gram Y^T Y", 999, 999, 10000, True, False), ("apply Y L^-T", 10000, 999, 999, False, True),
          ("K_ZX W", 10000, 1002, 1002, False, False)]
for name, m, n, k, ta, tb in shapes:
    A = torch.randn((k, m) if ta else (m, k), generator=g, device=dev, dtype=torch.float64)
    B = torch.randn((n, k) if tb else (k, n), generator=g, device=dev, dtype=torch.float64)
    C = ops.dgemm(A, B, ta, tb); torch.cuda.synchronize()
    ref = (A.T if ta else A) @ (B.T if tb else B)
    err = float((C - ref).abs().max() / ref.abs().max())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): ops.dgemm(A, B, ta, tb)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    e0.record()
    for _ in range(10): (A.T if ta else A) @ (B.T if tb else B)
    e1.record(); torch.cuda.synchronize()
    ms_t = e0.elapsed_time(e1) / 10
    print(f"{name:22s} {m}x{n}x{k}: {ms:7.3f} ms = {2e-9 * m * n * k / ms:6.2f} TFLOP/s   (torch/cuBLAS {ms_t:7.3f} ms = {2e-9 * m * n * k / ms_t:6.2f})  err {err:.1e}")
