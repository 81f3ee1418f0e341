"""Time the Caratheodory kernel alone (CUDA events). BASQ_CAR2=1 selects the previous kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from basq_b200 import ops
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
for (n, S) in [(100, 200), (500, 1000), (1000, 2000), (1000, 1500)]:
    A = torch.randn(n, S, generator=g, device=dev, dtype=torch.float64)
    A[0] = torch.rand(S, generator=g, device=dev, dtype=torch.float64) + 0.1
    for _ in range(2): om = ops.caratheodory(A)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): om = ops.caratheodory(A)
    e1.record(); torch.cuda.synchronize()
    res = float((A @ om - A.sum(1)).abs().max() / A.sum(1).abs().max())
    print(f"CAR n={n} S={S}: {e0.elapsed_time(e1)/5:.3f} ms (incl. ~0.1 ms host copy of A)  kept {int((om>0).sum())}  residual {res:.1e}  min omega {float(om.min()):.1e}")
