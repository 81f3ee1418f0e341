"""Time the round-1 set-sum kernel alone (bench workload shapes) through the staged session.
Usage: python scripts/time_setsum.py [N] ; BASQ_B200_LIB selects an A/B build of the library."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from basq_b200 import _lib, ops
from basq_b200.kernels import KernelSpec
dev = torch.device("cuda:0")
N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
M, q, d = 11002, 999, 10
g = torch.Generator(device=dev).manual_seed(0)
X = math.sqrt(2.0) * torch.randn(N, d, generator=g, device=dev)
Z = X[:M].clone()
U = torch.randn(q, M, generator=g, device=dev, dtype=torch.float64)
spec = KernelSpec(_lib.RBF, _lib.PLAIN, torch.tensor([2.5]), 1.0)
sess = ops.Session(spec, X, Z, U, N, 0)
ctx = _lib.context_for(dev)
F = int(os.environ.get("CELLS", "1"))
for _ in range(2):
    sess.pass_begin(N, 0, F)
ctx.profile(True); ctx.profile_read(True)
reps = int(os.environ.get("REPS", "3"))
import subprocess, threading
rows, stop = [], threading.Event()
def sample():
    while not stop.is_set():
        o = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-i", "0"], capture_output=True, text=True).stdout.strip()
        rows.append(o); stop.wait(0.05)
th = threading.Thread(target=sample, daemon=True)
if reps > 5: th.start()
for _ in range(reps):
    sess.pass_begin(N, 0, F)
torch.cuda.synchronize()
stop.set()
if reps > 5:
    th.join(); print("clock samples (MHz, W, power_cap):", rows[2:-1][:30])
p = ctx.profile_read(True)
ms = p["set_sum"][0] / reps
pairs = float(M) * N
print(f"{os.path.basename(os.environ.get('BASQ_B200_LIB', 'default')):24s} F={F:2d} N={N} set_sum {ms:8.3f} ms  {pairs / ms / 1e9:7.1f} Gpairs/s  "
      f"{pairs / (ms * 1e-3) / 148 / 1.965e9:6.2f} pairs/clk/SM@1965  projection {p['projection'][0] / reps:.2f} ms")
