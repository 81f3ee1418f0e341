"""Where does the end-to-end (host-buffer) call spend its extra time?  Times, for the bench workload,
(a) the device-resident path, (b) basq_recombine_host, (c) a bare pinned H2D copy of the candidates."""
import math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from basq_b200 import _lib, gp as bgp, ops
from basq_b200.kernels import spec_from_model
dev = torch.device("cuda:0")
N = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
d, M, q, n_obs = 10, 10000, 999, 1002
Xo, yo = bench.make_observations(d, n_obs)
model = bgp.FixedGP(Xo.to(dev, torch.float32), yo.to(dev), bgp.ScaleKernel(bgp.RBFKernel(2.5), 1.0), noise=1e-10)
kern = spec_from_model(model, _lib.PRED_COV)
g = torch.Generator(device=dev).manual_seed(1)
X = math.sqrt(2.0) * torch.randn(N, d, generator=g, device=dev)
Z = X[:M].clone()
Om = torch.randn(M, q, generator=g, device=dev, dtype=torch.float64)
Xh = torch.empty(N, d).pin_memory(); Xh.copy_(X)
Zh = Z.cpu().pin_memory(); Oh = Om.cpu().pin_memory()
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    each = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        each.append((time.perf_counter() - t0) * 1e3)
    if os.environ.get("PROBE_EACH"): print("   per call:", " ".join(f"{x:.1f}" for x in each), file=sys.stderr)
    return min(each)
def dev_path():
    _, U = ops.nystrom_basis(kern, Z, q, omega=Om, want_S=False)
    return ops.recombine(kern, X, Z, U)
print(f"N={N}: device path {t(dev_path):.2f} ms | host call {t(lambda: ops.recombine_host(kern, Xh, Zh, q, device=dev, seed=3)):.2f} ms | "
      f"bare H2D X {t(lambda: Xh.to(dev, non_blocking=True)):.2f} ms, Omega {t(lambda: Oh.to(dev, non_blocking=True)):.2f} ms")
s2 = torch.cuda.Stream()
def overlapped():
    with torch.cuda.stream(s2):
        Xd = Xh.to(dev, non_blocking=True)
    ev = s2.record_event()
    _, U = ops.nystrom_basis(kern, Z, q, omega=Om, want_S=False)
    torch.cuda.current_stream().wait_event(ev)
    return ops.recombine(kern, Xd, Z, U)
def nys_only():
    return ops.nystrom_basis(kern, Z, q, omega=Om, want_S=False)
def nys_plus_copy():
    with torch.cuda.stream(s2):
        Xd = Xh.to(dev, non_blocking=True)
    r = ops.nystrom_basis(kern, Z, q, omega=Om, want_S=False)
    s2.synchronize()
    return r
print(f"torch side-stream copy + device path {t(overlapped):.2f} ms | nystrom alone {t(nys_only):.2f} ms | nystrom with a concurrent X copy {t(nys_plus_copy):.2f} ms")
