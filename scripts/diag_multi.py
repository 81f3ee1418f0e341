"""Where a multi-rank step spends its time (rank 0 prints): device-resident step and the host-buffer step of
bench.py, with synchronising wall-clock marks.  torchrun --nproc-per-node 2 scripts/diag_multi.py"""
import math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dev = torch.device(f"cuda:{lr}")
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
from basq_b200 import _lib, gp as bgp, ops, sharded
from basq_b200.kernels import spec_from_model
import bench
a = bench.parse()
Xo, yo = bench.make_observations(a.d, a.n_obs)
model = bgp.FixedGP(Xo.to(dev, torch.float32), yo.to(dev), bgp.ScaleKernel(bgp.RBFKernel(a.lengthscale), 1.0), noise=a.noise)
kern = spec_from_model(model, _lib.PRED_COV)
N, q = a.N, a.n - 1
X = math.sqrt(2.0) * torch.randn(N, a.d, device=dev, dtype=torch.float32)
Z = X[: a.M].clone()
if world > 1: dist.broadcast(Z, 0)
Om = torch.randn(a.M, q, device=dev, dtype=torch.float64)
Xh = torch.empty(N, a.d, dtype=torch.float32).pin_memory(); Xh.copy_(X)
Zh, Oh = Z.cpu().pin_memory(), Om.cpu().pin_memory()
side = torch.cuda.Stream(dev)
ctx = _lib.context_for(dev)

def mark(tag, t0):
    torch.cuda.synchronize(dev)
    t = time.perf_counter()
    if rank == 0: print(f"   {tag:34s} {1e3 * (t - t0):8.2f} ms", flush=True)
    return t

for rep in range(3):
    if world > 1: dist.barrier()
    torch.cuda.synchronize(dev)
    if rank == 0: print(f"-- rep {rep}: device-resident step")
    t = time.perf_counter()
    _, U = ops.nystrom_basis(kern, Z, q, omega=Om, want_S=False); t = mark("nystrom_basis", t)
    sess = ops.Session(kern, X, Z, U, N * world, rank * N); t = mark("Session()", t)
    idx, w = sharded.recombine_sharded(sess, sess.n, sess.S, device=dev); t = mark("recombine_sharded", t)
    idx, w = sharded.gather_result(idx, w, sess.n); t = mark("gather_result", t)
    sess.close(); t = mark("session close", t)
    if rank == 0: print(f"-- rep {rep}: host-buffer step")
    t = time.perf_counter()
    main = torch.cuda.current_stream(dev)
    Zd = Zh.to(dev, non_blocking=True); Od = Oh.to(dev, non_blocking=True)
    with torch.cuda.stream(side):
        Xd = Xh.to(dev, non_blocking=True)
    ev = side.record_event(); t = mark("copies (synchronised)", t)
    _, U = ops.nystrom_basis(kern, Zd, q, omega=Od, want_S=False); t = mark("nystrom_basis", t)
    main.wait_event(ev)
    sess = ops.Session(kern, Xd, Zd, U, N * world, rank * N); t = mark("Session()", t)
    idx, w = sharded.recombine_sharded(sess, sess.n, sess.S, device=dev); t = mark("recombine_sharded", t)
    sess.close(); t = mark("session close", t)
    k, nprom = ctx.conditioning()
    if rank == 0: print(f"   kappa {k:.3g} promotions {nprom}")
if world > 1: dist.destroy_process_group()
