"""Small-shape pass over every kernel family of libbasq_b200.so, for compute-sanitizer
(scripts/sanitize.sh).  Shapes are chosen so that each tool finishes in minutes: the tcgen05 set-sum
kernel, the cooperative Caratheodory kernels, the Nystrom range finder (tcgen05 GEMM, cooperative
Cholesky), the fp64 DMMA GEMM, GP prediction, the non-linear modes and the candidate-side kernels."""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import basq_b200
from basq_b200 import _lib, ops, sampler
from oracle import gp_kernels as ogp

dev = torch.device("cuda:0")
which = set(sys.argv[1:]) or {"linear", "nonlinear", "fp64", "nystrom", "gp", "candidates", "gemm", "host"}
g = torch.Generator().manual_seed(0)
d, N, M, n = 6, 6000, 256, 24
X = (math.sqrt(2.0) * torch.randn(N, d, generator=g)).float().to(dev)
Z = X[:M].clone()
model = ogp.make_gp(d, 40, lengthscale=2.0, noise=1e-3, seed=1)
modell = ogp.make_gp(d, 40, lengthscale=2.0, noise=1e-3, seed=1, log_targets=True)
U = torch.linalg.qr(torch.randn(M, n - 1, generator=g, dtype=torch.float64)).Q.T.contiguous().to(dev)


def check(tag, idx, w):
    torch.cuda.synchronize()
    assert 1 <= len(idx) <= n and bool((w > 0).all()) and abs(float(w.sum()) - 1.0) < 1e-9, tag
    print(f"[sanitize] {tag}: ok ({len(idx)} points)", flush=True)


if "linear" in which:
    check("plain rbf fp32", *ops.recombine(model.covar_module.forward, X, Z, U))
    check("pred_cov fp32", *ops.recombine(ogp.VanillaGP(model).predictive_kernel, X, Z, U))
    check("wsabi-l fp32", *ops.recombine(ogp.WsabiGP(model).wsabil_kernel, X, Z, U))
    mu = torch.rand(N, generator=g, dtype=torch.float64); mu[::3] = 0; mu /= mu.sum()
    check("weighted fp32", *ops.recombine(model.covar_module.forward, X, Z, U, mu=mu.to(dev)))
if "nonlinear" in which:
    check("wsabi-m fp32", *ops.recombine(ogp.WsabiGP(model).wsabim_kernel, X, Z, U))
    check("mmlt fp32", *ops.recombine(ogp.ScaleMmltGP(modell).gspace_kernel, X, Z, U))
if "fp64" in which:
    check("pred_cov fp64", *ops.recombine(ogp.VanillaGP(model).predictive_kernel, X.double(), Z.double(), U))
if "nystrom" in which:
    S, Ub = ops.nystrom_basis(ogp.VanillaGP(model).predictive_kernel, Z, n - 1)
    torch.cuda.synchronize()
    assert float((Ub @ Ub.T - torch.eye(n - 1, dtype=torch.float64, device=dev)).abs().max()) < 1e-9
    print("[sanitize] nystrom: ok", flush=True)
if "gp" in which:
    mean, var = ops.gp_predict(ogp.VanillaGP(model).predictive_kernel, X)
    Phi = ops.features(ogp.VanillaGP(model).predictive_kernel, X[:500], Z, U)
    K = ops.gram(ogp.WsabiGP(model).wsabim_kernel, Z, X[:300])
    torch.cuda.synchronize()
    assert bool(torch.isfinite(mean).all()) and bool((var > 0).all()) and bool(torch.isfinite(Phi).all())
    print("[sanitize] gp predict / features / gram: ok", flush=True)
    # mean on the tensor-core set-sum kernel with the roles swapped (n_obs >= 64)
    model80 = ogp.make_gp(d, 80, lengthscale=2.0, noise=1e-3, seed=2)
    m80, _ = ops.gp_predict(ogp.VanillaGP(model80).predictive_kernel, X[:3000], want_var=False)
    check("wsabi-l fp32, n_obs = 80", *ops.recombine(ogp.WsabiGP(model80).wsabil_kernel, X, Z, U))
    torch.cuda.synchronize()
    assert bool(torch.isfinite(m80).all())
if "candidates" in which:
    Xs = sampler.sample_mvn(torch.zeros(d), 2.0 * torch.eye(d), 5000, seed=3, device=dev)
    w = sampler.calc_weights(ogp.VanillaGP(model).predictive_kernel, Xs, ratio=0.5)
    idx = sampler.sir_indices(w, 50, seed=1)
    lp = sampler.mvn_logpdf(Xs, torch.zeros(d), 2.0 * torch.eye(d))
    torch.cuda.synchronize()
    assert len(idx) == 50 and bool(torch.isfinite(lp).all())
    print("[sanitize] candidates: ok", flush=True)
if "gemm" in which:
    A = torch.randn(130, 300, generator=g, dtype=torch.float64).to(dev)
    B = torch.randn(300, 200, generator=g, dtype=torch.float64).to(dev)
    C1 = ops.dgemm(A, B)
    C2 = ops.tgemm(A, B.T.contiguous())
    torch.cuda.synchronize()
    assert float((C1 - A @ B).abs().max()) < 1e-10 and float((C2 - A @ B).abs().max()) < 1e-2
    print("[sanitize] gemm: ok", flush=True)
    # symmetric Gram product (lower tiles + mirror, split K) and the two-K-part tgemm (more tiles than SMs)
    Y = torch.randn(3000, 200, generator=g, dtype=torch.float64).to(dev)
    G1 = ops.dgemm(Y, Y, True, False)
    A2 = torch.randn(1300, 512, generator=g, dtype=torch.float64).to(dev)
    B2 = torch.randn(7000, 512, generator=g, dtype=torch.float64).to(dev)
    C3 = ops.tgemm(A2, B2)
    torch.cuda.synchronize()
    assert torch.equal(G1, G1.T) and float((G1 - Y.T @ Y).abs().max()) < 1e-9
    assert float((C3 - A2 @ B2.T).abs().max()) < 5e-3
    print("[sanitize] gram dgemm / k-split tgemm: ok", flush=True)
if "host" in which:
    # host-buffer entries: library-drawn test matrix, staged shard, row-sharded basis (one shard), block-cache trim
    from basq_b200 import sharded
    Xh, Zh = X.cpu().pin_memory(), Z.cpu()
    kern = model.covar_module.forward
    check("recombine_host (device-drawn test matrix)", *ops.recombine_host(kern, Xh, Zh, n - 1, seed=5))
    Ush = sharded.nystrom_basis_sharded(kern, Z, n - 1, seed=5)
    staged = ops.stage_candidates(Xh, device=dev)
    check("staged shard + sharded basis", *sharded.recombination_sharded(None, Z, n, kern, N, 0, Ush, staged=staged))
    ops.release_memory(dev)
    check("after trim", *ops.recombine(kern, X, Z, U))
print("[sanitize] done", flush=True)
