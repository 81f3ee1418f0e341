"""GPU parity tests of the building blocks, through the C ABI (basq_b200.ops -> libbasq_b200.so),
against the CPU oracle on the same seeded inputs.  Run on the B200 box: pytest -m gpu."""
import math
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import gp_kernels as ogp  # noqa: E402
from oracle import rchq as orchq  # noqa: E402


@pytest.fixture(scope="module")
def bq():
    import basq_b200
    from basq_b200 import _lib, ops
    return basq_b200, _lib, ops


DEV = "cuda"


def rel(a, b):
    a = a.detach().double().cpu(); b = b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-300))


# ------------------------------------------------------------------------------------------- dgemm
@pytest.mark.parametrize("ta,tb", [(False, False), (True, False), (False, True), (True, True)])
@pytest.mark.parametrize("m,n,k", [(1, 1, 1), (65, 33, 17), (128, 64, 256), (200, 130, 1001), (999, 2000, 130),
                                   (300, 999, 64), (64, 2002, 48)])
def test_dgemm(bq, ta, tb, m, n, k):
    _, _, ops = bq
    g = torch.Generator().manual_seed(m * 7 + n)
    A = torch.randn((k, m) if ta else (m, k), generator=g, dtype=torch.float64).to(DEV)
    B = torch.randn((n, k) if tb else (k, n), generator=g, dtype=torch.float64).to(DEV)
    C = ops.dgemm(A, B, ta, tb)
    ref = (A.T if ta else A) @ (B.T if tb else B)
    assert rel(C, ref) < 1e-13


@pytest.mark.parametrize("n,k,ta", [(1, 5, True), (64, 40, True), (65, 3000, True), (130, 517, False),
                                    (999, 10000, True), (400, 9000, False)])
def test_dgemm_gram_is_symmetric(bq, n, k, ta):
    """A^T A / A A^T (the CholeskyQR Gram matrix): only the tiles touching the lower triangle are computed,
    with the K range split, and the rest is mirrored - exactly symmetric, same values as the full product."""
    _, _, ops = bq
    g = torch.Generator().manual_seed(n + k)
    A = torch.randn((k, n) if ta else (n, k), generator=g, dtype=torch.float64).to(DEV)
    C = ops.dgemm(A, A, ta, not ta)
    ref = A.T @ A if ta else A @ A.T
    assert C.shape == (n, n) and torch.equal(C, C.T)
    assert rel(C, ref) < 1e-13


@pytest.mark.parametrize("m,n,k", [(1, 1, 1), (65, 33, 17), (128, 256, 32), (129, 257, 33), (300, 700, 1001),
                                   (999, 640, 2048)])
def test_tgemm_split_fp16(bq, m, n, k):
    """Tensor-core GEMM (tcgen05, row-scaled fp16 hi / lo split, three products): fp32-level accuracy against
    an fp64 product."""
    _, _, ops = bq
    g = torch.Generator().manual_seed(m * 11 + n)
    A = torch.randn(m, k, generator=g, dtype=torch.float64).to(DEV)
    B = torch.randn(n, k, generator=g, dtype=torch.float64).to(DEV)
    C = ops.tgemm(A, B)
    ref = A @ B.T
    # error relative to the magnitude of the sum's terms: the split drops the lo*lo product (2^-22) and
    # the accumulator in TMEM is fp32 (growth with k like an fp32 GEMM)
    scale = float((A.abs() @ B.abs().T).max())
    assert float((C - ref).abs().max()) / scale < 2e-6 * max(1.0, k / 512)
    # fp32-representable inputs with small integer values are reproduced exactly
    Ai = torch.randint(-8, 9, (m, k), generator=g).double().to(DEV)
    Bi = torch.randint(-8, 9, (n, k), generator=g).double().to(DEV)
    assert torch.equal(ops.tgemm(Ai, Bi), Ai @ Bi.T)


@pytest.mark.parametrize("m,n,k", [(999, 10000, 2000), (640, 2100, 700), (1000, 4000, 64)])
def test_tgemm_ksplit_and_row_ranges(bq, m, n, k):
    """More tiles than SMs and not a multiple of them: every tile is computed as two K parts that are added onto
    the zero-filled output (two addends: order-independent).  Rows and columns with very different magnitudes
    exercise the per-row power-of-two scales."""
    _, _, ops = bq
    g = torch.Generator().manual_seed(m + n + k)
    A = torch.randn(m, k, generator=g, dtype=torch.float64) * torch.logspace(-6, 6, m, dtype=torch.float64).unsqueeze(1)
    B = torch.randn(n, k, generator=g, dtype=torch.float64) * torch.logspace(3, -9, n, dtype=torch.float64).unsqueeze(1)
    A[5] = 0.0                                   # an all-zero row keeps scale 1
    A, B = A.to(DEV), B.to(DEV)
    C = ops.tgemm(A, B)
    ref = A @ B.T
    scale = A.abs() @ B.abs().T
    err = ((C - ref).abs() / scale.clamp_min(1e-300)).max()
    assert float(err) < 2e-6 * max(1.0, k / 512)
    assert torch.equal(C, ops.tgemm(A, B))       # run-to-run identical


# ------------------------------------------------------------------------------------------- base kernels
FAMS = [("rbf", 2.5, 0), ("matern", 1.5, 1), ("matern", 2.5, 2)]


def _spec(bq, fam_id, ls, os_=1.3):
    _, _lib, _ = bq
    from basq_b200.kernels import KernelSpec
    return KernelSpec(fam_id, _lib.PLAIN, torch.as_tensor(ls, dtype=torch.float64), os_)


@pytest.mark.parametrize("fam,nu,fam_id", FAMS)
@pytest.mark.parametrize("d", [1, 2, 3, 7, 10, 13, 20, 32])
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.float64, 1e-12)])
def test_gram_plain(bq, fam, nu, fam_id, d, dtype, tol):
    _, _, ops = bq
    g = torch.Generator().manual_seed(100 + d)
    X = (math.sqrt(2.0) * torch.randn(157, d, generator=g, dtype=torch.float64) + 0.3)
    Y = (math.sqrt(2.0) * torch.randn(211, d, generator=g, dtype=torch.float64) + 0.3)
    ls = 1.5 * math.sqrt(d / 2.0)
    K = ops.gram(_spec(bq, fam_id, ls), X.to(DEV, dtype), Y.to(DEV, dtype))
    ref = ogp.base_kernel(X.to(dtype).double(), Y.to(dtype).double(), fam, ls, 1.3, nu=nu)
    assert K.shape == (157, 211) and K.dtype == torch.float64
    assert rel(K, ref) < tol


def test_gram_ard_lengthscale(bq):
    _, _, ops = bq
    g = torch.Generator().manual_seed(3)
    X = torch.randn(50, 4, generator=g, dtype=torch.float64)
    ls = torch.tensor([0.7, 1.1, 2.0, 3.5], dtype=torch.float64)
    K = ops.gram(_spec(bq, 0, ls), X.to(DEV), X.to(DEV))
    assert rel(K, ogp.base_kernel(X, X, "rbf", ls, 1.3)) < 1e-12


# ------------------------------------------------------------------------------------------- posterior kernels
def _gp(d=4, n_obs=40, family="rbf", noise=1e-6, log_targets=False, mean_const=0.0, nu=2.5):
    return ogp.make_gp(d, n_obs, family=family, lengthscale=1.7, outputscale=1.2, noise=noise, nu=nu,
                       mean_const=mean_const, seed=7, log_targets=log_targets)


def _kernel_objects(model):
    return {
        "pred_cov": ogp.VanillaGP(model).predictive_kernel,
        "wsabil": ogp.WsabiGP(model, alpha=0.05).wsabil_kernel,
        "wsabim": ogp.WsabiGP(model, alpha=0.05).wsabim_kernel,
        "mmlt": ogp.ScaleMmltGP(model).gspace_kernel,
        "sober_pc": ogp.Kernel(model, "predictive_covariance"),
        "sober_w": ogp.Kernel(model, "weighted_predictive_covariance"),
        "sober_k": ogp.Kernel(model, "kernel"),
        "forward": model.covar_module.forward,
    }


@pytest.mark.parametrize("name", ["pred_cov", "wsabil", "wsabim", "mmlt", "sober_pc", "sober_w", "sober_k", "forward"])
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-4), (torch.float64, 1e-9)])
def test_gram_modes(bq, name, dtype, tol):
    _, _, ops = bq
    # fp32 kernel values carry ~1e-7 relative noise which W = (K + s2 I)^-1 amplifies by its
    # condition number: the fp32 case uses a well-conditioned GP, the fp64 case a stiff one
    model = _gp(noise=1e-2 if dtype == torch.float32 else 1e-6, log_targets=(name == "mmlt"),
                mean_const=0.1 if name.startswith("wsabi") else 0.0)
    kern = _kernel_objects(model)[name]
    g = torch.Generator().manual_seed(21)
    X = math.sqrt(2.0) * torch.randn(90, 4, generator=g, dtype=torch.float64)
    Y = math.sqrt(2.0) * torch.randn(120, 4, generator=g, dtype=torch.float64)
    K = ops.gram(kern, X.to(DEV, dtype), Y.to(DEV, dtype))
    ref = kern(X.to(dtype).double(), Y.to(dtype).double())
    assert rel(K, ref) < tol


def test_gram_basq_noise_diag_quirk(bq):
    """BASQ/_gp.py:275-276 adds lik_var to the first min(len) diagonal entries."""
    _, _, ops = bq
    model = _gp(noise=1e-3)
    kern = ogp.VanillaGP(model, add_noise_diag=True).predictive_kernel
    g = torch.Generator().manual_seed(5)
    X = torch.randn(30, 4, generator=g, dtype=torch.float64)
    K = ops.gram(kern, X.to(DEV), X[:20].to(DEV))
    assert rel(K, kern(X, X[:20])) < 1e-9


@pytest.mark.parametrize("family,nu", [("rbf", 2.5), ("matern", 2.5)])
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-4), (torch.float64, 1e-9)])
def test_gp_predict(bq, family, nu, dtype, tol):
    _, _, ops = bq
    model = _gp(family=family, nu=nu, mean_const=0.2, noise=1e-2 if dtype == torch.float32 else 1e-6)
    g = torch.Generator().manual_seed(9)
    X = math.sqrt(2.0) * torch.randn(3001, 4, generator=g, dtype=torch.float64)
    kern = ogp.VanillaGP(model).predictive_kernel
    mean, var = ops.gp_predict(kern, X.to(DEV, dtype))
    m_ref, v_ref = ogp.predict(X.to(dtype).double(), model)
    assert rel(mean, m_ref) < tol
    assert float((var.cpu() - v_ref).abs().max()) < tol * 1.2


def test_gp_predict_model_space(bq):
    _, _, ops = bq
    model = _gp(mean_const=0.1)
    g = torch.Generator().manual_seed(10)
    X = math.sqrt(2.0) * torch.randn(500, 4, generator=g, dtype=torch.float64)
    wl = ogp.WsabiGP(model, alpha=0.05)
    for kern, ref in [(wl.wsabil_kernel, wl.wsabil_predict(X)), (wl.wsabim_kernel, wl.wsabim_predict(X))]:
        mean, var = ops.gp_predict(kern, X.to(DEV), space=1)
        assert rel(mean, ref[0]) < 1e-9 and float((var.cpu() - ref[1]).abs().max()) < 1e-9
    mm = ogp.ScaleMmltGP(_gp(log_targets=True))
    mean, var = ops.gp_predict(mm.gspace_kernel, X.to(DEV), space=1)
    ref = mm.gspace_predict(X)
    assert rel(mean, ref[0]) < 1e-9 and float((var.cpu() - ref[1]).abs().max()) < 1e-9


# ------------------------------------------------------------------------------------------- Caratheodory
def _car_check(X, mu, w, idx, tol):
    A = torch.cat([torch.ones(len(X), 1, dtype=torch.float64), X.double()], 1)
    full = A.T @ mu.double()
    red = A[idx.cpu()].T @ w.double().cpu()
    res = float(torch.linalg.norm(full - red) / torch.linalg.norm(full))
    assert len(idx) <= X.shape[1] + 1, (len(idx), X.shape)
    assert bool((w > 0).all())
    assert bool((idx[1:] > idx[:-1]).all())
    assert res < tol, res
    return res


@pytest.mark.parametrize("tag", ["a", "b", "c", "d"])
def test_car_golden_inputs(bq, golden_dir, tag):
    basq_b200, _, _ = bq
    g = np.load(os.path.join(golden_dir, "car_kat.npz"))
    X = torch.from_numpy(g[f"X_{tag}"]); mu = torch.from_numpy(g[f"mu_{tag}"])
    w, idx, *_ = basq_b200.Tchernychova_Lyons_CAR(X.to(DEV), mu.to(DEV), DEV)
    _car_check(X, mu, w, idx, 1e-12)
    # the reference's own answer has the same size (generic position => exactly n+1 points)
    assert len(idx) == len(g[f"idx_{tag}"])


@pytest.mark.parametrize("S,q", [(3, 1), (10, 8), (11, 9), (64, 31), (200, 99), (1000, 499), (2000, 999), (777, 300)])
def test_car_random(bq, S, q):
    basq_b200, _, _ = bq
    g = torch.Generator().manual_seed(S + q)
    X = torch.randn(S, q, generator=g, dtype=torch.float64) * torch.logspace(0, -6, q, dtype=torch.float64)
    mu = torch.rand(S, generator=g, dtype=torch.float64) + 0.01
    mu /= mu.sum()
    w, idx, *_ = basq_b200.Tchernychova_Lyons_CAR(X.to(DEV), mu.to(DEV), DEV)
    _car_check(X, mu, w, idx, 1e-10)
    assert abs(float(w.sum()) - 1.0) < 1e-12


def test_car_rank_deficient_and_trivial(bq):
    basq_b200, _, _ = bq
    g = torch.Generator().manual_seed(1)
    X = torch.randn(40, 12, generator=g, dtype=torch.float64)
    X[:, 5] = X[:, 2] * 2.0 - X[:, 3]        # dependent feature
    X[7] = X[3]                               # duplicated point
    mu = torch.rand(40, generator=g, dtype=torch.float64) + 0.1
    w, idx, *_ = basq_b200.Tchernychova_Lyons_CAR(X.to(DEV), mu.to(DEV), DEV)
    _car_check(X, mu, w, idx, 1e-11)
    # nothing to do when there are at most n+1 points
    w2, idx2, *_ = basq_b200.Tchernychova_Lyons_CAR(X[:10].to(DEV), mu[:10].to(DEV), DEV)
    assert len(idx2) == 10 and torch.allclose(w2.cpu(), mu[:10])


def test_car_large_rank_deficient_falls_back(bq):
    """S = 2000 sets but only ~500 independent features: more non-basic sets than the
    register-resident kernel can hold -> the general kernel takes over (car.cu)."""
    basq_b200, _, _ = bq
    g = torch.Generator().manual_seed(4)
    S, q, r = 2000, 999, 500
    base = torch.randn(S, r, generator=g, dtype=torch.float64)
    mix = torch.randn(r, q - r, generator=g, dtype=torch.float64) / math.sqrt(r)
    X = torch.cat([base, base @ mix], 1)
    mu = torch.rand(S, generator=g, dtype=torch.float64) + 0.01
    mu /= mu.sum()
    w, idx, *_ = basq_b200.Tchernychova_Lyons_CAR(X.to(DEV), mu.to(DEV), DEV)
    A = torch.cat([torch.ones(S, 1, dtype=torch.float64), X], 1)
    res = float(torch.linalg.norm(A.T @ mu - A[idx.cpu()].T @ w.cpu()) / torch.linalg.norm(A.T @ mu))
    assert res < 1e-9 and len(idx) <= r + 1 and bool((w > 0).all())


# ------------------------------------------------------------------------------------------- features
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.float64, 1e-11)])
@pytest.mark.parametrize("name", ["forward", "pred_cov", "wsabil", "wsabim", "mmlt"])
def test_features(bq, name, dtype, tol):
    _, _, ops = bq
    model = _gp(noise=1e-2 if dtype == torch.float32 else 1e-6, log_targets=(name == "mmlt"))
    kern = _kernel_objects(model)[name]
    g = torch.Generator().manual_seed(33)
    X = math.sqrt(2.0) * torch.randn(1500, 4, generator=g, dtype=torch.float64)
    Z = math.sqrt(2.0) * torch.randn(60, 4, generator=g, dtype=torch.float64)
    U = torch.linalg.qr(torch.randn(60, 11, generator=g, dtype=torch.float64)).Q.T.contiguous()
    Phi = ops.features(kern, X.to(DEV, dtype), Z.to(DEV, dtype), U.to(DEV))
    ref = orchq.features(X.to(dtype).double(), U, Z.to(dtype).double(), kern)
    tol_eff = tol if name == "forward" or dtype == torch.float64 else 3e-4
    assert rel(Phi, ref) < tol_eff


# ------------------------------------------------------------------------------------------- Nystrom
@pytest.mark.parametrize("M,q", [(64, 9), (300, 99), (700, 200)])
def test_nystrom_basis(bq, M, q):
    _, _, ops = bq
    g = torch.Generator().manual_seed(M)
    Z = math.sqrt(2.0) * torch.randn(M, 5, generator=g, dtype=torch.float64)
    spec = _spec(bq, 0, 2.0, 1.0)
    torch.manual_seed(0)
    S, U = ops.nystrom_basis(spec, Z.to(DEV), q)
    U = U.cpu()
    assert U.shape == (q, M)
    assert float((U @ U.T - torch.eye(q, dtype=torch.float64)).abs().max()) < 1e-10
    K = ogp.base_kernel(Z, Z, "rbf", 2.0, 1.0)
    # captured energy vs. the reference's randomised basis and vs. the optimum
    torch.manual_seed(0)
    _, Uref = orchq.nystrom_basis(Z, q, lambda a, b: ogp.base_kernel(a, b, "rbf", 2.0, 1.0))
    err = lambda B: float(torch.linalg.norm(K - K @ B.T @ B))
    ev = torch.linalg.eigvalsh(K).flip(0)
    opt = float(torch.sqrt((ev[q:] ** 2).sum()))
    assert err(U) <= 1.5 * err(Uref) + 1e-9 * float(torch.linalg.norm(K))
    assert err(U) <= 3.0 * opt + 1e-9 * float(torch.linalg.norm(K))
    # Rayleigh quotients are the diagonal of U K U^T
    assert rel(S, torch.diagonal(U @ K @ U.T)) < 1e-9


@pytest.mark.parametrize("M,q,d,ls", [(300, 99, 5, 2.0), (700, 200, 5, 2.0), (1500, 99, 2, 1.0), (2000, 499, 10, 2.5)])
def test_nystrom_basis_fp32_tensor_path(bq, M, q, d, ls):
    """fp32 kernels: the subspace iteration's K Y products run on the tensor cores (tgemm, 3xTF32);
    the basis stays orthonormal to fp64 and captures what the reference's fp32 svd_lowrank captures
    (d = 2: fast spectral decay, numerically rank-deficient Gram matrix)."""
    _, _, ops = bq
    g = torch.Generator().manual_seed(M + q)
    Z = (math.sqrt(2.0) * torch.randn(M, d, generator=g)).float()
    spec = _spec(bq, 0, ls, 1.0)
    torch.manual_seed(0)
    S, U = ops.nystrom_basis(spec, Z.to(DEV), q)
    U = U.cpu()
    assert U.shape == (q, M) and U.dtype == torch.float64
    assert float((U @ U.T - torch.eye(q, dtype=torch.float64)).abs().max()) < 1e-10
    K = ogp.base_kernel(Z.double(), Z.double(), "rbf", ls, 1.0)
    torch.manual_seed(0)
    _, Uref = orchq.nystrom_basis(Z, q, lambda a, b: ogp.base_kernel(a, b, "rbf", ls, 1.0))   # fp32, as BASQ runs it
    Uref = torch.linalg.qr(Uref.double().T).Q.T
    err = lambda B: float(torch.linalg.norm(K - K @ B.T @ B))
    ev = torch.linalg.eigvalsh(K).flip(0)
    opt = float(torch.sqrt((ev[q:] ** 2).sum()))
    nK = float(torch.linalg.norm(K))
    assert err(U) <= 1.5 * err(Uref) + 1e-6 * nK
    assert err(U) <= 3.0 * opt + 1e-6 * nK
    assert rel(S, torch.diagonal(U @ K @ U.T)) < 1e-5
