"""The oracle restatement vs. golden vectors produced by the reference's own BASQ/_rchq.py
(oracle/make_golden.py).  CPU only."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import rchq
from oracle.gp_kernels import base_kernel


def _kernel(fam, ls):
    family = "rbf" if fam == 0 else "matern"
    return lambda x, y: base_kernel(x, y, family, ls, 1.0, nu=2.5)


@pytest.mark.parametrize("tag", ["a", "b", "c", "d"])
def test_car_matches_reference(golden_dir, tag):
    g = np.load(os.path.join(golden_dir, "car_kat.npz"))
    X = torch.from_numpy(g[f"X_{tag}"]); mu = torch.from_numpy(g[f"mu_{tag}"])
    w, idx = rchq.caratheodory(X, mu)
    assert np.array_equal(idx.numpy(), g[f"idx_{tag}"])
    np.testing.assert_allclose(w.numpy(), g[f"w_{tag}"], rtol=1e-9, atol=1e-14)
    # invariants: mass, barycentre, positivity, size
    A = torch.cat([torch.ones(len(X), 1, dtype=X.dtype), X], 1)
    assert len(w) <= X.shape[1] + 1 and bool((w > 0).all())
    res = torch.linalg.norm(A.T @ mu - A[idx].T @ w) / torch.linalg.norm(A.T @ mu)
    assert float(res) < 1e-12


TL_TAGS = ["rbf_d3", "rbf_d10", "m52_d5", "final_only", "trivial", "exact_mult"]


@pytest.mark.parametrize("tag", TL_TAGS)
def test_tl_matches_reference(golden_dir, tag):
    g = np.load(os.path.join(golden_dir, "tl_cases.npz"))
    X = torch.from_numpy(g[f"X_{tag}"]); Z = torch.from_numpy(g[f"Z_{tag}"]); U = torch.from_numpy(g[f"U_{tag}"])
    N, d, M, n, fam, ls = g[f"meta_{tag}"]
    kern = _kernel(int(fam), float(ls))
    w, idx = rchq.tchernychova_lyons(X, U, Z, kern, chunk=1)
    assert len(idx) <= int(n) and bool((w > 0).all())
    assert abs(float(w.sum()) - 1.0) < 1e-12
    Phi = rchq.features(X, U, Z, kern)
    mu = torch.full((len(X),), 1.0 / len(X), dtype=torch.float64)
    assert rchq.moment_residual(Phi, mu, idx, w) < 1e-11
    # same arithmetic, same call sequence -> same pivots as the reference
    assert np.array_equal(idx.numpy(), g[f"idx_{tag}"])
    np.testing.assert_allclose(w.numpy(), g[f"w_{tag}"], rtol=1e-8, atol=1e-14)
    # the golden result itself satisfies the invariant (pins the metric, not just the port)
    gi = torch.from_numpy(g[f"idx_{tag}"]); gw = torch.from_numpy(g[f"w_{tag}"])
    assert rchq.moment_residual(Phi, mu, gi, gw) < 1e-11


@pytest.mark.parametrize("tag", ["rbf_d3", "m52_d5"])
def test_tl_chunked_sum_is_equivalent(golden_dir, tag):
    """chunk>1 (used by the timed CPU baseline) changes only the summation order."""
    g = np.load(os.path.join(golden_dir, "tl_cases.npz"))
    X = torch.from_numpy(g[f"X_{tag}"]); Z = torch.from_numpy(g[f"Z_{tag}"]); U = torch.from_numpy(g[f"U_{tag}"])
    N, d, M, n, fam, ls = g[f"meta_{tag}"]
    kern = _kernel(int(fam), float(ls))
    w, idx = rchq.tchernychova_lyons(X, U, Z, kern, chunk=64)
    Phi = rchq.features(X, U, Z, kern)
    mu = torch.full((len(X),), 1.0 / len(X), dtype=torch.float64)
    assert rchq.moment_residual(Phi, mu, idx, w) < 1e-11
    assert len(idx) <= int(n)


def test_weighted_variant_preserves_moments():
    """SOBER-style honoured init_weights with 30% zeros, without SOBER's tail bug."""
    torch.manual_seed(5)
    N, d, M, n = 2000, 4, 40, 10
    X = math.sqrt(2.0) * torch.randn(N, d, dtype=torch.float64)
    Z = X[:M].clone()
    kern = _kernel(0, 1.5)
    mu = torch.rand(N, dtype=torch.float64)
    mu[torch.rand(N) < 0.3] = 0.0
    mu = mu / mu.sum()
    _, U = rchq.nystrom_basis(Z, n - 1, kern)
    w, idx = rchq.tchernychova_lyons(X, U, Z, kern, mu=mu, chunk=8)
    Phi = rchq.features(X, U, Z, kern)
    assert rchq.moment_residual(Phi, mu, idx, w) < 1e-11
    assert bool((mu[idx] > 0).all())


def test_objective_variant_invariants():
    """The oracle's restatement of SOBER's calc_obj path (SOBER/_rchq.py:67-69,138-146,177-196): the
    q + 1 moments are preserved with <= q + 1 points, and the rule's expected objective is not
    below the measure's (every step moves along obj . w_null >= 0).  SOBER/_rchq.py itself
    double-counts the tail sums, so there is no reference output to compare with ("parity unpinned")."""
    g = torch.Generator().manual_seed(9)
    N, d, M, n = 1203, 3, 40, 8
    X = math.sqrt(2.0) * torch.randn(N, d, generator=g, dtype=torch.float64)
    Z = math.sqrt(2.0) * torch.randn(M, d, generator=g, dtype=torch.float64)
    U = torch.linalg.qr(torch.randn(M, n - 1, generator=g, dtype=torch.float64)).Q.T.contiguous()
    kern = _kernel(0, 1.3)
    calc_obj = lambda x: torch.exp(-0.5 * ((x - 0.7) ** 2).sum(-1))
    w, idx = rchq.tchernychova_lyons_objective(X, U, Z, kern, calc_obj)
    assert len(idx) <= n and bool((w > 0).all()) and abs(float(w.sum()) - 1.0) < 1e-12
    mu = torch.full((N,), 1.0 / N, dtype=torch.float64)
    Phi = rchq.features(X, U, Z, kern)
    assert rchq.moment_residual(Phi, mu, idx, w) < 1e-11
    assert float(w @ calc_obj(X[idx])) >= float(mu @ calc_obj(X)) - 1e-12
