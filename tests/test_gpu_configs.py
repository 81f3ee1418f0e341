"""BASELINE.json configurations 2-5 at their FULL sizes on the device, through size-independent
properties (the CPU oracle cannot finish these): <= n points, ascending unique indices, w > 0,
sum w = 1, and moment preservation for a random subset of the Nystrom test functions, evaluated in
fp64 by the library's feature kernel by streaming over ALL candidates.  Candidates come from the
device sampler (csrc/candidates.cu), as in a BASQ iteration.  GP posterior moments over the candidate
set (config 5's acquisition pass) are spot-checked against the oracle on a sample."""
import math

import numpy as np
import pytest
import torch

from oracle import gp_kernels as ogp

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


@pytest.fixture(scope="module")
def lib():
    import basq_b200
    from basq_b200 import _lib, gp, ops, sampler
    from basq_b200.kernels import spec_from_model
    return basq_b200, _lib, gp, ops, sampler, spec_from_model


def _observations(d, n_obs, seed, log=False):
    g = torch.Generator().manual_seed(seed)
    X = math.sqrt(2.0) * torch.randn(n_obs, d, generator=g, dtype=torch.float64)
    centres = 1.5 * torch.randn(3, d, generator=g, dtype=torch.float64)
    y = sum(torch.exp(-0.25 * ((X - c) ** 2).sum(-1)) for c in centres) / 3.0
    return X, (torch.log(y + 1e-12) if log else y)


def _check(lib, kern, X, Z, n, n_check=24, batch=250_000, tol=1e-8):
    basq_b200, _lib, gp, ops, sampler, _ = lib
    N, M = len(X), len(Z)
    torch.manual_seed(0)
    _, U = ops.nystrom_basis(kern, Z, n - 1, want_S=False)
    idx, w = ops.recombine(kern, X, Z, U)
    torch.cuda.synchronize()
    idx_c, w_c = idx.cpu(), w.cpu()
    assert 1 <= len(idx_c) <= n and bool((w_c > 0).all())
    assert bool((idx_c[1:] > idx_c[:-1]).all()) and int(idx_c.min()) >= 0 and int(idx_c.max()) < N
    assert abs(float(w_c.sum()) - 1.0) < 1e-11
    g = torch.Generator().manual_seed(1)
    rows = torch.randperm(n - 1, generator=g)[:n_check].sort().values
    Us = U[rows.to(DEV)].contiguous()
    full = torch.zeros(len(rows), dtype=torch.float64, device=DEV)
    for i in range(0, N, batch):
        full += ops.features(kern, X[i:i + batch], Z, Us).sum(0)
    full /= N
    red = ops.features(kern, X[idx], Z, Us).T @ w
    res = float(torch.linalg.norm(full - red) / torch.linalg.norm(full))
    assert res < tol, res
    return idx, w


def test_config2_tutorial01_2d(lib):
    """Tutorial 01 shape: d = 2, N_rec = 1e6, N_nys = 1e4, batch n = 100, VBQ posterior covariance."""
    basq_b200, _lib, gp, ops, sampler, spec_from_model = lib
    d, N, M, n = 2, 1_000_000, 10_000, 100
    Xo, yo = _observations(d, 102, seed=3)
    model = gp.FixedGP(Xo.to(DEV, torch.float32), yo.to(DEV), gp.ScaleKernel(gp.RBFKernel(1.0), 1.0), noise=1e-4)
    kern = spec_from_model(model, _lib.PRED_COV)
    X = sampler.sample_mvn(torch.zeros(d), 2.0 * torch.eye(d), N, seed=21, device=DEV)
    _check(lib, kern, X, X[:M].clone(), n)


def test_config4_matern52_20d(lib):
    """Tutorial 02 shape: 20-D Matern-5/2 kernel, N_rec = 4e6, batch n = 500 (M = 5e3)."""
    basq_b200, _lib, gp, ops, sampler, spec_from_model = lib
    from basq_b200.kernels import KernelSpec
    d, N, M, n = 20, 4_000_000, 5_000, 500
    kern = KernelSpec(_lib.MATERN25, _lib.PLAIN, torch.tensor([4.0]), 1.0)
    X = sampler.sample_mvn(torch.zeros(d), 2.0 * torch.eye(d), N, seed=22, device=DEV)
    _check(lib, kern, X, X[:M].clone(), n)


def test_config5_wsabil_1e7(lib):
    """Tutorial 03 shape: 10-D, N_rec = 1e7, batch n = 1000, WSABI-L kernel on a GP with 1002 observations."""
    basq_b200, _lib, gp, ops, sampler, spec_from_model = lib
    d, N, M, n = 10, 10_000_000, 10_000, 1000
    Xo, yo = _observations(d, 1002, seed=5)
    model = gp.FixedGP(Xo.to(DEV, torch.float32), torch.sqrt(2.0 * yo).to(DEV), gp.ScaleKernel(gp.RBFKernel(2.5), 1.0),
                       noise=1e-4)
    kern = spec_from_model(model, _lib.WSABI_L, offset=0.0)
    X = sampler.sample_mvn(torch.zeros(d), 2.0 * torch.eye(d), N, seed=23, device=DEV)
    _check(lib, kern, X, X[:M].clone(), n, n_check=16)


def test_config5_mmlt(lib):
    """MMLT (log-likelihood modelling) kernel, the pairwise non-linear path: N_rec = 5e5 (the chunked
    path costs 2 n_obs M flop per candidate and round; 1e7 of them take minutes), n = 200."""
    basq_b200, _lib, gp, ops, sampler, spec_from_model = lib
    d, N, M, n = 10, 500_000, 2_000, 200
    Xo, yo = _observations(d, 202, seed=6, log=True)
    model = gp.FixedGP(Xo.to(DEV, torch.float32), (yo - yo.max()).to(DEV), gp.ScaleKernel(gp.RBFKernel(2.5), 1.0),
                       noise=1e-4)
    kern = spec_from_model(model, _lib.MMLT_G)
    X = sampler.sample_mvn(torch.zeros(d), 2.0 * torch.eye(d), N, seed=24, device=DEV)
    _check(lib, kern, X, X[:M].clone(), n, n_check=16)


def test_config5_gp_moments_over_1e7_candidates(lib):
    """GP posterior mean / variance over 1e7 candidates (the acquisition pass of config 5), then the
    importance weights they define; spot-checked against the oracle GP on 2000 of the candidates."""
    basq_b200, _lib, gp, ops, sampler, spec_from_model = lib
    d, N, n_obs = 10, 10_000_000, 1002
    Xo, yo = _observations(d, n_obs, seed=5)
    model = gp.FixedGP(Xo.to(DEV, torch.float32), yo.to(DEV), gp.ScaleKernel(gp.RBFKernel(2.5), 1.0), noise=1e-4)
    kern = spec_from_model(model, _lib.PRED_COV)
    X = sampler.sample_mvn(torch.zeros(d), 2.0 * torch.eye(d), N, seed=25, device=DEV)
    mean, var = ops.gp_predict(kern, X, space=0, want_var=True)
    assert mean.shape == (N,) and bool(torch.isfinite(mean).all()) and bool((var > 0).all())
    pick = torch.randperm(N, generator=torch.Generator().manual_seed(2))[:2000]
    omodel = ogp.ExactGP(Xo, yo, ogp.ScaleKernel(ogp.RBFKernel(2.5), 1.0), noise=1e-4)
    m_o, v_o = ogp.predict(X[pick.to(DEV)].cpu().double(), omodel)
    assert float((mean[pick.to(DEV)].cpu() - m_o).abs().max()) < 2e-5 * float(m_o.abs().max() + 1.0)
    assert float((var[pick.to(DEV)].cpu() - v_o).abs().max()) < 2e-4 * float(v_o.abs().max())
    w = sampler.calc_weights(kern, X, ratio=0.5)
    assert abs(float(w.sum()) - 1.0) < 1e-10 and bool((w >= 0).all())
    # weighted candidates feed recombination as init_weights (SOBER signature)
    idx, wq = ops.recombine(kern, X[:2_000_000], X[:2000].clone(),
                            ops.nystrom_basis(kern, X[:2000].clone(), 99, want_S=False)[1], mu=w[:2_000_000])
    assert 1 <= len(idx) <= 100 and abs(float(wq.sum()) - float(w[:2_000_000].sum())) < 1e-12
