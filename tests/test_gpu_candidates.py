"""Device candidate-side stages (csrc/candidates.cu) through the C ABI against the CPU oracle."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import gp_kernels as ogp
from oracle import sampler as osam

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


@pytest.fixture(scope="module")
def bs():
    from basq_b200 import sampler
    return sampler


def _prior(d, seed=0):
    g = torch.Generator().manual_seed(seed)
    A = torch.randn(d, d, generator=g, dtype=torch.float64)
    cov = A @ A.T / d + 0.5 * torch.eye(d, dtype=torch.float64)
    mean = torch.randn(d, generator=g, dtype=torch.float64)
    return mean, cov, torch.linalg.cholesky(cov)


@pytest.mark.parametrize("d", [1, 2, 5, 10, 32])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_sample_matches_oracle_stream(bs, d, dtype):
    mean, cov, L = _prior(d, seed=d)
    n, off, seed = 4097, (1 << 33) + 12345, 0xDEADBEEF12345
    X = bs.sample_mvn(mean, cov, n, seed=seed, offset=off, device=DEV, dtype=dtype).cpu().double().numpy()
    ref = osam.sample_mvn(mean.numpy(), L.numpy(), n, seed=seed, offset=off)
    tol = 2e-5 if dtype == torch.float32 else 1e-11
    assert np.abs(X - ref).max() <= tol * (1.0 + np.abs(ref).max())


@pytest.mark.parametrize("rows,cols", [(1000, 39), (257, 1), (64, 999)])
def test_standard_normals_match_oracle_stream(rows, cols):
    """basq_standard_normals (the library-drawn Nystrom test matrix): the sample_mvn stream with L = I and any
    number of columns, against oracle.sampler.standard_normals."""
    from basq_b200 import ops
    seed, off = 0xABCDEF987, 77
    out = ops.standard_normals(rows, cols, seed=seed, offset=off, device=DEV).cpu().numpy()
    ref = osam.standard_normals(seed, off, rows, cols)
    assert out.shape == ref.shape and np.abs(out - ref).max() < 1e-11
    assert ops.standard_normals(0, cols, device=DEV).shape == (0, cols)


def test_shards_are_slices_of_one_stream(bs):
    mean, cov, L = _prior(10, seed=1)
    full = bs.sample_mvn(mean, cov, 10_000, seed=7, device=DEV)
    a = bs.sample_mvn(mean, cov, 6_000, seed=7, offset=0, device=DEV)
    b = bs.sample_mvn(mean, cov, 4_000, seed=7, offset=6_000, device=DEV)
    assert torch.equal(torch.cat([a, b]), full)
    assert bs.sample_mvn(mean, cov, 0, seed=7, device=DEV).shape == (0, 10)


def test_sample_moments_large(bs):
    d = 10
    mean = torch.zeros(d, dtype=torch.float64); cov = 2.0 * torch.eye(d, dtype=torch.float64)   # main.py prior
    X = bs.sample_mvn(mean, cov, 4_000_000, seed=1, device=DEV).double()
    assert float(X.mean(0).abs().max()) < 5e-3
    assert float((torch.cov(X.T) - cov.to(DEV)).abs().max()) < 1e-2


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_logpdf_matches_torch_mvn(bs, dtype):
    mean, cov, L = _prior(7, seed=3)
    mvn = torch.distributions.MultivariateNormal(mean, cov)
    X = mvn.sample(torch.Size([3000])).to(dtype)
    ref = mvn.log_prob(X.double())
    out = bs.mvn_logpdf(X.to(DEV), mean, cov).cpu()
    assert float((out - ref).abs().max()) < 1e-10 * (1 + float(ref.abs().max()))


def _gp(d=4, n_obs=40, log_targets=False):
    return ogp.make_gp(d, n_obs, lengthscale=1.7, outputscale=1.2, noise=1e-4, seed=7, log_targets=log_targets)


@pytest.mark.parametrize("ratio", [0.0, 0.5, 1.0])
def test_calc_weights_matches_oracle(bs, ratio):
    model = _gp()
    kern = ogp.VanillaGP(model).predictive_kernel
    g = torch.Generator().manual_seed(5)
    X = math.sqrt(2.0) * torch.randn(5000, 4, generator=g, dtype=torch.float64)
    w = bs.calc_weights(kern, X.to(DEV), ratio=ratio).cpu().numpy()
    mean, var = ogp.predict(X, model)               # oracle GP (exact variance incl. noise)
    ref = osam.calc_weights(mean.numpy(), var.numpy(), np.zeros(len(X)), ratio)
    np.testing.assert_allclose(w, ref, rtol=1e-7, atol=1e-15)
    assert abs(w.sum() - 1.0) < 1e-12


def test_cleansing_matches_reference_golden(bs, golden_dir):
    g = np.load(os.path.join(golden_dir, "candidates.npz"))
    for tag in ("mixed", "zero", "plain"):
        out = bs.cleansing_weights(torch.from_numpy(g[f"in_{tag}"]).to(DEV)).cpu().numpy()
        np.testing.assert_allclose(out, g[f"out_{tag}"], rtol=1e-13, atol=0)


def test_prior_sampler_call_pattern(bs):
    """pts_nys, pts_rec, w = sampler(n_rec) as in BASQ/_sampler.py:21-34, then straight into recombination."""
    import basq_b200
    d = 3
    prior = torch.distributions.MultivariateNormal(torch.zeros(d), 2.0 * torch.eye(d))
    smp = bs.PriorSampler(prior, 20_000, 0.01, DEV, seed=11)
    pts_nys, pts_rec, w = smp(20_000)
    assert pts_rec.shape == (20_000, d) and pts_nys.shape == (200, d) and torch.equal(pts_nys, pts_rec[:200])
    assert abs(float(w.sum()) - 1.0) < 1e-4
    pts_nys2, pts_rec2, _ = smp(20_000)
    assert not torch.equal(pts_rec2, pts_rec)       # the stream continues
    cov = ogp.ScaleKernel(ogp.RBFKernel(1.5), 1.0)
    idx, wq = basq_b200.recombination(pts_rec, pts_nys, 20, cov.forward, DEV)
    assert 1 <= len(idx) <= 20 and bool((wq > 0).all())


def test_sir_matches_oracle_race_and_weights(bs):
    """UncertaintySampler.SIR (torch.multinomial without replacement): same draws as the oracle's
    exponential race on the same Philox stream, no repeats, zero-weight points never drawn, and
    inclusion frequencies proportional to the weights when n << N."""
    g = torch.Generator().manual_seed(3)
    N, n = 200_000, 1000
    w = torch.rand(N, generator=g, dtype=torch.float64) ** 3
    w[::5] = 0.0
    idx = bs.sir_indices(w.to(DEV), n, seed=99).cpu().numpy()
    ref = osam.sir_indices(w.numpy(), n, seed=99)
    assert np.array_equal(idx, ref)
    assert len(set(idx.tolist())) == n and bool((w[idx] > 0).all())
    # more draws than positive weights: every positive-weight point exactly once
    w2 = torch.zeros(5000, dtype=torch.float64); w2[torch.arange(0, 5000, 50)] = 1.0
    idx2 = bs.sir_indices(w2.to(DEV), 500, seed=1).cpu()
    assert sorted(idx2.tolist()) == list(range(0, 5000, 50))
    # frequencies: two weight classes 3 : 1, many independent seeds
    w3 = torch.ones(4000, dtype=torch.float64); w3[:2000] = 3.0
    hi = sum(int((bs.sir_indices(w3.to(DEV), 40, seed=s) < 2000).sum()) for s in range(200))
    assert abs(hi / (200 * 40) - 0.75) < 0.02
    Xs = bs.SIR(torch.arange(4000, dtype=torch.float32, device=DEV).unsqueeze(1), w3.to(DEV), 7, seed=5)
    assert Xs.shape == (7, 1)
