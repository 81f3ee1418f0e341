"""Oracle of the candidate-side stages (oracle/sampler.py) against known answers: Random123's
Philox4x32-10 vectors, torch's MultivariateNormal.log_prob (what the reference calls), the
reference's own cleansing_weights outputs (tests/golden/candidates.npz), and the calc_weights / lfi
formulas.  CPU only."""
import math
import os

import numpy as np
import torch

from oracle import sampler as osam


def test_philox_known_answers():
    """kat_vectors of the Random123 distribution, philox4x32 with 10 rounds."""
    def run(ctr, key):
        r = osam.philox4x32_10(np.array([ctr], dtype=np.uint32), np.array([key], dtype=np.uint32))[0]
        return [int(v) for v in r]
    assert run([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert run([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert run([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def _prior(d, seed=0):
    g = torch.Generator().manual_seed(seed)
    A = torch.randn(d, d, generator=g, dtype=torch.float64)
    cov = A @ A.T / d + 0.5 * torch.eye(d, dtype=torch.float64)
    mean = torch.randn(d, generator=g, dtype=torch.float64)
    return mean, cov, torch.linalg.cholesky(cov)


def test_logpdf_matches_torch_mvn():
    mean, cov, L = _prior(6)
    mvn = torch.distributions.MultivariateNormal(mean, cov)
    X = mvn.sample(torch.Size([500]))
    ref = mvn.log_prob(X).numpy()
    np.testing.assert_allclose(osam.mvn_logpdf(X.numpy(), mean.numpy(), L.numpy()), ref, rtol=1e-12, atol=1e-12)


def test_sample_moments_and_stream_slicing():
    mean, cov, L = _prior(5, seed=2)
    X = osam.sample_mvn(mean.numpy(), L.numpy(), 200_000, seed=42)
    assert np.abs(X.mean(0) - mean.numpy()).max() < 0.02
    assert np.abs(np.cov(X.T) - cov.numpy()).max() < 0.03
    # a shard is a slice of the one stream
    part = osam.sample_mvn(mean.numpy(), L.numpy(), 1000, seed=42, offset=150_000)
    assert np.array_equal(part, X[150_000:151_000])
    assert not np.array_equal(osam.sample_mvn(mean.numpy(), L.numpy(), 10, seed=43), X[:10])


def test_cleansing_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "candidates.npz"))
    for tag in ("mixed", "zero", "plain"):
        np.testing.assert_allclose(osam.cleansing_weights(g[f"in_{tag}"]), g[f"out_{tag}"], rtol=1e-15, atol=0)


def test_calc_weights_prior_cancels():
    """BASQ/_sampler.py:200-216 multiplies the prior density into f and g; the ratio is
    |m| / (r v + (1 - r) |m|), which is what the device kernel evaluates."""
    g = np.random.default_rng(0)
    m, v = g.normal(size=300), g.uniform(0.01, 2.0, size=300)
    lp = g.normal(-8.0, 2.0, size=300)
    for r in (0.0, 0.3, 0.5, 1.0):
        w = osam.calc_weights(m, v, lp, r)
        direct = np.abs(m) / (r * v + (1 - r) * np.abs(m)) if r < 1 else np.abs(m) / (r * v)
        np.testing.assert_allclose(w, direct / direct.sum(), rtol=1e-10)


def test_lfi_is_a_normal_cdf():
    m, v = np.array([1.0, 2.0, -1.0]), np.array([1.0, 4.0, 0.25])
    ref = torch.distributions.Normal(0.0, 1.0).cdf(torch.tensor((m - 1.0) / np.sqrt(v))).numpy()
    np.testing.assert_allclose(osam.lfi(m, v), ref, rtol=1e-12)
    np.testing.assert_allclose(osam.lfi(m, v, log=True), np.log(ref + torch.finfo().eps), rtol=1e-12)


def test_sir_race_is_proportional_without_replacement():
    """The oracle's exponential race behaves like torch.multinomial(weights, n) without replacement:
    no repeats, zero weights never drawn, first-draw frequencies proportional to the weights."""
    w = np.array([0.0, 1.0, 2.0, 5.0, 0.0, 2.0])
    firsts = np.zeros(6)
    for s in range(3000):
        d = osam.sir_indices(w, 3, seed=s)
        assert len(set(d.tolist())) == 3 and w[d].min() > 0
        firsts[d[0]] += 1
    np.testing.assert_allclose(firsts / 3000, w / w.sum(), atol=0.03)
    assert len(osam.sir_indices(w, 10, seed=1)) == 4
