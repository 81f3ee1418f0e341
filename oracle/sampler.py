"""TEST INFRASTRUCTURE - CPU restatement (numpy) of the candidate-side stages of the reference.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.

* ``philox4x32_10``: Salmon, Moraes, Dror, Shaw, "Parallel random numbers: as easy as 1, 2, 3"
  (SC'11), the generator behind torch's CUDA ``prior.sample`` (BASQ/_sampler.py:31).  Pinned by the
  known-answer vectors of the Random123 distribution (tests/test_candidates_cpu.py).
* ``sample_mvn``: mean + L z with Box-Muller normals, the stream layout documented in
  include/basq_b200.h (basq_sample_mvn).  The reference draws with torch's own generator, so the
  VALUES are implementation-defined; what is checked against the reference's behaviour is the
  distribution (moments) and, bit-for-bit, the Philox integers.
* ``mvn_logpdf``: prior.log_prob (torch.distributions.MultivariateNormal.log_prob).
* ``calc_weights``: UncertaintySampler.calc_weights, BASQ/_sampler.py:190-217.  Pinned by outputs of
  that method itself (oracle/make_golden_gp.py -> tests/golden/gp_kernels.npz).
* ``lfi``: PI_BQ.lfi, SOBER/_pi.py:121-139.  log=False pinned the same way; the reference's log=True
  branch raises NameError (the file uses ``torch.finfo`` without importing torch), so the log form
  follows the line as written, with the fp32 eps torch.finfo() returns by default.
* ``cleansing_weights``: WeightsStabiliser.cleansing_weights, SOBER/_weights.py:21-38.
"""
from __future__ import annotations

import math

import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(ctr, key):
    """ctr [..., 4], key [..., 2] uint32 arrays -> [..., 4] uint32 (10 rounds)."""
    c = [np.asarray(ctr[..., i], dtype=np.uint32).copy() for i in range(4)]
    k = [np.asarray(key[..., i], dtype=np.uint32).copy() for i in range(2)]
    for _ in range(10):
        p0 = M0 * c[0].astype(np.uint64)
        p1 = M1 * c[2].astype(np.uint64)
        n0 = (p1 >> np.uint64(32)).astype(np.uint32) ^ c[1] ^ k[0]
        n1 = (p1 & MASK).astype(np.uint32)
        n2 = (p0 >> np.uint64(32)).astype(np.uint32) ^ c[3] ^ k[1]
        n3 = (p0 & MASK).astype(np.uint32)
        c = [n0, n1, n2, n3]
        with np.errstate(over="ignore"):
            k = [(k[0] + W0).astype(np.uint32), (k[1] + W1).astype(np.uint32)]
    return np.stack(c, axis=-1)


def standard_normals(seed, offset, n, d):
    """z [n, d] fp64: Box-Muller on the Philox stream (counter = (row lo, row hi, block, 0))."""
    rows = (np.arange(n, dtype=np.uint64) + np.uint64(offset))
    z = np.zeros((n, 4 * ((d + 3) // 4)), dtype=np.float64)
    key = np.empty((n, 2), dtype=np.uint32)
    key[:, 0] = np.uint32(seed & 0xFFFFFFFF)
    key[:, 1] = np.uint32((seed >> 32) & 0xFFFFFFFF)
    for blk in range((d + 3) // 4):
        ctr = np.zeros((n, 4), dtype=np.uint32)
        ctr[:, 0] = (rows & MASK).astype(np.uint32)
        ctr[:, 1] = (rows >> np.uint64(32)).astype(np.uint32)
        ctr[:, 2] = blk
        r = philox4x32_10(ctr, key)
        u = ((r >> np.uint32(8)).astype(np.float64) + 0.5) * 2.0 ** -24
        for h in range(2):
            rad = np.sqrt(-2.0 * np.log(u[:, 2 * h]))
            ang = 2.0 * np.pi * u[:, 2 * h + 1]
            z[:, 4 * blk + 2 * h] = rad * np.cos(ang)
            z[:, 4 * blk + 2 * h + 1] = rad * np.sin(ang)
    return z[:, :d]


def sample_mvn(mean, chol, n, seed=0, offset=0):
    z = standard_normals(seed, offset, n, len(mean))
    return np.asarray(mean, dtype=np.float64)[None, :] + z @ np.asarray(chol, dtype=np.float64).T


def mvn_logpdf(X, mean, chol):
    X = np.asarray(X, dtype=np.float64)
    L = np.asarray(chol, dtype=np.float64)
    y = np.linalg.solve(L, (X - np.asarray(mean, dtype=np.float64)[None, :]).T)     # forward substitution
    d = X.shape[1]
    return -0.5 * (y * y).sum(0) - np.log(np.diag(L)).sum() - 0.5 * d * math.log(2.0 * math.pi)


def calc_weights(mean_rec, var_rec, log_prior, ratio):
    """BASQ/_sampler.py:200-216, line by line (numpy).  log_prior = prior.log_prob(pts_rec)."""
    mean_rec, var_rec, log_prior = (np.asarray(a, dtype=np.float64) for a in (mean_rec, var_rec, log_prior))
    with np.errstate(divide="ignore", invalid="ignore"):
        f_rec = np.exp(np.log(np.abs(mean_rec)) + log_prior)
        if ratio < 1:
            g_rec = np.exp(np.log(ratio * var_rec + (1 - ratio) * np.abs(mean_rec)) + log_prior)
        else:
            g_rec = np.exp(np.log(ratio) + np.log(var_rec) + log_prior)
        w = f_rec / g_rec
        return w / w.sum()


def lfi(mu_pred, var_pred, log=False):
    """SOBER/_pi.py:132-139: Normal(0,1).cdf((mu - 1) / sqrt(var)); log adds torch.finfo().eps (fp32)."""
    from scipy.special import ndtr
    v = ndtr((np.asarray(mu_pred, dtype=np.float64) - 1.0) / np.sqrt(np.asarray(var_pred, dtype=np.float64)))
    return np.log(v + np.finfo(np.float32).eps) if log else v


def sir_indices(weights, n_return, seed=0):
    """UncertaintySampler.SIR (BASQ/_sampler.py:104-118) draws torch.multinomial(weights, n) without
    replacement; restated as the exponential race the library runs (include/basq_b200.h:
    basq_sir_resample): key_i = -log(u_i) / w_i, u_i from Philox(seed; counter (i, 0, 0x53495200)),
    n smallest keys in increasing order.  The draws are distributed like multinomial's (Efraimidis &
    Spirakis 2006); the values depend on the generator, as in the reference."""
    w = np.asarray(weights, dtype=np.float64)
    N = len(w)
    i = np.arange(N, dtype=np.uint64)
    ctr = np.zeros((N, 4), dtype=np.uint32)
    ctr[:, 0] = (i & MASK).astype(np.uint32)
    ctr[:, 1] = (i >> np.uint64(32)).astype(np.uint32)
    ctr[:, 3] = 0x53495200
    key = np.empty((N, 2), dtype=np.uint32)
    key[:, 0] = np.uint32(seed & 0xFFFFFFFF)
    key[:, 1] = np.uint32((seed >> 32) & 0xFFFFFFFF)
    r = philox4x32_10(ctr, key)
    u = ((r[:, 0] >> np.uint32(5)).astype(np.float64) * 2.0 ** 26
         + (r[:, 1] >> np.uint32(6)).astype(np.float64) + 0.5) * 2.0 ** -53      # 53-bit uniform in (0, 1)
    with np.errstate(divide="ignore"):
        keys = np.where(w > 0, -np.log(u) / np.where(w > 0, w, 1.0), np.inf)
    order = np.lexsort((np.arange(N), keys))
    n = min(int(n_return), int((w > 0).sum()))
    return order[:n]


def cleansing_weights(weights, eps=float(np.finfo(np.float32).eps)):
    """SOBER/_weights.py:32-38."""
    w = np.array(weights, dtype=np.float64, copy=True)
    w[w < eps] = 0
    w[np.isinf(w)] = eps
    w[np.isnan(w)] = eps
    if not w.sum() == 0:
        w /= w.sum()
    else:
        w = np.ones_like(w) / len(w)
    return w
