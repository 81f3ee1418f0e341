"""Oracle restatement (torch CPU) of the kernel callables that feed recombination.

TEST INFRASTRUCTURE - see ``oracle/__init__.py``.  PINNED at the reference's call
sites: ``oracle/make_golden_gp.py`` runs the reference's own ``predictive_covariance``,
``WsabiGP``, ``VanillaGP``, ``ScaleMmltGP`` and ``Kernel`` code on a duck-typed
exact-GP model and ``tests/test_oracle_golden_gp.py`` checks this module against those
outputs.  The layer underneath belongs to gpytorch (absent here, no version pinned by
the reference) and is restated from its published definition: base kernels follow
``RBFKernel`` / ``MaternKernel`` / ``ScaleKernel.forward``; GP caches are exact fp64
Cholesky (the reference's LOVE variance is an approximation of the same quantity,
``BASQ/_gp.py:228``).  Each function cites the reference call site it follows.

The stand-in classes expose the gpytorch *attribute surface* the reference
introspects (``train_inputs``, ``likelihood.noise``, ``covar_module.outputscale``,
``covar_module.base_kernel.lengthscale``, ``prediction_strategy.{mean,covar}_cache``)
so the product's duck-typing shim can be exercised without gpytorch.
"""
from __future__ import annotations

import math
import types

import torch


# --------------------------------------------------------------------------- base kernels
def _sq_dist(x, y):
    """Pairwise squared Euclidean distance, direct differences (no cancellation)."""
    diff = x.unsqueeze(1) - y.unsqueeze(0)
    return (diff * diff).sum(-1)


def _sq_dist_matmul(x, y):
    """gpytorch's own distance (``gpytorch.kernels.kernel.sq_dist``): centre both operands by the mean of
    ``x``, then |x|^2 + |y|^2 - 2 x y^T as ONE matmul, clamped at zero.  O(a b) memory instead of the
    O(a b d) of the direct differences; used by the timed baselines (bench.py), where it is also what the
    reference actually executes."""
    adj = x.mean(-2, keepdim=True)
    x, y = x - adj, y - adj
    xn = (x * x).sum(-1, keepdim=True)
    yn = (y * y).sum(-1, keepdim=True)
    x_ = torch.cat([-2.0 * x, xn, torch.ones_like(xn)], dim=-1)
    y_ = torch.cat([y, torch.ones_like(yn), yn], dim=-1)
    return (x_ @ y_.T).clamp_min_(0.0)


def base_kernel(x, y, family, lengthscale, outputscale, nu=2.5, matmul_dist=False):
    """gpytorch ``ScaleKernel(RBFKernel|MaternKernel).forward``.

    RBF:     s * exp(-0.5 * |x-y|^2 / l^2)
    Matern:  s * poly_nu(sqrt(2nu) r) * exp(-sqrt(2nu) r),  r = |x-y| / l
             nu=1.5: 1 + sqrt3 r ; nu=2.5: 1 + sqrt5 r + 5/3 r^2
    ``lengthscale`` is a scalar or a length-d vector (ARD).
    """
    ls = torch.as_tensor(lengthscale, dtype=x.dtype, device=x.device).reshape(1, -1)
    r2 = _sq_dist_matmul(x / ls, y / ls) if matmul_dist else _sq_dist(x / ls, y / ls)
    if family == "rbf":
        return outputscale * torch.exp(-0.5 * r2)
    if family == "matern":
        r = torch.sqrt(r2.clamp_min(0.0))
        c = math.sqrt(2.0 * nu)
        e = torch.exp(-c * r)
        if abs(nu - 1.5) < 1e-12:
            poly = 1.0 + c * r
        elif abs(nu - 2.5) < 1e-12:
            poly = 1.0 + c * r + (5.0 / 3.0) * r2
        elif abs(nu - 0.5) < 1e-12:
            poly = 1.0
        else:
            raise ValueError("nu must be 0.5, 1.5 or 2.5")
        return outputscale * poly * e
    raise ValueError(family)


class RBFKernel:
    """Attribute stand-in for ``gpytorch.kernels.RBFKernel``."""

    def __init__(self, lengthscale):
        self.lengthscale = torch.as_tensor(lengthscale, dtype=torch.float64).reshape(1, -1)


class MaternKernel:
    """Attribute stand-in for ``gpytorch.kernels.MaternKernel``."""

    def __init__(self, lengthscale, nu=2.5):
        self.lengthscale = torch.as_tensor(lengthscale, dtype=torch.float64).reshape(1, -1)
        self.nu = nu


class ScaleKernel:
    """Attribute stand-in for ``gpytorch.kernels.ScaleKernel`` (``BASQ/_parameters.py:200-205``)."""

    def __init__(self, base, outputscale=1.0, matmul_dist=False):
        self.base_kernel = base
        self.outputscale = torch.as_tensor(float(outputscale), dtype=torch.float64)
        self.matmul_dist = matmul_dist      # gpytorch's matmul-based distance (timed baselines)

    def forward(self, x, y):
        fam = "rbf" if isinstance(self.base_kernel, RBFKernel) else "matern"
        return base_kernel(
            x, y, fam, self.base_kernel.lengthscale.to(x).reshape(-1),
            float(self.outputscale), nu=getattr(self.base_kernel, "nu", 2.5), matmul_dist=self.matmul_dist,
        )

    __call__ = forward


class ExactGP:
    """Attribute stand-in for the reference's ``ExactGPModel`` with FIXED hyper-parameters
    (``BASQ/_gp.py:9-30`` ConstantMean, ``SOBER/_gp.py:18`` ZeroMean -> ``mean_const=0``)."""

    def __init__(self, train_x, train_y, covar_module, noise=1e-10, mean_const=0.0):
        train_x = train_x.to(torch.float64)
        train_y = train_y.to(torch.float64)
        self.train_inputs = (train_x,)
        self.train_targets = train_y
        self.covar_module = covar_module
        self.likelihood = types.SimpleNamespace(noise=torch.tensor([float(noise)], dtype=torch.float64))
        self.mean_module = types.SimpleNamespace(constant=torch.tensor(float(mean_const), dtype=torch.float64))
        K = covar_module.forward(train_x, train_x)
        K = K + float(noise) * torch.eye(len(train_x), dtype=torch.float64)
        L = torch.linalg.cholesky(K)
        Linv = torch.linalg.solve_triangular(L, torch.eye(len(train_x), dtype=torch.float64), upper=False)
        # gpytorch: covar_cache S with S S^T = (K + s2 I)^-1 ; mean_cache = (K + s2 I)^-1 (y - m)
        self.prediction_strategy = types.SimpleNamespace(
            covar_cache=Linv.T.contiguous(),
            mean_cache=torch.cholesky_solve((train_y - mean_const).unsqueeze(1), L).squeeze(1),
        )

    def eval(self):
        return self

    def to(self, device):
        """Move the observations and caches (used by bench.py's torch-on-GPU baseline leg)."""
        self.train_inputs = (self.train_inputs[0].to(device),)
        self.train_targets = self.train_targets.to(device)
        ps = self.prediction_strategy
        ps.covar_cache, ps.mean_cache = ps.covar_cache.to(device), ps.mean_cache.to(device)
        return self


# --------------------------------------------------------------------------- GP prediction
def get_cov_cache(model):
    """``BASQ/_gp.py:233-256``: W = S S^T = (K_XX + s2 I)^-1, Xobs, noise."""
    S = model.prediction_strategy.covar_cache
    return S @ S.T, model.train_inputs[0], model.likelihood.noise


def predict(x, model):
    """Exact GP posterior mean / variance *including* likelihood noise
    (``BASQ/_gp.py:213-230``, ``SOBER/_gp.py:212-237``; exact instead of LOVE)."""
    W, Xobs, noise = get_cov_cache(model)
    x64 = x.to(torch.float64)
    KxX = model.covar_module.forward(x64, Xobs)
    mean = float(model.mean_module.constant) + KxX @ model.prediction_strategy.mean_cache
    kxx = float(model.covar_module.outputscale)
    var = kxx - ((KxX @ W) * KxX).sum(-1) + float(noise)
    return mean.to(x.dtype), var.to(x.dtype)


def predictive_covariance(x, y, model, add_noise_diag=False):
    """``BASQ/_gp.py:259-277`` (``add_noise_diag=True`` reproduces its ``+ lik_var`` on the
    first min(len) diagonal entries) / ``SOBER/_gp.py:281-305`` (no diagonal term)."""
    W, Xobs, noise = get_cov_cache(model)
    dt = x.dtype
    x64, y64 = x.to(torch.float64), y.to(torch.float64)
    Kxy = model.covar_module.forward(x64, y64)
    KxX = model.covar_module.forward(x64, Xobs)
    KXy = model.covar_module.forward(Xobs, y64)
    cov = Kxy - KxX @ W @ KXy
    if add_noise_diag:
        k = min(len(x), len(y))
        ii = torch.arange(k, device=cov.device)
        cov[ii, ii] = cov[ii, ii] + float(noise)
    return cov.to(dt)


# --------------------------------------------------------------------------- adaptor objects
class VanillaGP:
    """``BASQ/_vbq.py:119-151`` - the kernel object of the reference's default config."""

    def __init__(self, model, add_noise_diag=False):
        self.model = model
        self._diag = add_noise_diag

    def predictive_kernel(self, x, y):
        return predictive_covariance(x, y, self.model, self._diag)

    def predict(self, x):
        return predict(x, self.model)

    def predict_mean(self, x):
        return predict(x, self.model)[0]


class WsabiGP:
    """``BASQ/_wsabi.py:194-301`` kernels / predictors of the square-root warped GP."""

    def __init__(self, model, alpha=0.0, jitter=0.0, add_noise_diag=False):
        self.model = model
        self.alpha = alpha
        self.jitter = jitter
        self._diag = add_noise_diag

    def predictive_kernel(self, x, y):
        return predictive_covariance(x, y, self.model, self._diag)

    def _with_jitter(self, C, x, y):
        k = min(len(x), len(y))
        ii = torch.arange(k)
        C[ii, ii] = C[ii, ii] + self.jitter
        return C

    def wsabil_kernel(self, x, y):  # :205-226
        mx, _ = predict(x, self.model)
        my, _ = predict(y, self.model)
        C = predictive_covariance(x, y, self.model, self._diag)
        return self._with_jitter(mx.unsqueeze(1) * C * my.unsqueeze(0), x, y)

    def wsabim_kernel(self, x, y):  # :228-249
        mx, _ = predict(x, self.model)
        my, _ = predict(y, self.model)
        C = predictive_covariance(x, y, self.model, self._diag)
        return self._with_jitter(mx.unsqueeze(1) * C * my.unsqueeze(0) + 0.5 * C * C, x, y)

    def wsabil_predict(self, x):  # :251-263
        m, v = predict(x, self.model)
        return self.alpha + 0.5 * m * m, m * v * m

    def wsabim_predict(self, x):  # :265-277
        m, v = predict(x, self.model)
        return self.alpha + 0.5 * (m * m + v), m * v * m + 0.5 * v * v

    def wsabil_mean_predict(self, x):
        return self.wsabil_predict(x)[0]

    def wsabim_mean_predict(self, x):
        return self.wsabim_predict(x)[0]


class ScaleMmltGP:
    """``SOBER/BASQ/_scale_mmlt.py:199-278`` moment-matched log transform (MMLT)."""

    def __init__(self, model, jitter=0.0):
        self.model = model
        self.jitter = jitter

    def hspace_predict(self, x):
        return predict(x, self.model)

    def gspace_predict(self, x):  # :211-223
        mh, vh = predict(x, self.model)
        mg = torch.exp(mh + 0.5 * vh) - 1.0
        return mg, mg * mg * (torch.exp(vh) - 1.0)

    def gspace_mean_predict(self, x):
        return self.gspace_predict(x)[0]

    def hspace_kernel(self, x, y):
        return predictive_covariance(x, y, self.model, False)

    def gspace_kernel(self, x, y):  # :258-278
        gx = self.gspace_mean_predict(x)
        gy = self.gspace_mean_predict(y)
        C = self.hspace_kernel(x, y)
        out = gx.unsqueeze(1) * gy.unsqueeze(0) * (torch.exp(C) - 1.0)
        k = min(len(x), len(y))
        ii = torch.arange(k)
        out[ii, ii] = out[ii, ii] + self.jitter
        return out


class Kernel:
    """``SOBER/_kernel.py:4-47`` kernel adaptor (three modes)."""

    def __init__(self, model, mode="predictive_covariance"):
        self.model = model
        self.mode = mode

    def __call__(self, x, y):
        if self.mode == "predictive_covariance":
            return predictive_covariance(x, y, self.model, False)
        if self.mode == "weighted_predictive_covariance":
            mx, _ = predict(x, self.model)
            my, _ = predict(y, self.model)
            return mx.unsqueeze(1) * predictive_covariance(x, y, self.model, False) * my.unsqueeze(0)
        if self.mode == "kernel":
            return self.model.covar_module.forward(x, y)
        raise ValueError(
            'mode should be from ["predictive_covariance", "weighted_predictive_covariance", "kernel"]')


# --------------------------------------------------------------------------- synthetic workload
def make_gp(d, n_obs, family="rbf", lengthscale=1.5, outputscale=1.0, noise=1e-10, nu=2.5,
            mean_const=0.0, seed=0, log_targets=False):
    """Fixed-hyper-parameter GP on synthetic observations: prior N(0, 2 I_d),
    targets from a smooth positive test likelihood (a 3-component Gaussian mixture)."""
    g = torch.Generator().manual_seed(seed)
    X = math.sqrt(2.0) * torch.randn(n_obs, d, generator=g, dtype=torch.float64)
    centres = 1.5 * torch.randn(3, d, generator=g, dtype=torch.float64)
    y = torch.zeros(n_obs, dtype=torch.float64)
    for c in centres:
        y = y + torch.exp(-0.25 * ((X - c) ** 2).sum(-1)) / 3.0
    if log_targets:
        y = torch.log(y + 1.0)
    base = RBFKernel(lengthscale) if family == "rbf" else MaternKernel(lengthscale, nu)
    return ExactGP(X, y, ScaleKernel(base, outputscale), noise=noise, mean_const=mean_const)
