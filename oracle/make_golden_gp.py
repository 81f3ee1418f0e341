"""Golden vectors for the GP-side arithmetic of the path, produced by the REFERENCE's own files.

TEST INFRASTRUCTURE - run in the build container only:  ``python oracle/make_golden_gp.py``
(/root/reference does not exist on the GPU box; the committed tests/golden/gp_kernels.npz travels).

What runs here is the reference's code, unmodified, imported from the read-only tree:

  BASQ/_gp.py            predict, get_cov_cache, predictive_covariance          (:213-289)
  BASQ/_wsabi.py         WsabiGP.wsabil_kernel / wsabim_kernel / *_predict      (:205-277)
  BASQ/_vbq.py           VanillaGP.predictive_kernel / predict                  (:119-152)
  BASQ/_sampler.py       UncertaintySampler.calc_weights                        (:200-216)
  SOBER/_gp.py           predict, predict_mean, predictive_covariance           (:211-310)
  SOBER/_kernel.py       Kernel.__call__ (three modes)                          (:16-46)
  SOBER/BASQ/_scale_mmlt.py  ScaleMmltGP.gspace_predict / gspace_kernel         (:211-278)
  SOBER/_pi.py           PI_BQ.lfi                                              (:121-139)

Those files import gpytorch / botorch / matplotlib, none of which is installed here (and the reference
pins no version).  They only need the modules at import time plus a fitted *model object* at call time,
so this script (a) registers empty stand-in modules that satisfy the import statements and
(b) hands the reference a duck-typed model exposing exactly the attributes its functions touch:
``train_inputs``, ``train_targets``, ``likelihood`` (``noise``, ``eval()``, ``__call__``),
``covar_module.forward`` (+ ``outputscale``, ``base_kernel.lengthscale``),
``prediction_strategy.covar_cache``, ``eval()``, ``__call__``.  The model is the textbook exact GP
(gpytorch's published ScaleKernel(RBF/Matern).forward, exact posterior mean / variance, Gaussian
likelihood adding the noise to the marginal variance, covar_cache S with S S^T = (K + noise I)^-1),
written here independently of oracle/gp_kernels.py.  The stand-in therefore pins everything the
reference ITSELF computes on top of those primitives - the Woodbury covariance and its noise-diagonal
quirk, the WSABI-L/M and MMLT warps, the weighted kernel, the importance weights, LFI - and leaves
only gpytorch's own primitives as "published algorithm restated" (fast_pred_var's LOVE approximation
is a no-op context here: exact variance, as in the library).

All constructors of the reference classes fit a GP with gpytorch's optimisers; they are bypassed with
``object.__new__`` and the few attributes the kernel / predict methods read are set by hand.
"""
from __future__ import annotations

import contextlib
import math
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("BASQ_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


# ------------------------------------------------------------------ stand-in modules (import time only)
def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install_import_stubs():
    class _Anything:
        def __init__(self, *a, **k):
            pass

    @contextlib.contextmanager
    def _ctx(*a, **k):
        yield

    g = _module("gpytorch")
    g.models = _module("gpytorch.models", ExactGP=_Anything)
    g.means = _module("gpytorch.means", ConstantMean=_Anything, ZeroMean=_Anything)
    g.distributions = _module("gpytorch.distributions", MultivariateNormal=_Anything)
    g.likelihoods = _module("gpytorch.likelihoods", GaussianLikelihood=_Anything)
    g.constraints = _module("gpytorch.constraints", Interval=_Anything)
    g.mlls = _module("gpytorch.mlls", ExactMarginalLogLikelihood=_Anything)
    g.settings = _module("gpytorch.settings", fast_pred_var=_ctx, fast_computations=_ctx, cholesky_jitter=_ctx)
    g.priors = _module("gpytorch.priors")
    g.priors.torch_priors = _module("gpytorch.priors.torch_priors", GammaPrior=_Anything)
    b = _module("botorch")
    b.fit = _module("botorch.fit", fit_gpytorch_model=lambda *a, **k: None)
    mp = _module("matplotlib")
    mp.pyplot = _module("matplotlib.pyplot")


# ------------------------------------------------------------------ the fitted-model stand-in (call time)
def _scaled_dist(x, y, ls):
    d = x.unsqueeze(1) / ls - y.unsqueeze(0) / ls
    return (d * d).sum(-1)


class _Base:
    def __init__(self, family, lengthscale):
        self.family = family
        self.lengthscale = torch.as_tensor(lengthscale, dtype=torch.float64).reshape(1, -1)

    def __call__(self, x, y):
        r2 = _scaled_dist(x, y, self.lengthscale)
        if self.family == "rbf":
            return torch.exp(-0.5 * r2)
        r = torch.sqrt(r2.clamp_min(0))
        if self.family == "matern15":
            c = math.sqrt(3.0)
            return (1 + c * r) * torch.exp(-c * r)
        c = math.sqrt(5.0)
        return (1 + c * r + 5.0 / 3.0 * r2) * torch.exp(-c * r)


class _Scale:
    def __init__(self, base, outputscale):
        self.base_kernel = base
        self.outputscale = torch.tensor(float(outputscale), dtype=torch.float64)

    def forward(self, x, y):
        return self.outputscale * self.base_kernel(x, y)


class _Dist:
    def __init__(self, mean, variance):
        self.mean, self.variance = mean, variance


class _Likelihood:
    def __init__(self, noise):
        self.noise = torch.tensor([float(noise)], dtype=torch.float64)

    def eval(self):
        return self

    def __call__(self, dist):
        return _Dist(dist.mean, dist.variance + self.noise)


class _Strategy:
    pass


class StandInGP:
    def __init__(self, X, y, covar, noise, mean_const):
        self.train_inputs = (X,)
        self.train_targets = y
        self.covar_module = covar
        self.likelihood = _Likelihood(noise)
        self.mean_const = float(mean_const)
        K = covar.forward(X, X) + float(noise) * torch.eye(len(X), dtype=torch.float64)
        L = torch.linalg.cholesky(K)
        Linv = torch.linalg.solve_triangular(L, torch.eye(len(X), dtype=torch.float64), upper=False)
        self.prediction_strategy = _Strategy()
        self.prediction_strategy.covar_cache = Linv.T.contiguous()      # S S^T = (K + noise I)^-1
        self._alpha = Linv.T @ (Linv @ (y - self.mean_const))
        self._Linv = Linv

    def eval(self):
        return self

    def __call__(self, x):
        k = self.covar_module.forward(x, self.train_inputs[0])
        mean = self.mean_const + k @ self._alpha
        v = self._Linv @ k.T
        prior = self.covar_module.outputscale * torch.ones(len(x), dtype=torch.float64)
        return _Dist(mean, prior - (v * v).sum(0))


# ------------------------------------------------------------------ cases
def make_case(seed, d, n_obs, family, lengthscale, outputscale, noise, mean_const, nx, ny):
    g = torch.Generator().manual_seed(seed)
    X = torch.randn(n_obs, d, generator=g, dtype=torch.float64)
    y = torch.sin(X.sum(-1)) + 0.3 * torch.randn(n_obs, generator=g, dtype=torch.float64) + 1.5
    x = torch.randn(nx, d, generator=g, dtype=torch.float64)
    z = torch.randn(ny, d, generator=g, dtype=torch.float64)
    model = StandInGP(X, y, _Scale(_Base(family, lengthscale), outputscale), noise, mean_const)
    meta = dict(Xobs=X, yobs=y, x=x, z=z, lengthscale=torch.as_tensor(lengthscale, dtype=torch.float64).reshape(-1),
                outputscale=torch.tensor(float(outputscale)), noise=torch.tensor(float(noise)),
                mean_const=torch.tensor(float(mean_const)))
    return model, meta


CASES = {
    # tag: (seed, d, n_obs, family, lengthscale, outputscale, noise, mean_const, nx, ny)
    "rbf": (11, 3, 24, "rbf", 1.3, 1.7, 1e-4, 0.4, 17, 13),
    "ard": (12, 4, 30, "rbf", [0.9, 1.4, 2.0, 1.1], 0.8, 1e-3, 0.0, 15, 15),
    "m52": (13, 2, 20, "matern25", 1.1, 1.2, 1e-3, -0.2, 9, 21),
    "m32": (14, 3, 20, "matern15", 1.6, 2.0, 1e-2, 0.0, 12, 12),
}
FAMILY_CODE = {"rbf": 0, "matern15": 1, "matern25": 2}


def main():
    sys.path.insert(0, REF)
    torch.set_default_dtype(torch.float64)
    install_import_stubs()
    from BASQ import _gp as bgp, _wsabi as bwsabi, _vbq as bvbq, _sampler as bsampler
    from SOBER import _gp as sgp, _kernel as skernel, _pi as spi
    from SOBER.BASQ import _scale_mmlt as smmlt
    from torch.distributions.multivariate_normal import MultivariateNormal

    out = {}
    for tag, spec in CASES.items():
        model, meta = make_case(*spec)
        x, z = meta["x"], meta["z"]
        for k, v in meta.items():
            out[f"{tag}_{k}"] = v.numpy()
        out[f"{tag}_family"] = np.int64(FAMILY_CODE[spec[3]])

        def put(name, t):
            out[f"{tag}_{name}"] = t.detach().numpy().copy()

        # BASQ/_gp.py
        m, v = bgp.predict(x, model)
        put("basq_predict_mean", m); put("basq_predict_var", v)
        put("basq_predcov_xz", bgp.predictive_covariance(x, z, model))        # adds lik_var on min(nx, ny) diagonal
        put("basq_predcov_xx", bgp.predictive_covariance(x, x, model))
        # SOBER/_gp.py (no diagonal term)
        put("sober_predcov_xz", sgp.predictive_covariance(x, z, model))
        put("sober_predict_mean", sgp.predict_mean(x, model))
        # SOBER/_kernel.py
        for mode in ("predictive_covariance", "weighted_predictive_covariance", "kernel"):
            put(f"sober_kernel_{mode}", skernel.Kernel(model, mode=mode)(x, z))
        # BASQ/_vbq.py
        vb = object.__new__(bvbq.VanillaGP)
        vb.model = model
        put("vbq_kernel_xz", vb.predictive_kernel(x, z))
        mu, var = vb.predict(x)
        put("vbq_predict_mean", mu); put("vbq_predict_var", var)
        # BASQ/_wsabi.py
        ws = object.__new__(bwsabi.WsabiGP)
        ws.model = model
        ws.alpha = torch.tensor(0.37)
        ws.jitter = 0
        out[f"{tag}_wsabi_alpha"] = np.float64(0.37)
        put("wsabil_kernel_xz", ws.wsabil_kernel(x, z))
        put("wsabim_kernel_xz", ws.wsabim_kernel(x, z))
        put("wsabil_kernel_xx", ws.wsabil_kernel(x, x))
        put("wsabim_kernel_xx", ws.wsabim_kernel(x, x))
        for nm in ("wsabil_predict", "wsabim_predict"):
            mu, var = getattr(ws, nm)(x)
            put(f"{nm}_mean", mu); put(f"{nm}_var", var)
        put("wsabil_mean_predict", ws.wsabil_mean_predict(x))
        put("wsabim_mean_predict", ws.wsabim_mean_predict(x))
        # SOBER/BASQ/_scale_mmlt.py
        mm = object.__new__(smmlt.ScaleMmltGP)
        mm.model = model
        mm.jitter = torch.tensor(0.0)
        mu, var = mm.gspace_predict(x)
        put("mmlt_gspace_mean", mu); put("mmlt_gspace_var", var)
        put("mmlt_gspace_kernel_xz", mm.gspace_kernel(x, z))
        put("mmlt_gspace_kernel_xx", mm.gspace_kernel(x, x))
        put("mmlt_hspace_kernel_xz", mm.hspace_kernel(x, z))
        # SOBER/_pi.py
        pi = spi.PI_BQ(mm)
        put("lfi", pi.lfi(x, log=False))
        try:
            put("lfi_log", pi.lfi(x, log=True))
        except NameError:
            # SOBER/_pi.py never imports torch, so its log branch (`torch.finfo().eps`, :137) raises in the
            # reference itself; there is no reference output to pin for log=True.
            out[f"{tag}_lfi_log_raises"] = np.int64(1)
        # BASQ/_sampler.py: calc_weights needs self.model, self.prior, self.ratio only
        d = x.shape[1]
        g = torch.Generator().manual_seed(100 + spec[0])
        A = torch.randn(d, d, generator=g, dtype=torch.float64)
        cov = A @ A.T / d + 0.5 * torch.eye(d, dtype=torch.float64)
        loc = 0.1 * torch.randn(d, generator=g, dtype=torch.float64)
        prior = MultivariateNormal(loc, cov)
        put("prior_loc", loc); put("prior_cov", cov)
        put("prior_logprob", prior.log_prob(x))
        for ratio in (0.5, 1.0):
            us = object.__new__(bsampler.UncertaintySampler)
            us.model, us.prior, us.ratio = model, prior, ratio
            put(f"calc_weights_r{int(ratio * 10):02d}", us.calc_weights(x))

    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, "gp_kernels.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path)} bytes")


if __name__ == "__main__":
    main()
