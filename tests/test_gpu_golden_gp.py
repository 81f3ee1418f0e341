"""The CUDA path vs. outputs of the REFERENCE's own GP-side files (tests/golden/gp_kernels.npz, made by
oracle/make_golden_gp.py from BASQ/_gp.py, _wsabi.py, _vbq.py, _sampler.py, SOBER/_gp.py, _kernel.py,
_pi.py, SOBER/BASQ/_scale_mmlt.py).  Everything goes through the C ABI (basq_gram, basq_gp_predict,
basq_candidate_weights); the oracle classes appear only as the duck-typed kernel objects a reference
caller would hand to the library (kernels.py reads their attributes, it never calls them)."""
import os

import numpy as np
import pytest
import torch

from oracle import gp_kernels as ok

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")
TAGS = ["rbf", "ard", "m52", "m32"]


@pytest.fixture(scope="module")
def lib():
    from basq_b200 import ops, sampler
    return ops, sampler


def _load(golden_dir, tag):
    g = np.load(os.path.join(golden_dir, "gp_kernels.npz"))
    get = lambda k: g[f"{tag}_{k}"]
    fam = int(get("family"))
    ls = get("lengthscale")
    base = ok.RBFKernel(ls) if fam == 0 else ok.MaternKernel(ls, nu=1.5 if fam == 1 else 2.5)
    model = ok.ExactGP(torch.from_numpy(get("Xobs")), torch.from_numpy(get("yobs")),
                       ok.ScaleKernel(base, float(get("outputscale"))), noise=float(get("noise")),
                       mean_const=float(get("mean_const")))
    return get, model, torch.from_numpy(get("x")), torch.from_numpy(get("z"))


def _rel(a, ref):
    a = np.asarray(a.detach().double().cpu() if torch.is_tensor(a) else a, dtype=np.float64)
    return float(np.abs(a - ref).max() / max(np.abs(ref).max(), 1e-300))


# fp64 inputs (SOBER's global dtype): 1e-9; fp32 inputs (BASQ's default): kernel-value rounding times the
# conditioning of (K + noise I)^-1 of these small GPs
TOLS = [(torch.float64, 1e-9), (torch.float32, 5e-4)]


@pytest.mark.parametrize("tag", TAGS)
@pytest.mark.parametrize("dtype,tol", TOLS)
def test_gram_modes_match_reference(lib, golden_dir, tag, dtype, tol):
    ops, _ = lib
    get, model, x, z = _load(golden_dir, tag)
    alpha = float(get("wsabi_alpha"))
    ws = ok.WsabiGP(model, alpha=alpha, jitter=0.0, add_noise_diag=True)
    mm = ok.ScaleMmltGP(model, jitter=0.0)
    cases = [
        (ok.VanillaGP(model, add_noise_diag=True).predictive_kernel, "vbq_kernel_xz", z),     # BASQ/_vbq.py:119-128
        (ok.VanillaGP(model, add_noise_diag=True).predictive_kernel, "basq_predcov_xx", x),   # square: full noise diagonal
        (ws.wsabil_kernel, "wsabil_kernel_xz", z), (ws.wsabim_kernel, "wsabim_kernel_xz", z), # BASQ/_wsabi.py:205-249
        (ws.wsabil_kernel, "wsabil_kernel_xx", x), (ws.wsabim_kernel, "wsabim_kernel_xx", x),
        (mm.gspace_kernel, "mmlt_gspace_kernel_xz", z), (mm.gspace_kernel, "mmlt_gspace_kernel_xx", x),
        (ok.Kernel(model, "predictive_covariance"), "sober_kernel_predictive_covariance", z),  # SOBER/_kernel.py
        (ok.Kernel(model, "weighted_predictive_covariance"), "sober_kernel_weighted_predictive_covariance", z),
        (ok.Kernel(model, "kernel"), "sober_kernel_kernel", z),
    ]
    xd = x.to(DEV, dtype)
    for kern, key, other in cases:
        K = ops.gram(kern, xd, other.to(DEV, dtype))
        assert _rel(K, get(key)) < tol, (key, _rel(K, get(key)))


@pytest.mark.parametrize("tag", TAGS)
@pytest.mark.parametrize("dtype,tol", TOLS)
def test_predict_matches_reference(lib, golden_dir, tag, dtype, tol):
    ops, _ = lib
    get, model, x, _ = _load(golden_dir, tag)
    xd = x.to(DEV, dtype)
    mean, var = ops.gp_predict(ok.VanillaGP(model).predictive_kernel, xd)                     # BASQ/_gp.py:213-230
    assert _rel(mean, get("basq_predict_mean")) < tol and _rel(var, get("basq_predict_var")) < tol
    ws = ok.WsabiGP(model, alpha=float(get("wsabi_alpha")))
    for kern, nm in ((ws.wsabil_kernel, "wsabil_predict"), (ws.wsabim_kernel, "wsabim_predict")):
        mean, var = ops.gp_predict(kern, xd, space=1)                                         # BASQ/_wsabi.py:251-277
        assert _rel(mean, get(f"{nm}_mean")) < tol and _rel(var, get(f"{nm}_var")) < tol, nm
    mean, var = ops.gp_predict(ok.ScaleMmltGP(model).gspace_kernel, xd, space=1)              # _scale_mmlt.py:211-223
    assert _rel(mean, get("mmlt_gspace_mean")) < tol and _rel(var, get("mmlt_gspace_var")) < tol


@pytest.mark.parametrize("tag", TAGS)
def test_candidate_weights_match_reference(lib, golden_dir, tag):
    ops, sampler = lib
    get, model, x, _ = _load(golden_dir, tag)
    xd = x.to(DEV)
    kern = ok.VanillaGP(model).predictive_kernel
    for ratio, key in ((0.5, "calc_weights_r05"), (1.0, "calc_weights_r10")):                 # BASQ/_sampler.py:200-216
        w = sampler.calc_weights(kern, xd, ratio=ratio)
        assert _rel(w, get(key)) < 1e-9, key
    out = sampler.lfi(ok.ScaleMmltGP(model).gspace_kernel, xd, log=False)                     # SOBER/_pi.py:121-139
    assert float(np.abs(out.cpu().numpy() - get("lfi")).max()) < 1e-10
