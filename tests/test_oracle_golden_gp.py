"""oracle/gp_kernels.py and oracle/sampler.py vs. golden vectors produced by the REFERENCE's own
BASQ/_gp.py, _wsabi.py, _vbq.py, _sampler.py and SOBER/_gp.py, _kernel.py, _pi.py, BASQ/_scale_mmlt.py
(oracle/make_golden_gp.py runs those files on a duck-typed exact-GP model).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import gp_kernels as ok
from oracle import sampler as osamp

TAGS = ["rbf", "ard", "m52", "m32"]
RTOL, ATOL = 1e-10, 1e-12


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


def _close(a, ref):
    np.testing.assert_allclose(np.asarray(a), ref, rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("tag", TAGS)
def test_predict_and_predictive_covariance(golden_dir, tag):
    get, model, x, z = _load(golden_dir, tag)
    m, v = ok.predict(x, model)
    _close(m, get("basq_predict_mean")); _close(v, get("basq_predict_var"))
    _close(m, get("sober_predict_mean"))
    # BASQ/_gp.py:259-277 adds the likelihood noise to the first min(len) diagonal entries; SOBER/_gp.py does not
    _close(ok.predictive_covariance(x, z, model, add_noise_diag=True), get("basq_predcov_xz"))
    _close(ok.predictive_covariance(x, x, model, add_noise_diag=True), get("basq_predcov_xx"))
    _close(ok.predictive_covariance(x, z, model), get("sober_predcov_xz"))
    assert not np.allclose(get("basq_predcov_xz"), get("sober_predcov_xz"), rtol=0, atol=1e-9)  # the quirk is real


@pytest.mark.parametrize("tag", TAGS)
def test_adaptor_objects(golden_dir, tag):
    get, model, x, z = _load(golden_dir, tag)
    vb = ok.VanillaGP(model, add_noise_diag=True)
    _close(vb.predictive_kernel(x, z), get("vbq_kernel_xz"))
    m, v = vb.predict(x)
    _close(m, get("vbq_predict_mean")); _close(v, get("vbq_predict_var"))
    for mode in ("predictive_covariance", "weighted_predictive_covariance", "kernel"):
        _close(ok.Kernel(model, mode)(x, z), get(f"sober_kernel_{mode}"))


@pytest.mark.parametrize("tag", TAGS)
def test_wsabi(golden_dir, tag):
    get, model, x, z = _load(golden_dir, tag)
    ws = ok.WsabiGP(model, alpha=float(get("wsabi_alpha")), jitter=0.0, add_noise_diag=True)
    _close(ws.wsabil_kernel(x, z), get("wsabil_kernel_xz"))
    _close(ws.wsabim_kernel(x, z), get("wsabim_kernel_xz"))
    _close(ws.wsabil_kernel(x, x), get("wsabil_kernel_xx"))
    _close(ws.wsabim_kernel(x, x), get("wsabim_kernel_xx"))
    for nm in ("wsabil_predict", "wsabim_predict"):
        m, v = getattr(ws, nm)(x)
        _close(m, get(f"{nm}_mean")); _close(v, get(f"{nm}_var"))
    _close(ws.wsabil_mean_predict(x), get("wsabil_mean_predict"))
    _close(ws.wsabim_mean_predict(x), get("wsabim_mean_predict"))


@pytest.mark.parametrize("tag", TAGS)
def test_mmlt_and_lfi(golden_dir, tag):
    get, model, x, z = _load(golden_dir, tag)
    mm = ok.ScaleMmltGP(model, jitter=0.0)
    m, v = mm.gspace_predict(x)
    _close(m, get("mmlt_gspace_mean")); _close(v, get("mmlt_gspace_var"))
    _close(mm.gspace_kernel(x, z), get("mmlt_gspace_kernel_xz"))
    _close(mm.gspace_kernel(x, x), get("mmlt_gspace_kernel_xx"))
    _close(mm.hspace_kernel(x, z), get("mmlt_hspace_kernel_xz"))
    _close(osamp.lfi(m.numpy(), v.numpy()), get("lfi"))
    # the reference's log branch raises NameError (SOBER/_pi.py:137 uses torch without importing it):
    # nothing to pin, the oracle keeps the documented intent log(lfi + eps)
    assert int(get("lfi_log_raises")) == 1


@pytest.mark.parametrize("tag", TAGS)
def test_calc_weights(golden_dir, tag):
    get, model, x, z = _load(golden_dir, tag)
    m, v = ok.predict(x, model)
    chol = np.linalg.cholesky(get("prior_cov"))
    logp = osamp.mvn_logpdf(x.numpy(), get("prior_loc"), chol)
    _close(logp, get("prior_logprob"))
    for ratio, key in ((0.5, "calc_weights_r05"), (1.0, "calc_weights_r10")):
        _close(osamp.calc_weights(m.numpy(), v.numpy(), logp, ratio), get(key))
