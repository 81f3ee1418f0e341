"""Drop-in mirror of the reference's ``_rchq.py`` interface, running on libbasq_b200.so.

Same names, argument meaning and return conventions as ``BASQ/_rchq.py`` (and the extended
signature of ``SOBER/_rchq.py``), so a caller switches by changing one import
(``from basq_b200._rchq import recombination``).  Differences, all by design:

* ``kernel`` must be one of the reference's kernel *objects* (see ``basq_b200.kernels``); an
  arbitrary Python closure raises TypeError - there is no Python-callback or CPU fallback;
* the Caratheodory pivots differ from the reference's SVD-based ones, so the selected indices
  differ while the preserved moments and the quadrature estimates agree;
* inputs are never mutated (the SOBER variant mutates ``init_weights`` in place).
"""
from __future__ import annotations

import torch

from . import ops

# BASQ/_rchq.py:53 ignores ``init_weights`` (mu is reset to 1/N); SOBER/_rchq.py:60-61 honours it.
# The BASQ-signature entry point follows BASQ unless this switch is set.
HONOUR_INIT_WEIGHTS_IN_BASQ_SIGNATURE = False


def _as_weights(init_weights, N):
    if init_weights is None:
        return None
    if not torch.is_tensor(init_weights):
        return None  # the reference's default ``init_weights=0``
    if init_weights.ndim == 0:
        return None
    if init_weights.shape[0] != N:
        raise ValueError("init_weights must have one entry per candidate")
    return init_weights


def recombination(pts_rec, pts_nys, num_pts, kernel, device, *args, **kwargs):
    """``recombination(pts_rec, pts_nys, num_pts, kernel, device, init_weights=0)``
    (BASQ/_rchq.py:4-25) or
    ``recombination(pts_rec, pts_nys, num_pts, kernel, device, dtype, init_weights=None, calc_obj=None)``
    (SOBER/_rchq.py:6-31).  Returns ``(idx, w)`` on ``device``: at most ``num_pts`` ascending
    indices into ``pts_rec`` and positive weights with the same total mass."""
    args = list(args)
    dtype = kwargs.pop("dtype", None)
    sober = dtype is not None
    if args and isinstance(args[0], torch.dtype):
        dtype = args.pop(0)
        sober = True
    init_weights = kwargs.pop("init_weights", args.pop(0) if args else (None if sober else 0))
    calc_obj = kwargs.pop("calc_obj", args.pop(0) if args else None)
    if args or kwargs:
        raise TypeError(f"recombination() got unexpected arguments {args} {kwargs}")
    mu = _as_weights(init_weights, len(pts_rec))
    if not sober and not HONOUR_INIT_WEIGHTS_IN_BASQ_SIGNATURE:
        mu = None
    if dtype is not None:
        pts_rec, pts_nys = pts_rec.to(dtype), pts_nys.to(dtype)
    return rc_kernel_svd(pts_rec, pts_nys, num_pts, kernel, device, mu=mu, calc_obj=calc_obj)


def ker_svd_sparsify(pt, s, kernel, device=None):
    """(S, U [s, M]) - BASQ/_rchq.py:28-31.  U has orthonormal rows spanning the randomised range
    of kernel(pt, pt) (torch.svd_lowrank's range finder, niter=2, same RNG consumption); S are the
    Rayleigh quotients of those rows (the reference discards S, :36)."""
    device = torch.device(device) if device is not None else pt.device
    S, U = ops.nystrom_basis(kernel, pt, s, device=device)
    return S.to(pt.dtype), U.to(pt.dtype)


def rc_kernel_svd(samp, pt, s, kernel, device, mu=None, use_obj=True, calc_obj=None):
    """BASQ/_rchq.py:34-40 (SOBER/_rchq.py:33-46 with calc_obj): Nystrom basis, then the
    Tchernychova-Lyons loop.  Returns (idx, w)."""
    device = torch.device(device)
    _, U = ops.nystrom_basis(kernel, pt, s - 1, device=device, want_S=False)
    w_star, idx_star = Mod_Tchernychova_Lyons(samp, U, pt, kernel, device, mu=mu, calc_obj=calc_obj)
    return idx_star, w_star


def Mod_Tchernychova_Lyons(samp, U_svd, pt_nys, kernel, device, mu=None, use_obj=True, DEBUG=False, calc_obj=None):
    """BASQ/_rchq.py:43-130 (SOBER/_rchq.py:48-219 with ``calc_obj``: a callable samp -> [N] whose
    expectation under the rule is pushed up, :67-69).  Returns ``(w_star, idx_star)`` like the reference."""
    device = torch.device(device)
    weights = _as_weights(mu, len(samp))
    obj = None
    if calc_obj is not None:
        obj = -1.0 * torch.as_tensor(calc_obj(samp)).reshape(-1)          # SOBER/_rchq.py:69
    idx, w = ops.recombine(kernel, samp, pt_nys, U_svd, mu=weights, device=device, obj=obj)
    out_dtype = samp.dtype if samp.dtype in (torch.float32, torch.float64) else torch.float32
    return w.to(out_dtype), idx


def Tchernychova_Lyons_CAR(X, mu, device=None, DEBUG=False):
    """BASQ/_rchq.py:133-175: reduce the measure (X [N, n], mu [N]) to at most n + 1 points with the
    same mass and barycentre.  Returns the reference's 7-tuple (w_star, idx_star, nan, nan, 0., nan, nan)."""
    device = torch.device(device) if device is not None else X.device
    Xd = X.detach().to(device=device, dtype=torch.float64)
    mud = mu.detach().to(device=device, dtype=torch.float64)
    # unnormalised barycentre system: row 0 masses, rows 1.. mass-weighted coordinates
    A = torch.cat([mud.unsqueeze(0), (Xd * mud.unsqueeze(1)).T], dim=0).contiguous()
    omega = ops.caratheodory(A)
    keep = omega > 0
    w_star = (mud * omega)[keep].to(mu.dtype)
    idx_star = torch.nonzero(keep).squeeze(1)
    nan = float("nan")
    return w_star, idx_star, nan, nan, 0.0, nan, nan
