"""Oracle restatement (plain torch ops; CPU, or the inputs' device) of the reference's kernel
recombination, ``BASQ/_rchq.py``.

TEST INFRASTRUCTURE - see ``oracle/__init__.py``.  PINNED: checked against outputs of the
reference's own ``BASQ/_rchq.py`` (imported from the read-only reference tree by
``oracle/make_golden.py``) stored under ``tests/golden/``.

Notation: N candidates, M Nystrom landmarks, num_pts = q + 1 returned points, q test
functions phi_i(x) = U_i . k(Z, x), S = 2 (q + 1) sets per round.
"""
from __future__ import annotations

import torch


def caratheodory(X, mu):
    """``Tchernychova_Lyons_CAR`` (``BASQ/_rchq.py:133-175``).

    X [S, q] points, mu [S] positive weights -> (w [<= q+1] > 0, idx ascending) with the same
    mass and barycentre.  Null space of [1|X]^T from a full SVD (:140-143), then S-(q+1)
    ratio-test eliminations (:146-171).
    """
    S = X.shape[0]
    A = torch.cat([torch.ones(S, 1, dtype=X.dtype, device=X.device), X], dim=1)
    n = A.shape[1]
    _, _, Vh = torch.linalg.svd(A.T)
    Phi = Vh[n:, :].T.contiguous()
    mu = mu.clone()
    inf = torch.tensor(float("inf"), dtype=X.dtype, device=X.device)
    for _ in range(S - n):
        phi = Phi[:, 0]
        pos = phi > 0
        if not bool(pos.any()):
            # the reference's argmin over an empty selection (:152)
            raise RuntimeError("caratheodory: null vector has no positive entry")
        ratio = torch.where(pos, mu / phi, inf)
        k = int(torch.argmin(ratio))
        mu = mu - ratio[k] * phi
        mu[k] = 0.0
        rest = Phi[:, 1:]
        Phi = rest - torch.outer(phi, rest[k]) / phi[k]
        Phi[k, :] = 0.0
    keep = mu > 0
    return mu[keep], torch.nonzero(keep).squeeze(1)


def tchernychova_lyons(samp, U, pt_nys, kernel, mu=None, chunk=1):
    """``Mod_Tchernychova_Lyons`` (``BASQ/_rchq.py:43-130``).

    ``mu=None`` reproduces the BASQ variant (uniform 1/N, ``init_weights`` ignored, :53);
    a tensor reproduces the SOBER variant's honoured weights (``SOBER/_rchq.py:60-61``),
    without its tail double-count.  ``chunk`` = number of S-point strips per
    kernel call (1 = the reference's exact call sequence, :81-86; larger = same sums,
    fewer Python iterations, used by the timed CPU baseline).
    Returns (w_star, idx_star).
    """
    N = len(samp)
    q, M = U.shape
    S = 2 * (q + 1)
    dt, dev = U.dtype, samp.device
    mu = torch.full((N,), 1.0 / N, dtype=dt, device=dev) if mu is None else mu.to(dt).clone()
    alive = torch.arange(N, device=dev)[mu != 0]

    while True:
        R = len(alive)
        if R <= q + 1:                                              # :60-63
            idx = torch.nonzero(mu > 0).squeeze(1)
            return mu[idx], idx
        if R <= S:                                                  # :65-74 final exact stage
            feats = U @ kernel(pt_nys, samp[alive])
            w, keep = caratheodory(feats.T, mu[alive])
            alive = alive[keep]
            mu = torch.zeros_like(mu)
            mu[alive] = w
            return mu[mu > 0], alive

        E = R // S                                                  # :76-78
        body = alive[: E * S].reshape(E, S)
        tail = alive[E * S:]
        G = torch.zeros(M, S, dtype=dt, device=dev)
        for e0 in range(0, E, chunk):                               # :81-86
            e1 = min(E, e0 + chunk)
            ids = body[e0:e1].reshape(-1)
            Kb = kernel(pt_nys, samp[ids]) * mu[ids].unsqueeze(0)
            G += Kb.reshape(M, e1 - e0, S).sum(1) if e1 - e0 > 1 else Kb
        bary = (U @ G).T                                            # :88-89
        mass = mu[body].sum(0)                                      # :90
        if len(tail):                                               # :93-99 tail -> last set
            bary[-1] += (U @ kernel(pt_nys, samp[tail])) @ mu[tail]
            mass[-1] += mu[tail].sum()
        bary = bary / mass.unsqueeze(1)                             # :101

        w, keep = caratheodory(bary, mass.clone())                  # :103-105
        scale = torch.zeros(S, dtype=dt, device=dev)
        scale[keep] = w / mass[keep]
        mu[body.reshape(-1)] = (mu[body] * scale.unsqueeze(0)).reshape(-1)   # :107-115
        last_kept = bool(scale[-1] > 0)
        if len(tail):                                               # :117-127
            mu[tail] = mu[tail] * scale[-1]
        survivors = body[:, keep].reshape(-1)
        alive = torch.cat([survivors, tail]) if (len(tail) and last_kept) else survivors


def objective_step(bary, obj_bary, w, keep):
    """The extra null-space step of the objective-aware variant (``SOBER/_rchq.py:177-196`` and
    :86-104): among the kept points, move along the null vector of [bary | 1]^T in the direction
    with obj . w_null >= 0 until one more weight vanishes.  Returns (w, keep) with one point less
    (unchanged when the kept system has no null space, i.e. <= q + 1 points)."""
    Xp = torch.cat([bary[keep].T, torch.ones(1, len(keep), dtype=bary.dtype)], 0)      # :178-180
    if Xp.shape[1] <= Xp.shape[0]:
        return w, keep
    _, _, Vh = torch.linalg.svd(Xp)                                                   # :181
    w_null = Vh[-1]                                                                   # :182
    if float(obj_bary[keep] @ w_null) < 0:                                            # :183-184
        w_null = -w_null
    pos = w_null > 0                                                                  # :187
    if not bool(pos.any()):
        return w, keep
    alpha = torch.where(pos, w / w_null, torch.tensor(float("inf"), dtype=w.dtype))   # :188-189
    k = int(torch.argmin(alpha))                                                      # :190-191
    w = w - alpha[k] * w_null                                                         # :192
    w[k] = 0.0                                                                        # :193
    live = w > 0                                                                      # :195-196
    return w[live], keep[live]


def tchernychova_lyons_objective(samp, U, pt_nys, kernel, calc_obj, mu=None):
    """``Mod_Tchernychova_Lyons`` with ``calc_obj`` (``SOBER/_rchq.py:48-219``), restated on the set
    structure of :func:`tchernychova_lyons` (BASQ's tail handling: SOBER's own adds the tail sums
    twice, :128-135 and :153-164, which breaks the moments - SURVEY 2b - and is not reproduced).
    The objective obj = -calc_obj(samp) (:69) rides along as one more test function (:138-150);
    after every Caratheodory step the kept sets lose one more member by :func:`objective_step`.
    Returns (w_star, idx_star) with at most q + 1 points."""
    N = len(samp)
    q, M = U.shape
    S = 2 * (q + 1)
    dt = U.dtype
    mu = torch.full((N,), 1.0 / N, dtype=dt) if mu is None else mu.to(dt).clone()
    obj = -1.0 * calc_obj(samp).to(dt)
    alive = torch.arange(N)[mu != 0]
    while True:
        R = len(alive)
        if R <= q + 1:
            idx = torch.nonzero(mu > 0).squeeze(1)
            return mu[idx], idx
        if R <= S:                                                                    # :76-111
            feats = (U @ kernel(pt_nys, samp[alive])).T
            ext = torch.cat([feats, obj[alive].unsqueeze(1)], 1)
            w, keep = caratheodory(ext, mu[alive])
            w, keep = objective_step(feats, obj[alive], w, keep)
            alive = alive[keep]
            mu = torch.zeros_like(mu)
            mu[alive] = w
            if len(alive) <= q + 1:
                return mu[mu > 0], alive
            continue
        E = R // S
        body = alive[: E * S].reshape(E, S)
        tail = alive[E * S:]
        ids = body.reshape(-1)
        Kb = kernel(pt_nys, samp[ids]) * mu[ids].unsqueeze(0)
        G = Kb.reshape(M, E, S).sum(1)
        bary = (U @ G).T
        objs = (obj[body] * mu[body]).sum(0)
        mass = mu[body].sum(0)
        if len(tail):
            bary[-1] += (U @ kernel(pt_nys, samp[tail])) @ mu[tail]
            objs[-1] += obj[tail] @ mu[tail]
            mass[-1] += mu[tail].sum()
        bary = bary / mass.unsqueeze(1)
        objs = objs / mass
        w, keep = caratheodory(torch.cat([bary, objs.unsqueeze(1)], 1), mass.clone())  # :171-175
        w, keep = objective_step(bary, objs, w, keep)                                  # :177-196
        scale = torch.zeros(S, dtype=dt)
        scale[keep] = w / mass[keep]
        mu[body.reshape(-1)] = (mu[body] * scale.unsqueeze(0)).reshape(-1)
        last_kept = bool(scale[-1] > 0)
        if len(tail):
            mu[tail] = mu[tail] * scale[-1]
        survivors = body[:, keep].reshape(-1)
        alive = torch.cat([survivors, tail]) if (len(tail) and last_kept) else survivors


def nystrom_basis(pt, s, kernel):
    """``ker_svd_sparsify`` (``BASQ/_rchq.py:28-31``): randomised rank-s SVD of K(pt, pt)
    (torch.svd_lowrank, niter=2, consumes the global torch RNG); returns (S, U [s, M])."""
    _U, S, _ = torch.svd_lowrank(kernel(pt, pt), q=s)
    return S, -_U.T


def recombination(pts_rec, pts_nys, num_pts, kernel, init_weights=None, chunk=1, U=None):
    """``recombination`` -> ``rc_kernel_svd`` (``BASQ/_rchq.py:4-40``).  Returns (idx, w).
    ``U`` may be supplied to bypass the RNG-dependent basis (parity tests)."""
    if U is None:
        _, U = nystrom_basis(pts_nys, num_pts - 1, kernel)
    w, idx = tchernychova_lyons(pts_rec, U, pts_nys, kernel, mu=init_weights, chunk=chunk)
    return idx, w


# --------------------------------------------------------------------------- metrics
def features(X, U, pt_nys, kernel, block=8192):
    """phi(X) = (U @ k(Z, X))^T as [N, q] fp64 (the test functions of SURVEY 0)."""
    out = []
    for i in range(0, len(X), block):
        out.append((U.double() @ kernel(pt_nys, X[i:i + block]).double()).T)
    return torch.cat(out, 0)


def moment_residual(Phi, mu, idx, w):
    """|Phi^T mu - Phi[idx]^T w| / |Phi^T mu| over the q test functions plus the mass."""
    Phi1 = torch.cat([torch.ones(len(Phi), 1, dtype=torch.float64), Phi.double()], 1)
    full = Phi1.T @ mu.double()
    red = Phi1[idx].T @ w.double()
    return float(torch.linalg.norm(full - red) / torch.linalg.norm(full))


def quadrature(X, w, mean_predict, kernel):
    """``KernelQuadrature.quadrature`` tail (``BASQ/_quadrature.py:60-62``):
    EZy = w . m(X),  VarZy = w^T K(X, X) w."""
    return float(w @ mean_predict(X)), float(w @ kernel(X, X) @ w)
