"""Generate golden vectors under tests/golden/ by running the REFERENCE's own
``BASQ/_rchq.py`` (imported from the read-only tree; it needs only torch) on seeded inputs.

Run in the build container only:  ``python oracle/make_golden.py``
(/root/reference does not exist on the GPU box; the committed .npz files travel instead.)

The reference allocates its work tensors with the torch default dtype, so the default is
switched to float64 here: the stored outputs are the reference's fp64 behaviour.
"""
from __future__ import annotations

import math
import os
import sys

import numpy as np
import torch

REF = os.environ.get("BASQ_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def rbf(ls, os_=1.0):
    def k(x, y):
        d = x.unsqueeze(1) / ls - y.unsqueeze(0) / ls
        return os_ * torch.exp(-0.5 * (d * d).sum(-1))
    return k


def matern52(ls, os_=1.0):
    def k(x, y):
        d = x.unsqueeze(1) / ls - y.unsqueeze(0) / ls
        r2 = (d * d).sum(-1)
        r = torch.sqrt(r2)
        c = math.sqrt(5.0)
        return os_ * (1 + c * r + 5.0 / 3.0 * r2) * torch.exp(-c * r)
    return k


def main():
    sys.path.insert(0, REF)
    torch.set_default_dtype(torch.float64)
    from BASQ import _rchq as ref  # the reference implementation itself

    os.makedirs(OUT, exist_ok=True)
    dev = torch.device("cpu")

    # ---- Caratheodory known-answer tests (BASQ/_rchq.py:133-175)
    car = {}
    for tag, (S, q, seed) in {"a": (20, 9, 1), "b": (64, 31, 2), "c": (200, 99, 3), "d": (37, 12, 4)}.items():
        g = torch.Generator().manual_seed(seed)
        X = torch.randn(S, q, generator=g)
        mu = torch.rand(S, generator=g) + 0.05
        mu = mu / mu.sum()
        w, idx, *_ = ref.Tchernychova_Lyons_CAR(X.clone(), mu.clone(), dev)
        car.update({f"X_{tag}": X.numpy(), f"mu_{tag}": mu.numpy(), f"w_{tag}": w.numpy(), f"idx_{tag}": idx.numpy()})
    np.savez_compressed(os.path.join(OUT, "car_kat.npz"), **car)

    # ---- Tchernychova-Lyons loop with a given basis U (BASQ/_rchq.py:43-130)
    cases = {
        # tag: (N, d, M, num_pts, kernel family, lengthscale, seed)
        "rbf_d3": (3000, 3, 50, 8, "rbf", 1.2, 11),
        "rbf_d10": (5000, 10, 80, 16, "rbf", 2.5, 12),
        "m52_d5": (2500, 5, 64, 12, "matern52", 2.0, 13),
        "final_only": (30, 4, 40, 16, "rbf", 1.5, 14),      # n+1 < N <= S : final stage only
        "trivial": (10, 4, 40, 16, "rbf", 1.5, 15),         # N <= n+1 : returned untouched
        "exact_mult": (4 * 2 * 10, 3, 30, 10, "rbf", 1.0, 16),  # N multiple of S: no tail
    }
    tl = {}
    for tag, (N, d, M, n, fam, ls, seed) in cases.items():
        torch.manual_seed(seed)
        X = math.sqrt(2.0) * torch.randn(N, d)
        Z = math.sqrt(2.0) * torch.randn(M, d)
        kern = rbf(ls) if fam == "rbf" else matern52(ls)
        torch.manual_seed(seed + 1000)
        _, U = ref.ker_svd_sparsify(Z, n - 1, kern, dev)
        w, idx = ref.Mod_Tchernychova_Lyons(X, U, Z, kern, dev)
        tl.update({f"X_{tag}": X.numpy(), f"Z_{tag}": Z.numpy(), f"U_{tag}": U.numpy(),
                   f"w_{tag}": w.numpy(), f"idx_{tag}": idx.numpy(),
                   f"meta_{tag}": np.array([N, d, M, n, 0 if fam == "rbf" else 1, ls], dtype=np.float64)})
        # the public entry point must be the same thing (BASQ/_rchq.py:4-40) when re-seeded
        torch.manual_seed(seed + 1000)
        idx2, w2 = ref.recombination(X, Z, n, kern, dev)
        assert torch.equal(idx2, idx) and torch.equal(w2, w)
    np.savez_compressed(os.path.join(OUT, "tl_cases.npz"), **tl)
    print("golden vectors written to", os.path.normpath(OUT))


if __name__ == "__main__":
    main()
