"""Golden vectors for the candidate-side stages, produced by the REFERENCE's own code where it can be
imported here: ``SOBER/_weights.py`` (WeightsStabiliser.cleansing_weights) needs only torch.
``BASQ/_sampler.py`` (calc_weights) and ``SOBER/_pi.py`` (lfi) import gpytorch; they are run by
oracle/make_golden_gp.py, which registers stand-in modules for those imports.

Run in the build container only:  ``python oracle/make_golden_candidates.py``
"""
from __future__ import annotations

import importlib.util
import os

import numpy as np
import torch

REF = os.environ.get("BASQ_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def main():
    spec = importlib.util.spec_from_file_location("ref_weights", os.path.join(REF, "SOBER", "_weights.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    ws = mod.WeightsStabiliser()
    g = torch.Generator().manual_seed(3)
    out = {}
    w = torch.rand(257, generator=g, dtype=torch.float64)
    w[::7] = 0.0
    w[5] = float("nan"); w[11] = float("inf"); w[13] = 1e-9; w[17] = -0.3; w[19] = 3e-8
    out["in_mixed"] = w.numpy().copy()
    out["out_mixed"] = ws.cleansing_weights(w.clone()).numpy()
    z = torch.zeros(64, dtype=torch.float64); z[3] = 1e-12
    out["in_zero"] = z.numpy().copy()
    out["out_zero"] = ws.cleansing_weights(z.clone()).numpy()
    p = torch.rand(1000, generator=g, dtype=torch.float64) * 5.0
    out["in_plain"] = p.numpy().copy()
    out["out_plain"] = ws.cleansing_weights(p.clone()).numpy()
    np.savez_compressed(os.path.join(OUT, "candidates.npz"), **out)
    print("written", os.path.normpath(os.path.join(OUT, "candidates.npz")))


if __name__ == "__main__":
    main()
