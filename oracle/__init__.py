"""CPU oracle for the BASQ kernel-recombination hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and there only as the checker (or as
the thing timed on the host cores), never on the CUDA product path.

The reference (``ma921/BASQ``) is pure Python/torch, so the restatement is
written with torch CPU tensor ops (fp64 unless told otherwise) - the same
library the reference itself runs on, which also makes the timed CPU baseline
a fair multi-threaded one.

Pinning status
--------------
* ``oracle.rchq`` (recombination / Tchernychova-Lyons / Caratheodory) is
  PINNED: ``oracle/make_golden.py`` imports the reference's own
  ``BASQ/_rchq.py`` (it needs only torch) in the build container, runs it on
  seeded inputs and stores inputs + outputs under ``tests/golden/``;
  ``tests/test_oracle_golden.py`` checks the restatement against them.
* ``oracle.gp_kernels`` (predictive covariance, WSABI-L/M, MMLT, the adaptor
  objects) and ``oracle.sampler.calc_weights`` / ``lfi`` are PINNED at the
  reference's own call sites: ``oracle/make_golden_gp.py`` imports the
  reference's ``BASQ/_gp.py``, ``_wsabi.py``, ``_vbq.py``, ``_sampler.py``,
  ``SOBER/_gp.py``, ``_kernel.py``, ``_pi.py`` and ``SOBER/BASQ/_scale_mmlt.py``
  unmodified (empty stand-in modules satisfy their ``import gpytorch / botorch
  / matplotlib`` lines) and runs their functions on a duck-typed exact-GP
  model; ``tests/test_oracle_golden_gp.py`` checks the restatement against
  the stored outputs (``tests/golden/gp_kernels.npz``) to 1e-10.
  What stays "published algorithm restated" is the third-party layer below
  those call sites: gpytorch itself (absent from ``/root/reference`` and from
  this image, no version pinned by the reference) - ``ScaleKernel(RBFKernel /
  MaternKernel).forward``, the exact posterior mean / variance, the
  Gaussian-likelihood noise term and ``prediction_strategy.covar_cache``
  (S with S S^T = (K + noise I)^-1).  The reference's ``fast_pred_var`` (LOVE)
  variance is an approximation of the exact one used here.
"""
