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
* ``oracle.gp_kernels`` (gpytorch kernels, predictive covariance, WSABI-L/M,
  MMLT) is PARITY UNPINNED: the arithmetic lives in ``gpytorch`` which is not
  installed here (no version pinned by the reference either) and the reference
  ships no tests or golden vectors.  The functions restate the published
  gpytorch formulas and the reference's call sites (cited per function).
"""
