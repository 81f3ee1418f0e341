"""The C-ABI library loads on a CPU-only box and exports every symbol include/basq_b200.h declares;
compute entry points fail loudly (no CPU fallback) when there is no CUDA device."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "basq_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(basq_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported():
    from basq_b200 import _lib
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(_lib.lib, n), f"{n} declared in include/basq_b200.h but not exported"
    # and the binding covers all of them
    assert set(names) == set(_lib.SYMBOLS)
    assert _lib.lib.basq_abi_version() == 2


def test_struct_layout_matches_header():
    from basq_b200 import _lib
    # 4 x int32, double, 32 doubles, 3 doubles, 2 x int32, 3 pointers
    assert ctypes.sizeof(_lib.KernelDesc) == 16 + 8 + 8 * 32 + 24 + 8 + 32


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_no_cpu_fallback():
    from basq_b200 import _lib, ops
    from basq_b200.kernels import KernelSpec
    h = ctypes.c_void_p()
    assert _lib.lib.basq_ctx_create(0, None, ctypes.byref(h)) == _lib.ERR_CUDA
    assert b"no CPU fallback" in _lib.lib.basq_last_error()
    spec = KernelSpec(_lib.RBF, _lib.PLAIN, torch.tensor([1.0]), 1.0)
    X = torch.randn(10, 2)
    with pytest.raises(_lib.BasqError):
        ops.gram(spec, X, X)
    import basq_b200
    with pytest.raises(_lib.BasqError):
        basq_b200.recombination(X, X[:4], 3, spec, torch.device("cpu"))


def test_binding_argument_counts_match_header():
    """Every ctypes signature in basq_b200/_lib.py has as many arguments as the header's prototype
    (a mismatch would silently corrupt the call on the GPU box)."""
    from basq_b200 import _lib
    text = open(os.path.join(ROOT, "include", "basq_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = dict(re.findall(r"\b(basq_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S))
    assert set(protos) == set(_lib.SYMBOLS)
    for name, args in protos.items():
        args = " ".join(args.split())
        n_header = 0 if args in ("", "void") else args.count(",") + 1
        fn = getattr(_lib.lib, name)
        assert len(fn.argtypes) == n_header, f"{name}: header has {n_header} arguments, binding {len(fn.argtypes)}"
