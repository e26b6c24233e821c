"""CPU: the C-ABI library loads and exports every symbol include/lasr.h declares (no compute calls)."""
import ctypes
import os

import pytest


def test_library_exports_every_declared_symbol():
    from liteasr_b200 import _lib, build
    build.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _lib.declared_symbols()
    assert "lasr_gemm" in names and "lasr_ctc_fwdbwd" in names
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/lasr.h but not exported"
    assert lib.lasr_arch() == 100


def test_missing_library_fails_loudly(monkeypatch):
    from liteasr_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", os.path.join(os.path.dirname(_lib.LIB_PATH), "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.lib()


def test_ops_reject_cpu_tensors():
    import torch
    from liteasr_b200 import ops
    a = torch.zeros(8, 8, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.gemm(a, a, torch.zeros(8, 8), 8, 8, 8, lda=8, ldb=8, ldc=8)
