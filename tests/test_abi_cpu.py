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


def test_gemm_args_ctypes_layout_matches_the_header():
    """The ctypes mirror of ``lasr_gemm_args`` must list the same fields, in the same order and with the same C types as
    include/lasr.h (a drift would silently shift every later argument)."""
    import re
    from liteasr_b200 import _lib
    src = open(_lib.HEADER_PATH).read()
    body = re.search(r"typedef struct lasr_gemm_args \{(.*?)\} lasr_gemm_args;", src, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    ctype_of = {"const void*": ctypes.c_void_p, "void*": ctypes.c_void_p, "const float*": ctypes.c_void_p, "float*": ctypes.c_void_p,
                "int32_t": ctypes.c_int32, "int64_t": ctypes.c_int64, "float": ctypes.c_float, "uint32_t": ctypes.c_uint32}
    header = []
    for decl in body.split(";"):
        decl = " ".join(decl.split())
        if not decl:
            continue
        m = re.match(r"((?:const )?\w+\s?\*?)\s*(.+)", decl)
        ctype, names = m.group(1).replace(" *", "*").strip(), [n.strip() for n in m.group(2).split(",")]
        header += [(n, ctype_of[ctype]) for n in names]
    mirror = [(n, t) for n, t in _lib.GemmArgs._fields_]
    assert [n for n, _ in mirror] == [n for n, _ in header]
    for (n, t), (_, th) in zip(mirror, header):
        assert t is th, (n, t, th)
