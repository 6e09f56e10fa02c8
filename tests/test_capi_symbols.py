"""CPU: libb2m.so loads without a GPU and exports every entry point include/b2m.h declares; calls that
need a device fail with an error code, never a crash, and there is no CPU fallback to fall into."""
import ctypes
import os
import re

from marlin_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "b2m.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b2m_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported():
    L = _lib.lib()
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/b2m.h but not exported by libb2m.so"


def test_version_and_error_paths_without_gpu():
    L = _lib.lib()
    assert b"sm_100a" in L.b2m_version()
    h = ctypes.c_void_p()
    import torch
    if not torch.cuda.is_available():
        rc = L.b2m_ctx_create(0, ctypes.byref(h))
        assert rc == 8  # B2M_ERR_CUDA: no device, and no CPU fallback
        assert L.b2m_last_error()
    assert L.b2m_ntt(None, 0, None, 3, 0, 0) == 1  # B2M_ERR_INVALID_ARG, not a crash


def test_product_does_not_import_the_oracle():
    """The product path must never route through oracle/ (or any CPU fallback)."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "marlin_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in src and "from oracle" not in src and "oracle/" not in src, f
