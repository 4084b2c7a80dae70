"""CPU: the C-ABI library loads and exports every symbol include/emu_b200.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "emu_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(emub_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_documented_surface():
    from madaiemulator_b200 import engine
    assert set(_declared()) == set(engine.SYMBOLS)


def test_library_exports_every_declared_symbol():
    from madaiemulator_b200 import engine
    if not os.path.exists(engine.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    L = ctypes.CDLL(engine.LIB_PATH)
    for name in _declared():
        assert hasattr(L, name), name
    L.emub_version.restype = ctypes.c_char_p
    assert b"sm_100a" in L.emub_version()


def test_no_cpu_fallback_without_gpu():
    """The product path must fail loudly when there is no CUDA device (this container has none)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    from madaiemulator_b200 import engine
    with pytest.raises(engine.EmubError):
        engine.Context(0)


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "madaiemulator_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".c", ".h", ".cu", ".cuh", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "pyoracle" not in txt and "emu_oracle" not in txt and "libemu_ref" not in txt, f
