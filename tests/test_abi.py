"""The C-ABI library loads and exports every symbol include/svscope_b200.h declares (CPU only,
no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "svscope_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(svs_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from svscope_b200 import _lib
    from svscope_b200.csrc import build
    build.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 15
    for name in names:
        assert hasattr(lib, name), name
    # the ctypes table binds exactly the declared entry points
    assert sorted(_lib.SYMBOLS) == names
    L = _lib.load()
    assert b"sm_100a" in L.svs_version()


def test_no_cpu_fallback_without_device():
    import torch
    from svscope_b200 import _lib
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(RuntimeError):
        _lib.Context(0)
    from svscope_b200.spoa import poa
    with pytest.raises(RuntimeError):
        poa(["ACGT", "ACGT"], 1)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "svscope_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "oracle/" not in src, f


def test_scripts_do_not_use_the_oracle():
    """Only tests/, smoke() and bench.py's CPU legs may touch oracle/: the probes that compare
    against it live in tests/tools/."""
    sdir = os.path.join(ROOT, "scripts")
    for dirpath, _, files in os.walk(sdir):
        for f in files:
            if f.endswith((".py", ".sh", ".cu")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
