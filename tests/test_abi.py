"""The C-ABI library loads and exports every symbol include/edb.h declares (no compute calls)."""
import os
import re

import pytest

from easydist_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "edb.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(edb_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_bound_and_exported():
    if not os.path.exists(build.LIB_PATH) and build._nvcc() is None:
        pytest.skip("libedb.so not built and no nvcc here")
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in edb.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == declared
    assert lib.edb_version() == 100
    assert lib.edb_is_initialized() == 0


def test_product_has_no_cpu_fallback():
    """Host ops refuse CPU tensors instead of silently computing on the host."""
    import torch
    from easydist_b200 import reshard
    with pytest.raises(_lib.EdbError):
        reshard.all_gather_start(torch.zeros(4), 0, [0])
    with pytest.raises(_lib.EdbError):
        reshard.scatter_wrapper(torch.zeros(4, 4), 2, 0, 0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "easydist_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
