"""CPU checks of the drop-in boundary: libifl_b200.so builds for sm_100a, loads, and
exports every symbol include/ifl_b200.h declares.  No compute call is made here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "ifl_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ifl_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(ifl):
    if not os.path.exists(ifl.library_path()):
        ifl.build_library()
    lib = ctypes.CDLL(ifl.library_path())
    names = header_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), n


def test_binding_covers_header(ifl):
    from importlib import import_module
    binding = import_module("incremental-fluids_b200.binding")
    assert sorted(binding.EXPORTS) == header_symbols()


def test_no_cpu_fallback(ifl):
    """Without a CUDA device the product must fail loudly, not fall back to a CPU path."""
    from conftest import has_cuda
    if has_cuda():
        pytest.skip("CUDA device present")
    with pytest.raises(ifl.IflError):
        ifl.FluidSolver(32, 32, 0.1)


def test_product_does_not_reference_oracle():
    pkg = os.path.join(ROOT, "incremental-fluids_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")) or f == "Makefile":
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower(), os.path.join(dirpath, f)
