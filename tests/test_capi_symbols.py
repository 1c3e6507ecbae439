"""The C-ABI library loads on a machine without a GPU and exports every symbol
include/marian_b200.h declares (no compute calls here)."""
import ctypes
import os
import re

from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "marian_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mrn_\w+)\s*\(", text)))


def test_header_and_binding_agree(pkg):
    syms = declared_symbols()
    assert len(syms) > 50
    bound = set(pkg.DECLARED_SYMBOLS) | {"mrn_test_golden"}
    assert set(syms) == bound, set(syms) ^ bound


def test_product_library_exports_every_declared_symbol(pkg):
    assert os.path.exists(pkg.LIB_PATH), "run python __graft_entry__.py build"
    lib = ctypes.CDLL(pkg.LIB_PATH)
    tests = ctypes.CDLL(pkg.TESTS_LIB_PATH)
    for s in declared_symbols():
        holder = tests if s == "mrn_test_golden" else lib
        assert hasattr(holder, s), s
    lib.mrn_backend_name.restype = ctypes.c_char_p
    assert lib.mrn_backend_name() == b"cuda"


def test_product_never_links_the_oracle(pkg):
    import subprocess

    out = subprocess.run(["ldd", pkg.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out
