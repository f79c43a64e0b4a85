"""The C-ABI library loads and exports every symbol include/ima2p_b200.h declares (no compute calls: this
test runs without a GPU), and refuses to work without a device instead of falling back to the CPU."""
import ctypes as C
import os
import re

import pytest

from ima2p_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "ima2p_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ima2p_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def product_lib():
    import __graft_entry__ as g
    g.build()
    return capi.bind(capi.LIB_PATH)


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(capi.SIGNATURES)


def test_library_exports_every_declared_symbol(product_lib):
    for name in declared_symbols():
        assert hasattr(product_lib, name), name


def test_no_cpu_fallback_without_a_device(product_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    h = C.c_void_p()
    rc = product_lib.ima2p_engine_create(C.byref(h), 0, 2, 2, 0, 1, 64, 1)
    assert rc == capi.E_CUDA and b"no CUDA device" in product_lib.ima2p_last_error()
    l = C.c_void_p()
    z = (C.c_double * 4)()
    rc = product_lib.ima2p_lmode_create(C.byref(l), 0, 3, 2, 1, z, z, z, z, z, 0)
    assert rc == capi.E_CUDA


def test_product_binding_refuses_the_test_emulation(monkeypatch):
    """The host emulation of the kernels exists for the CPU test-suite only: the package's own loader rejects it."""
    import subprocess
    subprocess.run([os.path.join(ROOT, "tests", "hostemu", "build.sh")], check=True)
    emu = os.path.join(ROOT, "tests", "hostemu", "libima2p_hostemu.so")
    assert b"tests only" in capi.bind(emu).ima2p_version()
    monkeypatch.setattr(capi, "LIB_PATH", emu)
    monkeypatch.setattr(capi, "_LIB", None)
    with pytest.raises(ImportError):
        capi.lib()
    monkeypatch.setattr(capi, "_LIB", None)
