"""CPU: the C-ABI shared library loads and exports every symbol include/asm_b200.h declares; argument
validation that needs no GPU behaves as documented.  No compute calls here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "asm_b200.h")


@pytest.fixture(scope="module")
def lib():
    from style_transfer_based_holographic_imaging_b200 import build, _lib
    build.build()          # nvcc cross-compiles for sm_100a without a GPU
    return _lib.load()


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(asm_b200_[a-z_0-9]+)\s*\(", src)))


def test_header_declares_expected_entry_points():
    names = declared_functions()
    for n in ["asm_b200_abi_version", "asm_b200_strerror", "asm_b200_workspace_bytes", "asm_b200_forward",
              "asm_b200_adjoint", "asm_b200_grad_z", "asm_b200_launch_count", "asm_b200_profile",
              "asm_b200_unwrap_workspace_bytes", "asm_b200_unwrap"]:
        assert n in names


def test_library_exports_every_declared_symbol(lib):
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} is declared in asm_b200.h but not exported"
    from style_transfer_based_holographic_imaging_b200 import _lib
    assert sorted(_lib.SYMBOLS) == declared_functions()


def test_abi_version_and_strerror(lib):
    assert lib.asm_b200_abi_version() == 1
    assert lib.asm_b200_strerror(0) == b"ok"
    for code in range(-7, 0):
        assert lib.asm_b200_strerror(code).startswith(b"asm_b200")


def test_workspace_bytes(lib):
    ws = lib.asm_b200_workspace_bytes
    assert ws(4, 1, 1024, 0) > 4 * 1024 * 1024 * 8 // 2
    assert ws(5, 1, 128, 1) > 0
    assert ws(1, 1, 4096, 0) > 0 and ws(1, 1, 2048, 1) > 0
    for bad in [(0, 1, 128, 0), (1, 0, 128, 0), (1, 1, 101, 1), (1, 1, 3000, 0), (1, 1, 8, 0), (1, 1, 4096, 1), (1, 1, 16, 0)]:
        assert ws(*bad) == 0, bad
    assert ws(1, 1, 16, 1) > 0      # 16 padded -> FFT 32
    assert ws(3, 1, 92, 1) > 0 and ws(3, 1, 100, 0) > 0 and ws(1, 1, 45, 0) > 0   # not powers of two: matrix-product path


def test_argument_validation_without_gpu(lib):
    E_NULL, E_SHAPE, E_MODE, E_WORKSPACE, E_OPTICS = -1, -2, -3, -4, -5
    vp = ctypes.c_void_p
    dummy = vp(256)   # never dereferenced: validation fails first
    f = lib.asm_b200_forward
    assert f(dummy, None, dummy, 0, dummy, None, 1, 1, 101, 1, 0, 0, 532e-9, 1.5e-6, 1.0, 1.0, dummy, 1 << 30, None) == E_SHAPE
    assert f(dummy, None, dummy, 0, dummy, None, 1, 1, 128, 0, 9, 0, 532e-9, 1.5e-6, 1.0, 1.0, dummy, 1 << 30, None) == E_MODE
    assert f(dummy, None, dummy, 0, dummy, None, 1, 1, 128, 0, 1, 0, 532e-9, 1.5e-6, 1.0, 1.0, dummy, 1 << 30, None) == E_NULL   # amp/phase needs in1
    assert f(dummy, None, dummy, 0, dummy, None, 1, 1, 128, 0, 0, 0, -1.0, 1.5e-6, 1.0, 1.0, dummy, 1 << 30, None) == E_OPTICS
    assert f(dummy, None, dummy, 0, dummy, None, 1, 1, 128, 0, 0, 0, 532e-9, 1.5e-6, 1.0, 1.0, dummy, 16, None) == E_WORKSPACE
    assert f(dummy, None, dummy, 0, dummy, None, 1, 1, 128, 0, 0, 0, 532e-9, 1.5e-6, 1.0, 1.0, vp(257), 1 << 30, None) == E_WORKSPACE
    assert f(dummy, None, dummy, 0, dummy, None, 1, 2, 128, 0, 0, 3, 532e-9, 1.5e-6, 1.0, 1.0, dummy, 1 << 30, None) == E_MODE  # cat outputs need C == 1
    a = lib.asm_b200_adjoint
    assert a(dummy, None, dummy, 0, None, None, dummy, None, 1, 1, 128, 0, 1, 0, 532e-9, 1.5e-6, 1.0, dummy, 1 << 30, None) == E_MODE
    g = lib.asm_b200_grad_z
    assert g(dummy, None, dummy, 0, None, None, 0, dummy, 1, 1, 128, 0, 0, 532e-9, 1.5e-6, 1.0, dummy, 1 << 30, None) == E_NULL
