"""The C-ABI shared library: loads, exports every symbol include/fvy.h declares, and fails loudly
(no CPU fallback) when there is no CUDA device.  No compute calls here."""
import ctypes as C
import os
import re

import pytest

from face_vijnana_yolov3_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "fvy.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fvy_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = L.load()
    declared = _header_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/fvy.h but not exported by libfvy.so"
    assert sorted(L.SYMBOLS) == declared, "ctypes binding list and header disagree"


def test_struct_layouts():
    assert C.sizeof(L.FvyDet) == 32
    assert C.sizeof(L.FvyConfig) == 40
    assert C.sizeof(L.FvyPostParams) == 8 + 8 + 4 + 4 + 72 + 4 + 4   # trailing pad to 8


def test_version_and_error_strings():
    lib = L.load()
    assert b"sm_100a" in lib.fvy_version()
    assert isinstance(lib.fvy_last_error(), bytes)


def test_invalid_arguments_are_rejected_without_touching_the_gpu():
    lib = L.load()
    h = C.c_void_p()
    cfg = L.FvyConfig(device=0, net_h=400, net_w=416, head=L.HEAD_YOLO3, nb_class=1, bb_info_c_size=6, max_batch=1)
    assert lib.fvy_create(C.byref(cfg), C.byref(h)) == L.FVY_E_INVALID and b"multiple of 32" in lib.fvy_last_error()
    cfg = L.FvyConfig(device=0, net_h=416, net_w=416, head=7, nb_class=1, bb_info_c_size=6, max_batch=1)
    assert lib.fvy_create(C.byref(cfg), C.byref(h)) == L.FVY_E_INVALID
    cfg = L.FvyConfig(device=0, net_h=416, net_w=416, head=L.HEAD_YOLO3, nb_class=0, bb_info_c_size=6, max_batch=1)
    assert lib.fvy_create(C.byref(cfg), C.byref(h)) == L.FVY_E_INVALID
    assert lib.fvy_create(None, C.byref(h)) == L.FVY_E_INVALID
    assert lib.fvy_forward(None, None, 0, 1, None, None, None) == L.FVY_E_INVALID
    with pytest.raises(ValueError):
        L.check(L.FVY_E_INVALID)
    with pytest.raises(OverflowError):
        L.check(L.FVY_E_RANGE)


def test_no_cpu_fallback():
    """Without a CUDA device creation fails with FVY_E_CUDA; with one it succeeds.  Either way the
    product never computes on the host."""
    import torch
    lib = L.load()
    h = C.c_void_p()
    cfg = L.FvyConfig(device=0, net_h=64, net_w=64, head=L.HEAD_NONE, nb_class=1, bb_info_c_size=6, max_batch=1)
    rc = lib.fvy_create(C.byref(cfg), C.byref(h))
    if torch.cuda.is_available():
        assert rc == L.FVY_OK
        lib.fvy_destroy(h)
    else:
        assert rc == L.FVY_E_CUDA and b"no CPU fallback" in lib.fvy_last_error()
        with pytest.raises(L.FvyError):
            L.check(rc)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "face_vijnana_yolov3_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f"{f} imports the oracle"
                assert "liboracle" not in txt
