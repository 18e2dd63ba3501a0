"""The C-ABI library loads on a CPU-only box and exports every symbol include/b200pcg.h declares;
without a device every compute entry point fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from firefoam_dev_b200 import _lib
from conftest import ROOT, has_gpu


def header_functions():
    src = open(os.path.join(ROOT, "include", "b200pcg.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_\w+)\s*\(", src)))


def test_header_and_binding_agree():
    assert header_functions() == sorted(_lib.ABI_SYMBOLS)


def test_library_exports_every_declared_symbol():
    L = C.CDLL(_lib.PCG_SO)
    for name in header_functions():
        assert hasattr(L, name), name
    assert L.b200_abi_version() == 2


def test_struct_layouts():
    assert C.sizeof(_lib.Controls) == 32
    assert C.sizeof(_lib.Perf) == 72
    assert C.sizeof(_lib.Iface) == 24


@pytest.mark.skipif(has_gpu(), reason="CPU-only behaviour")
def test_no_cpu_fallback():
    L = _lib.load_pcg()
    assert L.b200_device_count() == 0
    h = C.c_void_p()
    rc = L.b200_ctx_create(-1, 0, 1, None, C.byref(h))
    assert rc == _lib.B200_ENODEVICE and not h
    assert b"no CPU fallback" in L.b200_last_error(None)
    from firefoam_dev_b200 import Context, B200Error
    with pytest.raises(B200Error):
        Context()


def test_bad_arguments_rejected_before_any_device_work():
    L = _lib.load_pcg()
    h = C.c_void_p()
    assert L.b200_ctx_create(-1, 2, 2, None, C.byref(h)) == _lib.B200_EINVAL      # rank >= nranks
    assert L.b200_ctx_create(-1, 0, 2, None, C.byref(h)) == _lib.B200_EINVAL      # no uid
    assert L.b200_set_addressing(None, 0, 0, 0, None, None, 0, None) == _lib.B200_EINVAL
