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
    assert C.sizeof(_lib.SmoothControls) == 40


def test_ctypes_mirrors_match_the_header(tmp_path):
    """sizeof / offsetof of every struct of include/b200pcg.h as a C compiler lays it out == the ctypes mirrors the
    tests, the bench and replay.py use (b200_dump grew three fields in ABI version 2)."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if not gcc:
        pytest.skip("no gcc")
    fields = {"b200_controls": (_lib.Controls, ["tolerance", "relTol", "maxIter", "minIter", "precond", "reserved"]),
              "b200_smooth_controls": (_lib.SmoothControls, ["tolerance", "relTol", "maxIter", "minIter", "nSweeps", "smoother",
                                                             "sweepMode", "reserved"]),
              "b200_perf": (_lib.Perf, ["initialResidual", "finalResidual", "normFactor", "nIterations", "converged", "singular",
                                        "nColours", "solveMs", "setupMs", "h2dMs", "d2hMs"]),
              "b200_iface": (_lib.Iface, ["nbrRank", "nFaces", "faceCells", "tag"]),
              "b200_prgh_terms": (_lib.PrghTerms, [f[0] for f in _lib.PrghTerms._fields_]),
              "b200_dump": (_lib.Dump, [f[0] for f in _lib.Dump._fields_])}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "b200pcg.h"', "int main(void) {"]
    for st, (_, names) in fields.items():
        lines.append(f'    printf("{st} %zu\\n", sizeof({st}));')
        for n in names:
            lines.append(f'    printf("{st}.{n} %zu\\n", offsetof({st}, {n}));')
    lines += ["    return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    r = subprocess.run([gcc, "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True).stdout.splitlines())
    for st, (cls, names) in fields.items():
        assert int(out[st]) == C.sizeof(cls), st
        for n in names:
            assert int(out[f"{st}.{n}"]) == getattr(cls, n).offset, f"{st}.{n}"


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
