"""Matrix dump / replay format: the C-ABI writer/reader (csrc/dump.cpp, the code
the OpenFOAM adapter calls), the pure-numpy reader, the committed fixtures, and -- on a GPU -- the
replay of the fixtures through the CUDA path and the native tools/b200replay binary."""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import pytest

from firefoam_dev_b200 import _lib, meshgen as mg, replay
from oracle import oracle as orc
from conftest import ROOT, has_gpu

GOLD = os.path.join(ROOT, "tests", "golden")


def test_fixture_steckler_pins_the_golden_log_count():
    """The committed steckler dump is the system behind log.fireFoam:92: the oracle needs the log's 29
    DICPCG iterations on exactly the bytes in the file, and reproduces the stored solution."""
    d = replay.read_dump(os.path.join(GOLD, "steckler_ph_rgh_c1.b200sys"))
    log = json.load(open(os.path.join(GOLD, "steckler_log.json")))
    assert d.fieldName == "ph_rgh" and d.system.addr.nCells == 9000 and d.system.addr.nFaces == 24868
    assert d.reference["nIterations"] == log["ph_rgh"][0]["iters"] == 29
    assert d.controls["preconditioner"] == "DIC" and d.controls["relTol"] == 0.01
    psi = d.psi0.copy()
    perf = orc.pcg_solve(d.system, psi, "DIC", d.controls["tolerance"], d.controls["relTol"], d.controls["maxIter"])
    assert perf.nIterations == 29
    assert perf.finalResidual == d.reference["finalResidual"]
    # ... and the log's printed final residual (0.0080439052) to its 8 digits
    assert perf.finalResidual == pytest.approx(log["ph_rgh"][0]["final"], rel=1e-7)
    assert np.array_equal(psi, d.psi)


def test_fixture_G_equation():
    """the P1 G equation (P1.C:238-244) as a committed system; the oracle reproduces the stored
    DICPCG line on the bytes in the file."""
    d = replay.read_dump(os.path.join(GOLD, "steckler_G_p1.b200sys"))
    assert d.fieldName == "G" and d.system.addr.nCells == 9000 and d.controls["relTol"] == 0.0
    psi = d.psi0.copy()
    perf = orc.pcg_solve(d.system, psi, "DIC", d.controls["tolerance"], d.controls["relTol"], d.controls["maxIter"])
    assert perf.nIterations == d.reference["nIterations"] == 62 and np.array_equal(psi, d.psi)


def test_fixture_singlebox():
    d = replay.read_dump(os.path.join(GOLD, "singlebox_ph_rgh_c1.b200sys"))
    assert d.system.addr.nCells == 245 and d.controls["preconditioner"] == "diagonal"
    psi = d.psi0.copy()
    perf = orc.pcg_solve(d.system, psi, "diagonal", d.controls["tolerance"], d.controls["relTol"], d.controls["maxIter"])
    assert perf.nIterations == d.reference["nIterations"] and np.array_equal(psi, d.psi)


def test_fixture_U_transport_asymmetric():
    """an asymmetric system with the reference's U controls (smoothSolver + symGaussSeidel, tol 1e-6,
    maxIter 10; fvSolution:48-55) as a committed dump: `lower` travels, the controls are smoothSolver's, and the
    oracle reproduces the stored line on the bytes in the file."""
    d = replay.read_dump(os.path.join(GOLD, "steckler_U_transport.b200sys"))
    s = d.system
    assert d.fieldName == "Ux" and s.addr.nCells == 9000 and s.lower is not None and d.header["symmetric"] is False
    assert np.abs(s.lower - s.upper).max() > 0
    assert d.smooth == {"smoother": "symGaussSeidel", "tolerance": 1e-6, "relTol": 0.0, "maxIter": 10, "minIter": 0,
                        "nSweeps": 1, "B200": {"sweepMode": "exact"}}
    psi = d.psi0.copy()
    perf = orc.smooth_solve(s, psi, smoother="symGaussSeidel", tolerance=1e-6, relTol=0.0, maxIter=10)
    assert perf.nIterations == d.reference["nIterations"] == 7 and perf.finalResidual == d.reference["finalResidual"]
    assert np.array_equal(psi, d.psi)
    # the symmetric fixtures are untouched by the format extension
    assert replay.read_dump(os.path.join(GOLD, "steckler_G_p1.b200sys")).smooth is None


def test_roundtrip_asymmetric_smooth_controls(tmp_path):
    """write (C ABI) -> read (numpy, C ABI) of an asymmetric system with smoothSolver controls"""
    from firefoam_dev_b200 import cases
    t = cases.transport_system(mg.hex_block(6, 5, 4), seed=3)
    L = _lib.load_pcg()
    for sm, mode, nS in (("GaussSeidel", "multicolour", 2), ("symGaussSeidel", "exact", 1)):
        p = tmp_path / f"U_{sm}.b200sys"
        ctl = {"smoother": sm, "tolerance": 1e-7, "relTol": 0.1, "maxIter": 12, "minIter": 1, "nSweeps": nS,
               "B200": {"sweepMode": mode}}
        replay.write_dump(p, t, np.zeros(t.addr.nCells), ctl, fieldName="Uy",
                          reference={"initialResidual": 1.0, "finalResidual": 1e-8, "nIterations": 4}, solverName="smoothSolver")
        d = replay.read_dump(p)
        assert d.smooth == ctl and d.controls == ctl
        assert np.array_equal(d.system.lower, t.lower) and np.array_equal(d.system.upper, t.upper)
        assert all(int(x["offset"]) % 64 == 0 for x in d.header["arrays"])
        h = C.c_void_p()
        assert L.b200_dump_read(str(p).encode(), C.byref(h)) == 0, L.b200_dump_last_error()
        dd = L.b200_dump_get(h).contents
        assert dd.haveSmooth == 1 and dd.smooth.nSweeps == nS and dd.smooth.maxIter == 12 and dd.smooth.minIter == 1
        assert dd.smooth.smoother == {"GaussSeidel": 0, "symGaussSeidel": 1}[sm]
        assert dd.smooth.sweepMode == {"multicolour": 0, "exact": 1}[mode]
        assert dd.smooth.tolerance == 1e-7 and dd.smooth.relTol == 0.1
        got = np.ctypeslib.as_array(C.cast(dd.lower, C.POINTER(C.c_double)), shape=(t.addr.nFaces,))
        assert np.array_equal(got, t.lower)
        L.b200_dump_free(h)
    # a symmetric PCG dump has neither
    p = tmp_path / "sym.b200sys"
    b = mg.hex_block(4, 4, 4)
    replay.write_dump(p, b, np.zeros(b.addr.nCells), {"preconditioner": "diagonal"})
    h = C.c_void_p()
    assert L.b200_dump_read(str(p).encode(), C.byref(h)) == 0
    dd = L.b200_dump_get(h).contents
    assert dd.haveSmooth == 0 and not dd.lower
    L.b200_dump_free(h)


def _pbicg_dump(tmp_path, mode="exact"):
    from firefoam_dev_b200 import cases
    t = cases.transport_system(mg.hex_block(9, 8, 7), seed=4, kappa=0.05)
    psi = np.zeros(t.addr.nCells)
    perf = orc.pbicg_solve(t, psi, "DILU", tolerance=1e-8, relTol=0.0, maxIter=1000)
    ctl = {"solver": "PBiCG", "preconditioner": "DILU", "tolerance": 1e-8, "relTol": 0.0, "maxIter": 1000, "minIter": 0}
    if mode == "exact":
        ctl["B200"] = {"diluMode": "exact"}
    p = tmp_path / f"Yi_{mode}.b200sys"
    replay.write_dump(p, t, np.zeros(t.addr.nCells), ctl, fieldName="Yi", psi=psi,
                      reference={"initialResidual": perf.initialResidual, "finalResidual": perf.finalResidual,
                                 "nIterations": perf.nIterations, "converged": perf.converged}, solverName="DILUPBiCG")
    return p, t, psi, perf, ctl


def test_roundtrip_pbicg_controls(tmp_path):
    """a PBiCG solve travels too: `lower`, "solver": "PBiCG", the asymmetric preconditioner code"""
    L = _lib.load_pcg()
    for mode, code in (("exact", 3), ("multicolour", 2)):
        p, t, psi, perf, ctl = _pbicg_dump(tmp_path, mode)
        d = replay.read_dump(p)
        assert d.bicg and d.smooth is None and d.header["controls"]["solver"] == "PBiCG"
        assert d.header["controls"]["precondCode"] == code and d.header["symmetric"] is False
        want = {k: v for k, v in ctl.items() if k != "solver"}
        assert d.controls == want
        assert np.array_equal(d.system.lower, t.lower) and np.array_equal(d.psi, psi)
        h = C.c_void_p()
        assert L.b200_dump_read(str(p).encode(), C.byref(h)) == 0, L.b200_dump_last_error()
        dd = L.b200_dump_get(h).contents
        assert dd.havePBiCG == 1 and dd.haveSmooth == 0 and dd.controls.precond == code and dd.controls.maxIter == 1000
        assert dd.perf.nIterations == perf.nIterations and dd.solverName == b"DILUPBiCG"
        L.b200_dump_free(h)
    # neither flag on the other kinds of dump
    for name in ("steckler_U_transport.b200sys", "steckler_G_p1.b200sys"):
        h = C.c_void_p()
        assert L.b200_dump_read(os.path.join(GOLD, name).encode(), C.byref(h)) == 0
        assert L.b200_dump_get(h).contents.havePBiCG == 0
        L.b200_dump_free(h)
        assert not replay.read_dump(os.path.join(GOLD, name)).bicg


@pytest.mark.gpu
def test_replay_pbicg_dump_on_gpu(tmp_path):
    """python replay and the native tool re-solve a PBiCG dump through b200_bicg_solve: level-scheduled DILU needs
    the dumped (oracle) iteration count"""
    from firefoam_dev_b200 import Context
    p, t, psi_ref, perf_ref, _ = _pbicg_dump(tmp_path, "exact")
    c = Context(device=0)
    try:
        psi, perf, d = replay.replay(str(p), context=c)
        assert perf.nIterations == perf_ref.nIterations == d.reference["nIterations"]
        assert np.abs(psi - psi_ref).max() <= 1e-10 * np.abs(psi_ref).max()
        assert str(perf).startswith("DILUB200PBiCG:  Solving for Yi, Initial residual = 1, ")
    finally:
        c.close()
    exe = os.path.join(ROOT, "firefoam-dev_b200", "b200replay")
    if os.path.exists(exe):
        r = subprocess.run([exe, str(p)], capture_output=True, text=True, timeout=120)
        assert r.returncode == 0, r.stdout + r.stderr
        assert "DILUB200PBiCG:  Solving for Yi, Initial residual = 1, " in r.stdout and "MISMATCH" not in r.stdout
        assert r.stdout.count(f"No Iterations {perf_ref.nIterations}") == 2


def test_roundtrip_multirank_with_interfaces(tmp_path):
    """write (C ABI) -> read (numpy) and read (C ABI): every array bit-identical, including processor
    interfaces; arrays 64-byte aligned; empty patches and ragged sizes survive."""
    subs = [mg.hex_block(7, 6, 5, 2, 2, 1, r) for r in range(4)]
    L = _lib.load_pcg()
    for r, s in enumerate(subs):
        p = tmp_path / f"p_rgh_3_p{r}.b200sys"
        psi0 = np.random.default_rng(r).standard_normal(s.addr.nCells)
        ctl = {"preconditioner": "diagonal", "tolerance": 1e-7, "relTol": 0.05, "maxIter": 321, "minIter": 2}
        replay.write_dump(p, s, psi0, ctl, fieldName="p_rgh", rank=r, nranks=4, solveIndex=3, time=0.125,
                          reference={"initialResidual": 0.5, "finalResidual": 1e-8, "nIterations": 17},
                          solverName="diagonalPCG")
        d = replay.read_dump(p)
        assert (d.rank, d.nranks, d.fieldName) == (r, 4, "p_rgh")
        assert d.header["solveIndex"] == 3 and d.header["time"] == 0.125
        assert d.controls == ctl and d.reference["nIterations"] == 17 and d.psi is None
        a, b = d.system.addr, s.addr
        assert np.array_equal(a.lowerAddr, b.lowerAddr) and np.array_equal(a.upperAddr, b.upperAddr)
        for x, y in ((d.system.diag, s.diag), (d.system.upper, s.upper), (d.system.source, s.source), (d.psi0, psi0)):
            assert np.array_equal(x, y)
        assert len(a.interfaces) == len(b.interfaces) > 0
        for k, (ia, ib) in enumerate(zip(a.interfaces, b.interfaces)):
            assert ia.neighbProcNo == ib.neighbProcNo and np.array_equal(ia.faceCells, ib.faceCells)
            assert np.array_equal(d.system.bou[k], s.bou[k])
        assert all(int(x["offset"]) % 64 == 0 for x in d.header["arrays"])
        # C reader
        h = C.c_void_p()
        assert L.b200_dump_read(str(p).encode(), C.byref(h)) == 0, L.b200_dump_last_error()
        dd = L.b200_dump_get(h).contents
        assert (dd.nCells, dd.nFaces, dd.nIfaces, dd.rank, dd.nranks) == (b.nCells, b.nFaces, len(b.interfaces), r, 4)
        assert dd.controls.maxIter == 321 and dd.controls.minIter == 2 and dd.controls.precond == 1
        assert dd.perf.nIterations == 17 and dd.solverName == b"diagonalPCG" and dd.havePerf == 1
        got = np.ctypeslib.as_array(C.cast(dd.diag, C.POINTER(C.c_double)), shape=(b.nCells,))
        assert np.array_equal(got, s.diag)
        fc = np.ctypeslib.as_array(C.cast(dd.ifaces[0].faceCells, C.POINTER(C.c_int32)), shape=(dd.ifaces[0].nFaces,))
        assert np.array_equal(fc, b.interfaces[0].faceCells)
        assert json.loads(L.b200_dump_header_json(h).decode())["nCells"] == b.nCells
        L.b200_dump_free(h)


def test_edge_cases_and_errors(tmp_path):
    L = _lib.load_pcg()
    # empty system
    from firefoam_dev_b200.ldu import LduAddressing
    from firefoam_dev_b200.meshgen import System
    e = System(LduAddressing(0, [], []), np.zeros(0), np.zeros(0), np.zeros(0), [])
    p = tmp_path / "empty.b200sys"
    replay.write_dump(p, e, np.zeros(0), {"preconditioner": "none"})
    d = replay.read_dump(p)
    assert d.system.addr.nCells == 0 and d.reference is None
    # not a dump / truncated
    bad = tmp_path / "bad.b200sys"
    bad.write_bytes(b"FoamFile { version 2.0; }")
    with pytest.raises(ValueError):
        replay.read_dump(bad)
    h = C.c_void_p()
    assert L.b200_dump_read(str(bad).encode(), C.byref(h)) != 0 and not h
    assert b"not a b200 system dump" in L.b200_dump_last_error()
    good = open(os.path.join(GOLD, "singlebox_ph_rgh_c1.b200sys"), "rb").read()
    cut = tmp_path / "cut.b200sys"
    cut.write_bytes(good[:len(good) // 2])
    with pytest.raises(ValueError):
        replay.read_dump(cut)
    assert L.b200_dump_read(str(cut).encode(), C.byref(h)) != 0
    assert L.b200_dump_write(None, None) != 0
    assert L.b200_dump_read(str(tmp_path / "missing.b200sys").encode(), C.byref(h)) != 0


@pytest.mark.gpu
def test_replay_fixtures_on_gpu(ctx):
    for name in ("steckler_ph_rgh_c1.b200sys", "singlebox_ph_rgh_c1.b200sys", "steckler_G_p1.b200sys"):
        psi, perf, d = replay.replay(os.path.join(GOLD, name), context=ctx)
        assert perf.nIterations == d.reference["nIterations"]
        assert abs(perf.finalResidual - d.reference["finalResidual"]) <= 1e-9 * d.reference["finalResidual"]
        assert np.abs(psi - d.psi).max() <= 1e-12 * np.abs(d.psi).max()
    assert "DICB200PCG" in str(replay.replay(os.path.join(GOLD, "steckler_ph_rgh_c1.b200sys"), context=ctx)[1])


@pytest.mark.gpu
def test_replay_asymmetric_fixture_on_gpu():
    """the U-shaped fixture through B200smoothSolver: sweepMode exact (as dumped) reproduces the stored line and
    solution bit for bit; the default multicolour sweeps converge under the same controls"""
    from firefoam_dev_b200 import Context
    c = Context(device=0)
    try:
        path = os.path.join(GOLD, "steckler_U_transport.b200sys")
        psi, perf, d = replay.replay(path, context=c)
        assert perf.nIterations == d.reference["nIterations"] == 7 and np.array_equal(psi, d.psi)
        assert perf.finalResidual == pytest.approx(d.reference["finalResidual"], rel=1e-9)
        assert str(perf).startswith("B200smoothSolver:  Solving for Ux, Initial residual = 1, ")
        d.controls["B200"] = {"sweepMode": "multicolour"}
        psi2, perf2, _ = replay.replay(d, context=c)
        assert perf2.finalResidual < 1e-5 and np.abs(psi2 - d.psi).max() <= 1e-4 * np.abs(d.psi).max()
    finally:
        c.close()


@pytest.mark.gpu
def test_native_replay_tool_asymmetric():
    exe = os.path.join(ROOT, "firefoam-dev_b200", "b200replay")
    if not os.path.exists(exe):
        pytest.skip("b200replay not built")
    r = subprocess.run([exe, os.path.join(GOLD, "steckler_U_transport.b200sys")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "B200smoothSolver:  Solving for Ux, Initial residual = 1, " in r.stdout and "asymmetric" in r.stdout
    assert r.stdout.count("No Iterations 7") == 2 and "MISMATCH" not in r.stdout


@pytest.mark.gpu
def test_native_replay_tool():
    exe = os.path.join(ROOT, "firefoam-dev_b200", "b200replay")
    if not os.path.exists(exe):
        pytest.skip("b200replay not built")
    r = subprocess.run([exe, os.path.join(GOLD, "steckler_ph_rgh_c1.b200sys"),
                        os.path.join(GOLD, "singlebox_ph_rgh_c1.b200sys")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "DICB200PCG:  Solving for ph_rgh, Initial residual = 1, " in r.stdout
    assert r.stdout.count("No Iterations 29") == 2 and "MISMATCH" not in r.stdout


def test_dic_modes_survive_the_dump(tmp_path):
    """`B200 { dicMode exact | eisenstat; }` travels as the preconditioner code of the dumped controls."""
    s = mg.hex_block(5, 4, 3)
    L = _lib.load_pcg()
    for mode, code in (("multicolour", 2), ("exact", 3), ("eisenstat", 4)):
        p = tmp_path / f"p_rgh_{mode}.b200sys"
        ctl = {"preconditioner": "DIC", "tolerance": 1e-6, "relTol": 0.0, "maxIter": 1000, "minIter": 0}
        if mode != "multicolour":
            ctl["B200"] = {"dicMode": mode}
        replay.write_dump(p, s, np.zeros(s.addr.nCells), ctl)
        d = replay.read_dump(p)
        assert d.controls == ctl and d.header["controls"]["precondCode"] == code
        assert d.header["controls"]["preconditioner"] == "DIC"
        h = C.c_void_p()
        assert L.b200_dump_read(str(p).encode(), C.byref(h)) == 0, L.b200_dump_last_error()
        assert L.b200_dump_get(h).contents.controls.precond == code
        L.b200_dump_free(h)


def test_native_reader_rejects_hostile_headers(tmp_path):
    """The native reader (b200replay feeds it files from disk) must not trust header numbers: wrapping header
    length / offsets, negative or fractional counts, and key names that appear as string VALUES."""
    import struct
    L = _lib.load_pcg()
    good = open(os.path.join(GOLD, "singlebox_ph_rgh_c1.b200sys"), "rb").read()
    hlen = struct.unpack("<Q", good[8:16])[0]
    hdr = good[16:16 + hlen].decode()
    h = C.c_void_p()

    def read(blob, name):
        p = tmp_path / name
        p.write_bytes(blob)
        rc = L.b200_dump_read(str(p).encode(), C.byref(h))
        if rc == 0:
            L.b200_dump_free(h)
        return rc, L.b200_dump_last_error().decode()

    def with_header(new):
        assert len(new) == len(hdr)          # same length: the array offsets stay valid
        return good[:16] + new.encode() + good[16 + hlen:]

    assert read(good, "ok.b200sys")[0] == 0
    # header length that wraps 16 + hlen around 2^64
    rc, msg = read(good[:8] + struct.pack("<Q", 2**64 - 8) + good[16:], "wrap.b200sys")
    assert rc != 0 and "truncated header" in msg
    # negative / fractional / huge sizes
    for a, b in (('"nCells": 245', '"nCells": -45'), ('"nCells": 245', '"nCells": 2.5'), ('"nFaces": 616', '"nFaces": 9e9')):
        assert a in hdr
        rc, msg = read(with_header(hdr.replace(a, b, 1)), "neg.b200sys")
        assert rc != 0 and "non-negative integers" in msg, (b, msg)
    # an array offset near 2^64 (off + bytes wraps) and a count far beyond the file
    i = hdr.index('"offset": "') + len('"offset": "')
    rc, msg = read(with_header(hdr[:i] + "18446744073709551608" + hdr[i + 20:]), "off.b200sys")
    assert rc != 0 and "out of bounds" in msg
    a = '"count": 616'
    rc, msg = read(with_header(hdr.replace(a, '"count": 9e9', 1)), "cnt.b200sys")
    assert rc != 0 and ("bad array count" in msg or "mandatory" in msg)
    # a field called like a key: "fieldName": "nCells" must not be parsed as the nCells entry
    assert '"fieldName": "ph_rgh"' in hdr
    rc, msg = read(with_header(hdr.replace('"fieldName": "ph_rgh"', '"fieldName": "nCells"', 1)), "key.b200sys")
    assert rc == 0, msg
    assert L.b200_dump_read(str(tmp_path / "key.b200sys").encode(), C.byref(h)) == 0
    dd = L.b200_dump_get(h).contents
    assert dd.nCells == 245 and dd.fieldName == b"nCells"
    L.b200_dump_free(h)
