"""Whole-p_rghEqn assembly: fvm::ddt + explicit fvc terms + fvc::div + laplacian +
explicit sources + the boundary fold of solveSegregated (solver/pEqn.H:26-37, solver/phrghEqn.H:43-46).
CPU: the oracle's operator-by-operator restatement against an independent numpy formulation and
against the golden log; GPU: the single-pass sorted-segment kernel against the oracle, bit for bit."""
import json
import os

import numpy as np
import pytest

from firefoam_dev_b200 import meshgen as mg
from firefoam_dev_b200.cases import StecklerHydrostatic
from firefoam_dev_b200.meshgen import System
from oracle import oracle as orc
from helpers import random_ldu
from conftest import ROOT


def random_terms(s, seed, nExplicit=2, with_ddt=True, with_div=True, with_su=True, with_boundary=True,
                 lapSign=-1.0, divSign=-1.0):
    a = s.addr
    N, F = a.nCells, a.nFaces
    rng = np.random.default_rng(seed)
    t = {"rDeltaT": 1.0 / 0.0037, "V": rng.uniform(0.5, 2.0, N) * 1e-3,
         "gamma_f": rng.uniform(0.1, 1.0, F), "magSf": rng.uniform(0.5, 1.5, F) * 1e-2,
         "deltaCoeffs": rng.uniform(50, 150, F), "lapSign": lapSign, "divSign": divSign}
    if with_ddt:
        t.update(psi=rng.uniform(1.0, 1.3, N) * 1e-5, psi0=rng.uniform(1.0, 1.3, N) * 1e-5,
                 p0=rng.standard_normal(N) * 10.0)
    t["explicit"] = [rng.standard_normal(N) for _ in range(nExplicit)]
    if with_div:
        t["phi"] = rng.standard_normal(F) * 1e-3
    if with_su:
        t["Su"] = rng.standard_normal(N)
    if with_boundary:
        nB = max(1, N // 3)
        cells = rng.integers(0, N, size=nB).astype(np.int32)   # several faces per cell, unsorted (patch order)
        cells[: min(5, nB)] = cells[0]
        t.update(bCells=cells, bPhi=rng.standard_normal(nB) * 1e-3, bInternal=rng.uniform(0, 2, nB),
                 bBoundary=rng.standard_normal(nB))
        t["bBoundary"][::4] = 0.0                                # coupled patches: boundaryCoeffs not folded
    return t


def numpy_assembly(a, t):
    """Independent (vectorised, different summation order) statement of the same equations."""
    N = a.nCells
    l, u = a.lowerAddr, a.upperAddr
    V = t["V"]
    U = t["deltaCoeffs"] * (t["gamma_f"] * t["magSf"])
    lapdiag = -(np.bincount(l, U, N) + np.bincount(u, U, N))
    diag = (t["rDeltaT"] * t["psi"] * V if "psi" in t else np.zeros(N)) + t["lapSign"] * lapdiag
    src = t["rDeltaT"] * t["psi0"] * t["p0"] * V if "psi" in t else np.zeros(N)
    for e in t["explicit"]:
        src = src - V * e
    if "phi" in t:
        div = np.bincount(l, t["phi"], N) - np.bincount(u, t["phi"], N)
        if "bPhi" in t:
            div = div + np.bincount(t["bCells"], t["bPhi"], N)
        src = src + t["divSign"] * div
    if "Su" in t:
        src = src + V * t["Su"]
    if "bInternal" in t:
        diag = diag + np.bincount(t["bCells"], t["bInternal"], N)
        src = src + np.bincount(t["bCells"], t["bBoundary"], N)
    return t["lapSign"] * U, diag, src


@pytest.mark.parametrize("lapSign,divSign", [(-1.0, -1.0), (1.0, 1.0)])
def test_oracle_restatement_matches_independent_formulation(lapSign, divSign):
    for s in (mg.hex_block(9, 7, 5), random_ldu(700, 5.0, seed=2), mg.bcc_poly(4, 4, 5)):
        t = random_terms(s, 3, lapSign=lapSign, divSign=divSign)
        up, dg, src = orc.assemble_p_rgh(s.addr.lowerAddr, s.addr.upperAddr, s.addr.nCells, t)
        up2, dg2, src2 = numpy_assembly(s.addr, t)
        assert np.array_equal(up, up2)
        np.testing.assert_allclose(dg, dg2, rtol=1e-13, atol=1e-15 * np.abs(dg2).max())
        np.testing.assert_allclose(src, src2, rtol=1e-11, atol=1e-13 * np.abs(src2).max())


def test_oracle_reduces_to_the_laplacian_assembly():
    s = mg.hex_block(8, 6, 5)
    a = s.addr
    t = {"V": np.ones(a.nCells), "gamma_f": s.gamma_f, "magSf": s.magSf, "deltaCoeffs": s.deltaCoeffs,
         "lapSign": -1.0, "explicit": []}
    up, dg, src = orc.assemble_p_rgh(a.lowerAddr, a.upperAddr, a.nCells, t)
    up_ref, dg_ref = orc.laplacian_assemble(a.lowerAddr, a.upperAddr, a.nCells, s.gamma_f, s.magSf, s.deltaCoeffs, -1.0)
    assert np.array_equal(up, up_ref) and np.array_equal(dg, dg_ref) and not src.any()


def hydrostatic_terms(case):
    """ph_rghEqn of solver/phrghEqn.H:43-46: fvm::laplacian(rhof, ph_rgh) == fvc::div(phig), top patch
    fixedValue 0 (internalCoeffs = -rho_b*magSf*deltaCoeffs_b), the fixedFluxPressure patches cancel."""
    a = case.addr
    l, u = a.lowerAddr, a.upperAddr
    rhof = 0.5 * (case.rho[l] + case.rho[u])
    phig = -rhof * case.ghf * ((case.rho[u] - case.rho[l]) * case.deltaCoeffs) * case.magSf
    dx, dy, dz = case.d
    nB = case.top.size
    return {"V": np.full(case.N, dx * dy * dz), "gamma_f": rhof, "magSf": case.magSf,
            "deltaCoeffs": case.deltaCoeffs, "lapSign": 1.0, "phi": phig, "divSign": 1.0, "explicit": [],
            "bCells": case.top.astype(np.int32), "bPhi": np.zeros(nB),
            "bInternal": np.full(nB, -case.rho_top * (dx * dz) * (2.0 / dy)), "bBoundary": np.zeros(nB)}


def test_ph_rghEqn_through_the_full_assembly_reproduces_the_golden_log_counts():
    """The steckler hydrostatic loop (log.fireFoam:92-96) with the equation assembled by the p_rghEqn
    restatement (fvc::div with its /V, *V round trip, boundary fold) instead of cases.py's shortcut."""
    log = json.load(open(os.path.join(ROOT, "tests", "golden", "steckler_log.json")))
    case = StecklerHydrostatic()
    psi = case.ph_rgh.copy()
    iters = []
    for k in range(5):
        up, dg, src = orc.assemble_p_rgh(case.addr.lowerAddr, case.addr.upperAddr, case.N, hydrostatic_terms(case))
        perf = orc.pcg_solve(System(case.addr, dg, up, src, []), psi, "DIC", case.TOL, case.RELTOL, 1000)
        var = case.update(psi)
        iters.append(perf.nIterations)
        assert perf.finalResidual == pytest.approx(log["ph_rgh"][k]["final"], rel=1e-7 if k < 3 else 1e-6)
        assert var == pytest.approx(log["variation"][k]["value"], rel=5e-8)
    assert iters == [e["iters"] for e in log["ph_rgh"]] == [29, 32, 7, 0, 0]


@pytest.mark.gpu
@pytest.mark.parametrize("renumber", ["0", "1"])
def test_gpu_assembly_bit_exact(renumber):
    from test_gpu_parity import _ctx_with_env
    import torch
    c = _ctx_with_env({"B200PCG_RENUMBER": renumber})
    try:
        for i, s in enumerate((mg.hex_block(24, 20, 16), random_ldu(5001, 6.0, seed=7), mg.bcc_poly(9, 8, 10),
                               mg.hex_block(3, 2, 1))):
            a = s.addr
            c.set_addressing(a)
            for kw in (dict(), dict(lapSign=1.0, divSign=1.0, with_ddt=False, nExplicit=0, with_su=False),
                       dict(with_boundary=False, with_div=False), dict(nExplicit=8)):
                t = random_terms(s, 10 + i, **kw)
                ref = orc.assemble_p_rgh(a.lowerAddr, a.upperAddr, a.nCells, t)
                got = c.assemble_p_rgh(t)
                for x, y, nm in zip(got, ref, ("upper", "diag", "source")):
                    assert np.array_equal(x, y), (nm, kw)
                # device entry point
                dev = {k: (torch.from_numpy(np.ascontiguousarray(v, dtype=np.float64)).cuda()
                           if isinstance(v, np.ndarray) and k != "bCells" else v) for k, v in t.items()}
                dev["explicit"] = [torch.from_numpy(e).cuda() for e in t["explicit"]]
                outs = [torch.empty(n, dtype=torch.float64, device="cuda") for n in (a.nFaces, a.nCells, a.nCells)]
                c.assemble_p_rgh_device(dev, *outs)
                for x, y in zip(outs, ref):
                    assert np.array_equal(x.cpu().numpy(), y)
        # errors
        from firefoam_dev_b200 import B200Error
        t = random_terms(s, 1, nExplicit=9)
        with pytest.raises(B200Error):
            c.assemble_p_rgh(t)
    finally:
        c.close()


@pytest.mark.gpu
def test_gpu_hydrostatic_loop_with_device_assembly(ctx):
    """ph_rghEqn assembled by the CUDA kernel and solved by B200PCG (DIC-exact): the golden log's counts."""
    from firefoam_dev_b200 import B200PCG, LduMatrix
    case = StecklerHydrostatic()
    ctx.set_addressing(case.addr)
    psi = case.ph_rgh.copy()
    iters = []
    log = json.load(open(os.path.join(ROOT, "tests", "golden", "steckler_log.json")))
    for k in range(5):
        up, dg, src = ctx.assemble_p_rgh(hydrostatic_terms(case))
        ctl = {"preconditioner": "DIC", "tolerance": case.TOL, "relTol": case.RELTOL, "B200": {"dicMode": "exact"}}
        perf = B200PCG("ph_rgh", LduMatrix(case.addr, dg, up), [], None, [], ctl, context=ctx).solve(psi, src)
        var = case.update(psi)
        iters.append(perf.nIterations)
        assert perf.finalResidual == pytest.approx(log["ph_rgh"][k]["final"], rel=1e-7 if k < 3 else 1e-6)
        assert var == pytest.approx(log["variation"][k]["value"], rel=5e-8)
    assert iters == [29, 32, 7, 0, 0]
