"""GPU parity of the PBiCG + DILU path (what the reference's other cases select for their transport
equations -- cases/wallFireSpread2D/system/fvSolution:66-73 -- and what produced its 2.4.x steckler logs) through the
C ABI (b200_bicg_solve) against oracle/bicg_oracle.c on the same seeded inputs.

Bars: DILU exact (level-scheduled) / diagonal: identical iteration counts, solution within 1e-10 relative (only the
order of the global dot-product sums differs from the CPU); `none`: counts within 3, solution 1e-5 (un-preconditioned
BiCG amplifies the summation order, like `none` PCG); DILU-class (multicolour): same solution within 1e-8 at a tight
tolerance; on a SYMMETRIC matrix DILU exact reproduces the digit-pinned DICPCG line of log.fireFoam:92."""
import os

import numpy as np
import pytest

import helpers
from conftest import ROOT
from firefoam_dev_b200 import cases, meshgen
from firefoam_dev_b200.meshgen import System
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def systems():
    yield "random", cases.transport_system(helpers.random_ldu(3000, 6, 17), seed=5, kappa=0.05)
    yield "hex", cases.transport_system(meshgen.hex_block(23, 17, 19), seed=6, kappa=0.05)
    yield "poly", cases.transport_system(meshgen.bcc_poly(9, 8, 7), seed=7, kappa=0.1)
    b = helpers.random_ldu(2500, 5, 23)
    yield "symmetric", System(b.addr, b.diag, b.upper, b.source, [], b.xstar)


SYSTEMS = list(systems())
IDS = [n for n, _ in SYSTEMS]


@pytest.fixture(scope="module")
def gctx():
    from firefoam_dev_b200 import Context
    c = Context(device=0)
    yield c
    c.close()


def solver(gctx, s, **ctl):
    from firefoam_dev_b200 import B200PBiCG
    return B200PBiCG("Yi", s.matrix, [], None, [], ctl, context=gctx)


@pytest.mark.parametrize("name,s", SYSTEMS, ids=IDS)
@pytest.mark.parametrize("pre", ["DILU-exact", "diagonal", "none"])
def test_pbicg_matches_oracle(gctx, name, s, pre):
    N = s.addr.nCells
    opre = "DILU" if pre == "DILU-exact" else pre
    extra = {"B200": {"diluMode": "exact"}} if pre == "DILU-exact" else {}
    for ctl in (dict(tolerance=1e-8, relTol=0.0, maxIter=1000),          # wallFireSpread2D/system/fvSolution:66-73
                dict(tolerance=1e-30, relTol=0.0, maxIter=5), dict(tolerance=1e-3, minIter=3, maxIter=1000)):
        ref = np.zeros(N)
        pr = orc.pbicg_solve(s, ref, opre, **ctl)
        psi = np.zeros(N)
        perf = solver(gctx, s, preconditioner=opre, **extra, **ctl).solve(psi, s.source)
        loose = pre == "none"
        assert abs(perf.nIterations - pr.nIterations) <= (3 if loose else 0), (ctl, perf.nIterations, pr.nIterations)
        assert perf.initialResidual == pytest.approx(pr.initialResidual, rel=1e-12)
        assert np.abs(psi - ref).max() <= (1e-5 if loose else 1e-10) * np.abs(ref).max()
        if not loose:
            assert perf.finalResidual == pytest.approx(pr.finalResidual, rel=1e-6)
            assert perf.converged == bool(pr.converged)
    assert str(perf).startswith({"DILU-exact": "DILUB200PBiCG", "diagonal": "diagonalB200PBiCG", "none": "noneB200PBiCG"}[pre] +
                                ":  Solving for Yi, Initial residual = ")


@pytest.mark.parametrize("name,s", SYSTEMS, ids=IDS)
def test_dilu_class_solution_parity_and_transliteration(gctx, name, s):
    N = s.addr.nCells
    ref = np.zeros(N)
    pr = orc.pbicg_solve(s, ref, "DILU", tolerance=1e-12, maxIter=3000)
    psi = np.zeros(N)
    perf = solver(gctx, s, preconditioner="DILU", tolerance=1e-12, maxIter=3000).solve(psi, s.source)
    assert pr.finalResidual < 1e-12 and perf.converged and perf.finalResidual < 1e-12
    assert np.linalg.norm(psi - ref) <= 1e-8 * np.linalg.norm(ref)
    assert str(perf).startswith("DILU(mc)B200PBiCG:  Solving for Yi")
    # the numpy transliteration of the same kernels on the same (multicolour) plan
    pv = helpers.PlanView(1, s.addr, renumber=-1)
    got, n, init, final = helpers.pbicg_emulated(pv, s, np.zeros(N), precond="DILU", tol=1e-8, maxIter=1000)
    psi = np.zeros(N)
    perf = solver(gctx, s, preconditioner="DILU", tolerance=1e-8, maxIter=1000).solve(psi, s.source)
    assert perf.nIterations == n and np.abs(psi - got).max() <= 1e-10 * np.abs(got).max()


def test_symmetric_steckler_system_reproduces_the_pinned_dicpcg_line(gctx):
    """DILU of a symmetric matrix is DIC, BiCG is CG: the system behind log.fireFoam:92 through B200PBiCG"""
    from firefoam_dev_b200 import replay
    d = replay.read_dump(os.path.join(ROOT, "tests", "golden", "steckler_ph_rgh_c1.b200sys"))
    psi = d.psi0.copy()
    perf = solver(gctx, d.system, preconditioner="DILU", tolerance=d.controls["tolerance"], relTol=d.controls["relTol"],
                  maxIter=d.controls["maxIter"], B200={"diluMode": "exact"}).solve(psi, d.system.source)
    assert perf.nIterations == 29
    assert perf.finalResidual == pytest.approx(d.reference["finalResidual"], rel=1e-8)
    assert np.abs(psi - d.psi).max() <= 1e-10 * np.abs(d.psi).max()


def test_zero_system_line_device_entry_and_errors(gctx):
    import torch
    from firefoam_dev_b200 import B200Error
    from firefoam_dev_b200.ldu import make_bicg_controls
    _, s = SYSTEMS[1]
    N = s.addr.nCells
    z = System(s.addr, s.diag, s.upper, np.zeros(N), [], None, lower=s.lower)
    psi = np.zeros(N)
    perf = solver(gctx, z, preconditioner="DILU", tolerance=1e-8).solve(psi, z.source)
    # cases/steckler/original/darwinIntel64/log.fireFoam:161
    assert str(perf).endswith("Solving for Yi, Initial residual = 0, Final residual = 0, No Iterations 0") and not psi.any()
    with pytest.raises(ValueError):
        solver(gctx, s, preconditioner="DIC")
    # device entry point, bit-reproducible re-runs
    gctx.set_addressing(s.addr)
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    d, up, lo, b = t(s.diag), t(s.upper), t(s.lower), t(s.source)
    ctl, _ = make_bicg_controls(dict(preconditioner="DILU", tolerance=1e-9, maxIter=500))
    outs = []
    for _ in range(2):
        x = torch.zeros(N, dtype=torch.float64, device=dev)
        p = gctx.bicg_solve_device(d, up, lo, b, x, ctl)
        torch.cuda.synchronize()
        outs.append((x.cpu().numpy(), p.nIterations, p.finalResidual))
    assert outs[0][1] == outs[1][1] and outs[0][2] == outs[1][2] and np.array_equal(outs[0][0], outs[1][0])
    psi = np.zeros(N)
    ph = solver(gctx, s, preconditioner="DILU", tolerance=1e-9, maxIter=500).solve(psi, s.source)
    assert ph.nIterations == outs[0][1] and np.array_equal(psi, outs[0][0])
    # the PCG and smoothSolver paths still work on the same context afterwards (shared buffers: t, dT, eD, valT)
    from firefoam_dev_b200 import B200PCG, B200smoothSolver, LduMatrix
    g = meshgen.hex_block(23, 17, 19)
    for pre in ("DIC", "diagonal"):
        x1, x2 = np.zeros(N), np.zeros(N)
        B200PCG("p", LduMatrix(g.addr, g.diag, g.upper), [], None, [], dict(preconditioner=pre, tolerance=1e-8), context=gctx).solve(x1, g.source)
        assert np.abs(x1 - g.xstar).max() < 1e-3 * np.abs(g.xstar).max()
    x3 = np.zeros(N)
    p3 = B200smoothSolver("U", s.matrix, [], None, [], dict(smoother="symGaussSeidel", tolerance=1e-9, maxIter=500), context=gctx).solve(x3, s.source)
    assert p3.converged
