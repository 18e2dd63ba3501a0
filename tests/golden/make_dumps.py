"""Generates the committed `.b200sys` fixtures (format regression corpus) with the
CPU oracle as the dumping solver:

  singlebox_ph_rgh_c1.b200sys   first hydrostatic corrector of the restated cases/singleBox base block
                                (245 cells, PCG + diagonal; BASELINE config 1)
  steckler_ph_rgh_c1.b200sys    first hydrostatic corrector of cases/steckler (9 000 cells, DICPCG,
                                tol 1e-6 relTol 0.01: the system behind log.fireFoam:92, 29 iterations)

  steckler_G_p1.b200sys         the P1 radiation model's G equation restated on the steckler topology (9 000
                                cells, DICPCG, tol 1e-6 relTol 0: SURVEY.md 8f-4, the reference's other symmetric solve)

  steckler_U_transport.b200sys  a U-shaped ASYMMETRIC transport system (ddt + div + laplacian, cases.transport_system)
                                on the steckler topology with the reference's U controls (`smoothSolver`,
                                `symGaussSeidel`, tol 1e-6, relTol 0, maxIter 10; fvSolution:48-55): SURVEY.md 8f-4,
                                the dumped reference is the oracle's smoothSolver (sweepMode exact reproduces it)

The systems come from firefoam-dev_b200/cases.py (the case files restated by hand: no OpenFOAM here),
the recorded reference lines from oracle/ (plain-C restatement of OpenFOAM's PCG).  Run from the repo
root: python tests/golden/make_dumps.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from firefoam_dev_b200 import replay  # noqa: E402
from firefoam_dev_b200.cases import SingleBoxHydrostatic, StecklerHydrostatic  # noqa: E402
from firefoam_dev_b200.meshgen import System  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def first_corrector(case, pre, name):
    m, src = case.assemble(lambda g, s, d, sign, d0: orc.laplacian_assemble(
        case.addr.lowerAddr, case.addr.upperAddr, case.N, g, s, d, sign, d0))
    sysm = System(case.addr, m.diag, m.upper, src, [])
    psi0 = case.ph_rgh.copy()
    psi = psi0.copy()
    perf = orc.pcg_solve(sysm, psi, pre, case.TOL, case.RELTOL, 1000)
    ctl = {"preconditioner": pre, "tolerance": case.TOL, "relTol": case.RELTOL, "maxIter": 1000}
    if pre == "DIC":
        ctl["B200"] = {"dicMode": "exact"}     # the dumped reference is OpenFOAM's DIC, not the multicolour IC0
    ref = {"initialResidual": perf.initialResidual, "finalResidual": perf.finalResidual,
           "nIterations": perf.nIterations, "converged": perf.converged, "singular": perf.singular}
    path = os.path.join(HERE, name)
    replay.write_dump(path, sysm, psi0, ctl, fieldName="ph_rgh", psi=psi, reference=ref,
                      solverName=pre + "PCG", solveIndex=0, time=0.0)
    print(name, os.path.getsize(path), "bytes;", pre + "PCG", perf.nIterations, "iterations")


def g_equation(name):
    """P1 G equation on the steckler topology (cases.p1_G_terms; P1.C:238-244, fvSolution:75-81: PCG + DIC,
    tol 1e-6, relTol 0), assembled by the oracle's fvMatrix algebra."""
    from firefoam_dev_b200.cases import p1_G_terms
    case, t = p1_G_terms()
    up, dg, src = orc.assemble_p_rgh(case.addr.lowerAddr, case.addr.upperAddr, case.N, t)
    sysm = System(case.addr, dg, up, src, [])
    psi0 = np.zeros(case.N)
    psi = psi0.copy()
    perf = orc.pcg_solve(sysm, psi, "DIC", 1e-6, 0.0, 1000)
    ctl = {"preconditioner": "DIC", "tolerance": 1e-6, "relTol": 0.0, "maxIter": 1000, "B200": {"dicMode": "exact"}}
    ref = {"initialResidual": perf.initialResidual, "finalResidual": perf.finalResidual,
           "nIterations": perf.nIterations, "converged": perf.converged, "singular": perf.singular}
    path = os.path.join(HERE, name)
    replay.write_dump(path, sysm, psi0, ctl, fieldName="G", psi=psi, reference=ref, solverName="DICPCG",
                      solveIndex=0, time=0.0)
    print(name, os.path.getsize(path), "bytes; DICPCG", perf.nIterations, "iterations")


def u_transport(name):
    """cases/steckler/system/fvSolution:48-55 on a synthetic U-shaped matrix (the reference ships none)."""
    from firefoam_dev_b200.cases import transport_system
    case = StecklerHydrostatic()
    m, src = case.assemble(lambda g, s, d, sign, d0: orc.laplacian_assemble(
        case.addr.lowerAddr, case.addr.upperAddr, case.N, g, s, d, sign, d0))
    t = transport_system(System(case.addr, m.diag, m.upper, src, [], None), seed=9, kappa=0.6)
    psi0 = np.zeros(case.N)
    psi = psi0.copy()
    ctl = {"smoother": "symGaussSeidel", "tolerance": 1e-6, "relTol": 0.0, "maxIter": 10, "minIter": 0, "nSweeps": 1,
           "B200": {"sweepMode": "exact"}}
    perf = orc.smooth_solve(t, psi, smoother="symGaussSeidel", tolerance=1e-6, relTol=0.0, maxIter=10)
    ref = {"initialResidual": perf.initialResidual, "finalResidual": perf.finalResidual,
           "nIterations": perf.nIterations, "converged": perf.converged, "singular": 0}
    path = os.path.join(HERE, name)
    replay.write_dump(path, t, psi0, ctl, fieldName="Ux", psi=psi, reference=ref, solverName="smoothSolver",
                      solveIndex=0, time=0.0)
    print(name, os.path.getsize(path), "bytes; smoothSolver", perf.nIterations, "sweeps, final", perf.finalResidual)


first_corrector(SingleBoxHydrostatic(), "diagonal", "singlebox_ph_rgh_c1.b200sys")
first_corrector(StecklerHydrostatic(), "DIC", "steckler_ph_rgh_c1.b200sys")
g_equation("steckler_G_p1.b200sys")
u_transport("steckler_U_transport.b200sys")
