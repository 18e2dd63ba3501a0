"""Pins the CPU oracle against the reference's only golden artefact for this path:
cases/steckler/original/linux64/log.fireFoam:92-100 (tests/golden/steckler_log.json), following
the recipe of SURVEY.md Appendix B with ONE correction found in round 2
(tools/kat/steckler_kat_candidates.py): the density on the fixed-value `top` patch is that of the
patch-face mixture as read from 0/ (N2 `calculated; value uniform 0` -> O2 only).  With it the KAT is
DIGIT-level: all five DICPCG lines (counts 29, 32, 7, 0, 0; initial and final residuals) and the three
printed hydrostatic variations reproduce to the printed digits (last-digit rounding of quantities that
sit on the fp64 cancellation noise of p = ph_rgh + rho*gh + pRef)."""
import json
import os

import numpy as np
import pytest

from firefoam_dev_b200.cases import StecklerHydrostatic
from firefoam_dev_b200.meshgen import System
from oracle import oracle as orc
from helpers import hydrostatic_loop, STECKLER_DIAG_COUNTS as DIAG_COUNTS

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "steckler_log.json")))


def run(pre):
    case = StecklerHydrostatic()
    a = case.addr
    lap = lambda g, s, d, sign, d0: orc.laplacian_assemble(a.lowerAddr, a.upperAddr, a.nCells, g, s, d, sign, d0)
    solve = lambda m, b, psi: orc.pcg_solve(System(a, m.diag, m.upper, b), psi, pre, case.TOL, case.RELTOL)
    return case, hydrostatic_loop(case, lap, solve)


def test_mesh_matches_appendix_b():
    case = StecklerHydrostatic()
    assert (case.N, case.F) == (9000, 24868)       # 25 650 box faces - 782 baffle faces
    l, u = case.addr.lowerAddr, case.addr.upperAddr
    assert np.all(l < u) and np.all(np.diff(l) >= 0)


def test_dic_lines_match_golden_log_to_the_printed_digits():
    _, res = run("DIC")
    gold = GOLD["ph_rgh"]
    assert [g["solver"] for g in gold] == ["DICPCG"] * 5
    # exact iteration counts of all five correctors (log.fireFoam:92,94,96,98,100)
    assert [r[2] for r in res] == [g["iters"] for g in gold] == [29, 32, 7, 0, 0]
    assert res[0][0] == pytest.approx(1.0, abs=1e-14) and gold[0]["initial"] == 1.0
    # printed initial / final residuals: 8 significant digits in the log.  Correctors 1-3 agree to <= 4e-8;
    # the residuals of correctors 4 and 5 (1e-6 of the norm factor, dominated by the cancellation noise of
    # rho differences) to 4e-7.
    for k in range(5):
        tol = 1e-7 if k < 3 else 1e-6
        assert res[k][0] == pytest.approx(gold[k]["initial"], rel=tol), k
        assert res[k][1] == pytest.approx(gold[k]["final"], rel=tol), k


def test_hydrostatic_functional_matches_golden_log():
    _, res = run("DIC")
    # gMax-gMin(ph_rgh) after every corrector (log.fireFoam:93,95,97,99,101): 2e-8 relative = one unit of
    # the last printed digit
    for k in range(5):
        assert res[k][3] == pytest.approx(GOLD["variation"][k]["value"], rel=5e-8), k


def test_diagonal_counts_self_derived():
    """PCG+diagonal is run by no shipped case: UNPINNED by the reference.  These counts are the
    oracle's own and guard against regressions only."""
    _, res = run("diagonal")
    assert [r[2] for r in res] == DIAG_COUNTS
    assert res[4][3] == pytest.approx(GOLD["variation"][4]["value"], rel=1e-5)


def test_oracle_controls_semantics():
    """maxIter+1 loop bodies (nIterations++ < maxIter), minIter, relTol."""
    from firefoam_dev_b200 import meshgen as mg
    s = mg.hex_block(8, 6, 5)
    psi = np.zeros(s.addr.nCells)
    p = orc.pcg_solve(s, psi, "diagonal", 1e-30, 0.0, maxIter=3)
    assert p.nIterations == 4 and not p.converged
    psi[:] = 0
    p = orc.pcg_solve(s, psi, "diagonal", 1e-6, 0.5, maxIter=100)
    assert p.converged and p.finalResidual < 0.5 * p.initialResidual
    # converged initial guess: 0 iterations unless minIter
    psi = s.xstar * (1 + 1e-9)
    p = orc.pcg_solve(s, psi, "diagonal", 1e-6, 0.0)
    assert p.nIterations == 0 and p.converged
    p = orc.pcg_solve(s, psi, "diagonal", 1e-6, 0.0, minIter=2)
    assert p.nIterations == 2
    # exact solution: r == 0 -> wApA == 0 -> singular break inside the first loop body
    psi = s.xstar.copy()
    p = orc.pcg_solve(s, psi, "diagonal", 1e-6, 0.0, minIter=2)
    assert p.singular and p.nIterations == 0


def test_singlebox_config1_diagonal_plumbing():
    """BASELINE config 1: singleBox base block (7 x 5 x 7), PCG + diagonal, 1 rank, oracle only.
    Nothing in the reference pins these numbers (self-derived regression values)."""
    from firefoam_dev_b200.cases import SingleBoxHydrostatic
    case = SingleBoxHydrostatic()
    assert (case.N, case.F) == (245, 616)
    a = case.addr
    lap = lambda g, s, d, sign, d0: orc.laplacian_assemble(a.lowerAddr, a.upperAddr, a.nCells, g, s, d, sign, d0)
    solve = lambda m, b, psi: orc.pcg_solve(System(a, m.diag, m.upper, b), psi, "diagonal", case.TOL, case.RELTOL)
    res = hydrostatic_loop(case, lap, solve)
    assert [r[2] for r in res] == [12, 8, 9, 0, 0]
    # hydrostatic variation ~ rho0 * g * psi*... small and positive; converged value stable
    assert res[4][3] == pytest.approx(0.0021817037, rel=1e-6)


def test_steckler_p_rgh_synthetic_config2():
    """BASELINE config 2 (replay stand-in): p_rgh-shaped systems on the steckler topology."""
    from firefoam_dev_b200.cases import steckler_p_rgh_system
    s = steckler_p_rgh_system()
    a = s.addr
    up, dg = orc.laplacian_assemble(a.lowerAddr, a.upperAddr, a.nCells, s.gamma_f, s.magSf, s.deltaCoeffs, -1.0, s.diag0)
    assert np.array_equal(up, s.upper) and np.array_equal(dg, s.diag)
    its = {}
    for pre in ("diagonal", "DIC"):
        for rt in (0.01, 0.0):
            psi = np.zeros(a.nCells)
            p = orc.pcg_solve(s, psi, pre, 1e-6, rt)
            assert p.converged
            its[(pre, rt)] = p.nIterations
    assert its[("DIC", 0.01)] < its[("diagonal", 0.01)] < its[("diagonal", 0.0)]
    # same order of magnitude as the reference's real p_rgh solves (9-21 / 22-28 DICPCG iterations,
    # cases/steckler/original/linux64/log.fireFoam:220-1231)
    assert 3 <= its[("DIC", 0.01)] <= 30 and 15 <= its[("DIC", 0.0)] <= 60
