"""CPU tests of the data structure + operation order the smoothSolver kernels use (numpy transliteration in
tests/helpers.py: PlanView.gs_rows / gs_residual / smooth_solve_emulated) against oracle/smooth_oracle.c.

Exact mode (Levels plan: rows ordered by the dependency level of OpenFOAM's own cell-order recurrence, entries
[lower neighbours | upper neighbours] each in ascending face order) must reproduce the sequential sweeps of
GaussSeidelSmoother.C / symGaussSeidelSmoother.C BIT FOR BIT; the multicolour mode is a different ordering of the
unknowns (GS-class): it must converge to the same solution."""
import numpy as np
import pytest

import helpers
from firefoam_dev_b200 import cases, meshgen
from oracle import oracle as orc

LEVELS, MULTICOLOUR = 2, 1


def systems():
    yield "random", cases.transport_system(helpers.random_ldu(120, 6, 17), seed=5)
    yield "hex", cases.transport_system(meshgen.hex_block(7, 5, 6), seed=6)
    yield "poly", cases.transport_system(meshgen.bcc_poly(4, 3, 3), seed=7, kappa=0.3)
    b = helpers.random_ldu(90, 5, 23)
    yield "symmetric", meshgen.System(b.addr, b.diag, b.upper, b.source, [], b.xstar)


@pytest.mark.parametrize("name,s", list(systems()), ids=[n for n, _ in systems()])
@pytest.mark.parametrize("smoother", ["GaussSeidel", "symGaussSeidel"])
def test_level_scheduled_sweeps_are_bit_identical(name, s, smoother):
    pv = helpers.PlanView(LEVELS, s.addr)
    N = s.addr.nCells
    for nS in (1, 2):
        psi = np.zeros(N)
        orc.smooth_solve(s, psi, smoother=smoother, nSweeps=-nS)
        got, n, _, _ = helpers.smooth_solve_emulated(pv, s, np.zeros(N), smoother=smoother, nSweeps=-nS, mode="exact")
        assert n == nS and np.array_equal(got, psi)
    psi = np.zeros(N)
    p = orc.smooth_solve(s, psi, smoother=smoother, tolerance=1e-7, maxIter=300)
    got, n, init, final = helpers.smooth_solve_emulated(pv, s, np.zeros(N), smoother=smoother, tol=1e-7, maxIter=300,
                                                          mode="exact")
    assert n == p.nIterations and np.array_equal(got, psi)
    assert init == pytest.approx(p.initialResidual, rel=1e-12) and final == pytest.approx(p.finalResidual, rel=1e-9)


@pytest.mark.parametrize("name,s", list(systems()), ids=[n for n, _ in systems()])
def test_multicolour_sweeps_converge_to_the_same_solution(name, s):
    pv = helpers.PlanView(MULTICOLOUR, s.addr)
    N = s.addr.nCells
    psi = np.zeros(N)
    p = orc.smooth_solve(s, psi, tolerance=1e-12, maxIter=2000)
    got, n, init, final = helpers.smooth_solve_emulated(pv, s, np.zeros(N), tol=1e-12, maxIter=2000)
    assert p.finalResidual < 1e-12 and final < 1e-12
    assert init == pytest.approx(p.initialResidual, rel=1e-12)
    assert np.linalg.norm(got - psi) <= 1e-8 * np.linalg.norm(psi)
    assert n <= 2 * p.nIterations + 2          # a different ordering of the same smoother, not a different method
    # the residual it reports (last group's share formed in-kernel) is the true residual of what it returns, up to
    # the rounding of rounding-level terms
    true = np.abs(orc.residual_asym(s, got)[0]).sum() / p.normFactor
    assert final == pytest.approx(true, rel=1e-3)
    for sm in ("GaussSeidel", "symGaussSeidel"):
        got, n, init, final = helpers.smooth_solve_emulated(pv, s, np.zeros(N), smoother=sm, tol=1e-6, maxIter=10, nSweeps=2)
        true = np.abs(orc.residual_asym(s, got)[0]).sum() / p.normFactor
        assert final == pytest.approx(true, rel=1e-7) and n % 2 == 0


@pytest.mark.parametrize("smoother", ["GaussSeidel", "symGaussSeidel"])
def test_skipped_groups_are_bit_neutral(smoother):
    """the sweeps skip two recomputations (last group of the forward half at the start of the reverse half, group 0
    at the start of the next forward half): same bits as visiting every group every time.  (Level plans of the hex
    system; the multicolour plan of the polyhedral one: more than two colours.)"""
    sysd = dict(systems())
    for ordering, mode, s in ((LEVELS, "exact", sysd["hex"]), (MULTICOLOUR, "multicolour", sysd["poly"])):
        pv = helpers.PlanView(ordering, s.addr)
        assert pv.nColours > 2
        N = s.addr.nCells
        low = s.upper if s.lower is None else s.lower
        val = pv.values_asym(s.upper, low, s.addr.lowerAddr)
        d, b, x = pv.to_internal(s.diag), pv.to_internal(s.source), np.zeros(N)
        for _ in range(3):
            for k in range(pv.nColours):
                pv.gs_rows(k, d, val, b, x)
            if smoother == "symGaussSeidel":
                for k in range(pv.nColours - 1, -1, -1):
                    pv.gs_rows(k, d, val, b, x)
        got, _, _, _ = helpers.smooth_solve_emulated(pv, s, np.zeros(N), smoother=smoother, nSweeps=-3, mode=mode)
        assert np.array_equal(got, pv.to_natural(x))


def test_two_colour_symmetric_sweep_is_two_red_black_sweeps():
    """On a two-colour plan the reverse half of a symmetric sweep only recomputes one colour: a red-black symGaussSeidel
    sweep would be ONE red-black Gauss-Seidel sweep, half as strong as upstream's forward + reverse pass.  The
    multicolour mode executes it as two red-black sweeps: same row updates as a symmetric sweep, and a sweep count
    close to upstream's natural-order symGaussSeidel instead of twice it."""
    s = dict(systems())["hex"]
    pv = helpers.PlanView(MULTICOLOUR, s.addr)
    assert pv.nColours == 2
    N = s.addr.nCells
    a = helpers.smooth_solve_emulated(pv, s, np.zeros(N), smoother="symGaussSeidel", nSweeps=-3)
    b = helpers.smooth_solve_emulated(pv, s, np.zeros(N), smoother="GaussSeidel", nSweeps=-6)
    assert np.array_equal(a[0], b[0])
    psi = np.zeros(N)
    up = orc.smooth_solve(s, psi, smoother="symGaussSeidel", tolerance=1e-9, maxIter=500)       # upstream's order
    mc = helpers.smooth_solve_emulated(pv, s, np.zeros(N), smoother="symGaussSeidel", tol=1e-9, maxIter=500)
    gs = helpers.smooth_solve_emulated(pv, s, np.zeros(N), smoother="GaussSeidel", tol=1e-9, maxIter=500)
    assert abs(mc[1] - up.nIterations) <= max(2, up.nIterations // 4), (mc[1], up.nIterations)
    assert gs[1] >= 2 * mc[1] - 2


@pytest.mark.parametrize("smoother", ["GaussSeidel", "symGaussSeidel"])
@pytest.mark.parametrize("nSweeps", [1, 2])
def test_lagged_residual_of_two_colour_plans(smoother, nSweeps):
    """hex meshes are two-coloured: the residual of an iteration is completed by the first pass of the next one
    (no residual kernel).  Same iterates, same sweep counts, same residuals (to the rounding of the evaluation order)
    as the form with an explicit residual pass; stops on the iteration that converged, not one later."""
    _, s = list(systems())[1]
    pv = helpers.PlanView(MULTICOLOUR, s.addr)
    assert pv.nColours == 2
    N = s.addr.nCells
    for ctl in (dict(tol=1e-9, maxIter=500), dict(tol=1e-6, maxIter=10), dict(tol=1e-30, maxIter=7),
                dict(tol=1e-3, minIter=4, maxIter=100), dict(tol=1e-30, relTol=0.01, maxIter=100)):
        a = helpers.smooth_solve_emulated(pv, s, np.zeros(N), smoother=smoother, nSweeps=nSweeps, lag=True, **ctl)
        b = helpers.smooth_solve_emulated(pv, s, np.zeros(N), smoother=smoother, nSweeps=nSweeps, lag=False, **ctl)
        assert a[1] == b[1] and np.array_equal(a[0], b[0])
        assert a[3] == pytest.approx(b[3], rel=1e-6)
