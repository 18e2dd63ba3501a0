"""CPU tests of oracle/smooth_oracle.c (smoothSolver + GaussSeidel / symGaussSeidel on asymmetric lduMatrices,
SURVEY.md 8f-4) against an INDEPENDENT formulation: dense matrices and scipy triangular solves.

The reference pins these solves by residual lines only (cases/steckler/original/linux64/log.fireFoam:172-178 ...,
261 lines) and their matrices need the whole solver, so the oracle is PARITY UNPINNED for this algorithm; the one
line that needs nothing else (zero field, zero source -> 0, 0, No Iterations 0; log.fireFoam:176) is checked here."""
import numpy as np
import pytest
from scipy.linalg import solve_triangular

from firefoam_dev_b200 import cases, meshgen
from oracle import oracle as orc
import helpers


def dense(s):
    a = s.addr
    A = np.diag(np.asarray(s.diag, dtype=np.float64))
    low = s.upper if s.lower is None else s.lower
    A[a.lowerAddr, a.upperAddr] = s.upper
    A[a.upperAddr, a.lowerAddr] = low
    return A


def small_transport(N=60, deg=5, seed=3, **kw):
    return cases.transport_system(helpers.random_ldu(N, deg, seed), **kw)


def test_amul_sumA_residual_against_dense():
    s = small_transport()
    A = dense(s)
    assert np.abs(A - A.T).max() > 1e-3            # the matrix really is asymmetric
    x = np.random.default_rng(0).standard_normal(s.addr.nCells)
    np.testing.assert_allclose(orc.amul_asym(s, x)[0], A @ x, rtol=0, atol=1e-13 * np.abs(A).sum(1).max())
    np.testing.assert_allclose(orc.sumA_asym(s), A.sum(1), rtol=1e-13)
    np.testing.assert_allclose(orc.residual_asym(s, x)[0], s.source - A @ x, rtol=0, atol=1e-12)


@pytest.mark.parametrize("smoother", ["GaussSeidel", "symGaussSeidel"])
def test_sweeps_equal_triangular_solves(smoother):
    s = small_transport(seed=5)
    A = dense(s)
    DL, U = np.tril(A), np.triu(A, 1)
    DU, L = np.triu(A), np.tril(A, -1)
    x = np.zeros(s.addr.nCells)
    ref = x.copy()
    for nS in (1, 3):
        psi = x.copy()
        perf = orc.smooth_solve(s, psi, smoother=smoother, nSweeps=-nS)
        assert (perf.nIterations, perf.initialResidual, perf.finalResidual) == (nS, 0.0, 0.0)
        ref = x.copy()
        for _ in range(nS):
            ref = solve_triangular(DL, s.source - U @ ref, lower=True)
            if smoother == "symGaussSeidel":
                ref = solve_triangular(DU, s.source - L @ ref, lower=False)
        np.testing.assert_allclose(psi, ref, rtol=1e-12, atol=1e-14)


def test_control_flow_and_log_line_for_a_zero_system():
    s = small_transport(seed=7, kappa=0.05)
    N = s.addr.nCells
    A = dense(s)
    # cases/steckler/original/linux64/log.fireFoam:176  "Solving for H2O, Initial residual = 0, Final residual = 0,
    # No Iterations 0": zero field, zero source
    z = meshgen.System(s.addr, s.diag, s.upper, np.zeros(N), [], None, lower=s.lower)
    psi = np.zeros(N)
    p = orc.smooth_solve(z, psi, tolerance=1e-8, maxIter=10)
    assert (p.initialResidual, p.finalResidual, p.nIterations) == (0.0, 0.0, 0) and not psi.any()
    # the reference's own controls: tolerance 1e-6, relTol 0, maxIter 10 (fvSolution:48-55)
    psi = np.zeros(N)
    p = orc.smooth_solve(s, psi, tolerance=1e-6, relTol=0.0, maxIter=10)
    assert 1 <= p.nIterations <= 10
    nf = p.normFactor
    assert p.finalResidual == pytest.approx(np.abs(s.source - A @ psi).sum() / nf, rel=1e-10)
    assert (p.nIterations == 10) or p.finalResidual < 1e-6
    # maxIter caps: (nIterations += nSweeps) < maxIter is tested AFTER the increment
    for nS, mx, expect in ((1, 3, 3), (2, 3, 4), (4, 10, 12)):
        psi = np.zeros(N)
        p = orc.smooth_solve(s, psi, tolerance=1e-30, maxIter=mx, nSweeps=nS)
        assert p.nIterations == expect
    # minIter forces iterations on a converged system; relTol stops early
    psi = s.xstar.copy()
    p = orc.smooth_solve(s, psi, tolerance=1e-3, minIter=2)
    assert p.nIterations == 2
    psi = np.zeros(N)
    p = orc.smooth_solve(s, psi, tolerance=1e-30, relTol=0.1, maxIter=1000)
    assert p.finalResidual < 0.1 * p.initialResidual and p.nIterations < 1000


def test_symmetric_matrix_through_the_asymmetric_path():
    b = helpers.random_ldu(50, 5, 11)
    s1 = meshgen.System(b.addr, b.diag, b.upper, b.source, [], b.xstar)
    s2 = meshgen.System(b.addr, b.diag, b.upper, b.source, [], b.xstar, lower=b.upper.copy())
    p1, p2 = np.zeros(50), np.zeros(50)
    a = orc.smooth_solve(s1, p1, tolerance=1e-9, maxIter=200)
    c = orc.smooth_solve(s2, p2, tolerance=1e-9, maxIter=200)
    assert a.nIterations == c.nIterations and np.array_equal(p1, p2)
    np.testing.assert_array_equal(orc.amul_asym(s1, b.xstar)[0], orc.amul(b, b.xstar)[0])


def test_two_ranks_couple_like_a_jacobi_interface():
    """processor patches are explicit in the sweeps (psi of the neighbour rank from the START of the sweep):
    R-rank symGaussSeidel == block triangular solves with the off-rank blocks lagged."""
    g = meshgen.hex_block(6, 5, 4)
    s = cases.transport_system(g, seed=21)
    N = s.addr.nCells
    c2p = (np.arange(N) % 6 >= 3).astype(np.int32)              # split in x
    subs = meshgen.decompose(s, c2p, 2)
    A = dense(s)
    x = np.random.default_rng(2).standard_normal(N)
    ys = orc.amul_asym(subs, [x[p.cells] for p in subs])
    y = np.empty(N)
    for p, yp in zip(subs, ys):
        y[p.cells] = yp
    np.testing.assert_allclose(y, A @ x, rtol=0, atol=1e-12 * np.abs(A).sum(1).max())
    psis = [np.zeros(p.addr.nCells) for p in subs]
    for p in subs:
        p.source = s.source[p.cells].copy()
    perf = orc.smooth_solve(subs, psis, nSweeps=-2)
    assert perf.nIterations == 2
    ref = np.zeros(N)
    for _ in range(2):
        new = ref.copy()
        for p in subs:
            own = p.cells
            other = np.setdiff1d(np.arange(N), own)
            Ap = A[np.ix_(own, own)]
            rhs = s.source[own] - A[np.ix_(own, other)] @ ref[other]
            xx = solve_triangular(np.tril(Ap), rhs - np.triu(Ap, 1) @ ref[own], lower=True)
            xx = solve_triangular(np.triu(Ap), rhs - np.tril(Ap, -1) @ xx, lower=False)
            new[own] = xx
        ref = new
    got = np.empty(N)
    for p, q in zip(subs, psis):
        got[p.cells] = q
    np.testing.assert_allclose(got, ref, rtol=1e-11, atol=1e-13)
    # and the solve converges to the global solution with the same controls on 1 and 2 ranks
    psis = [np.zeros(p.addr.nCells) for p in subs]
    perf2 = orc.smooth_solve(subs, psis, tolerance=1e-10, maxIter=500)
    psi1 = np.zeros(N)
    perf1 = orc.smooth_solve(s, psi1, tolerance=1e-10, maxIter=500)
    assert perf2.finalResidual < 1e-10 and perf1.finalResidual < 1e-10
    assert perf2.initialResidual == pytest.approx(perf1.initialResidual, rel=1e-12)
    for p, q in zip(subs, psis):
        np.testing.assert_allclose(q, psi1[p.cells], rtol=0, atol=1e-8 * np.abs(psi1).max())
