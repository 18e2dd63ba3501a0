"""CPU tests of oracle/bicg_oracle.c (PBiCG + DILU on asymmetric lduMatrices) against independent
formulations: a dense numpy bi-conjugate gradient with a dense DILU factorisation, and -- on a SYMMETRIC matrix, where
DILU == DIC and BiCG == CG -- the digit-pinned DICPCG of oracle/pcg_oracle.c.

The reference's DILUPBiCG log lines (cases/steckler/original/darwinIntel64/log.fireFoam: 207 of them) need matrices
only the whole solver can assemble: PARITY UNPINNED except for the line that needs none (:161, zero system)."""
import numpy as np
import pytest
from scipy.linalg import solve_triangular

import helpers
from firefoam_dev_b200 import cases, meshgen
from oracle import oracle as orc


def dense(s):
    a = s.addr
    A = np.diag(np.asarray(s.diag, dtype=np.float64))
    low = s.upper if s.lower is None else s.lower
    A[a.lowerAddr, a.upperAddr] = s.upper
    A[a.upperAddr, a.lowerAddr] = low
    return A


def dense_dilu(A):
    """E (the DILU diagonal), and M^-1 r = (E+U)^-1 E (E+L)^-1 r as dense triangular solves."""
    n = A.shape[0]
    E = np.diag(A).copy()
    for i in range(n):
        for j in range(i):
            if A[i, j] != 0.0 or A[j, i] != 0.0:
                E[i] -= A[j, i] * A[i, j] / E[j]
    L, U = np.tril(A, -1), np.triu(A, 1)
    lowT, upT = np.diag(E) + L, np.diag(E) + U
    apply = lambda r: solve_triangular(upT, E * solve_triangular(lowT, r, lower=True), lower=False)
    applyT = lambda r: solve_triangular(lowT.T, E * solve_triangular(upT.T, r, lower=True), lower=False)
    return E, apply, applyT


def dense_pbicg(A, b, x0, apply, applyT, tol, maxIter):
    x = x0.copy()
    wA, wT = A @ x, A.T @ x
    rA, rT = b - wA, b - wT
    xRef = x.mean()
    sA = A.sum(1) * xRef
    nf = (np.abs(wA - sA) + np.abs(b - sA)).sum() + 1e-20
    init = final = np.abs(rA).sum() / nf
    n, rho = 0, 1e20
    pA = pT = None
    if final >= tol:
        while True:
            rho_old = rho
            wA, wT = apply(rA), applyT(rT)
            rho = wA @ rT
            pA, pT = (wA, wT) if n == 0 else (wA + (rho / rho_old) * pA, wT + (rho / rho_old) * pT)
            wA, wT = A @ pA, A.T @ pT
            alpha = rho / (wA @ pT)
            x, rA, rT = x + alpha * pA, rA - alpha * wA, rT - alpha * wT
            final = np.abs(rA).sum() / nf
            n += 1
            if not (n - 1 < maxIter and final >= tol):
                break
    return x, n, init, final


def small_transport(N=70, deg=5, seed=3, **kw):
    return cases.transport_system(helpers.random_ldu(N, deg, seed), **kw)


def test_tmul_and_dilu_against_dense():
    s = small_transport()
    A = dense(s)
    x = np.random.default_rng(0).standard_normal(s.addr.nCells)
    np.testing.assert_allclose(orc.tmul_asym(s, x), A.T @ x, rtol=0, atol=1e-13 * np.abs(A).sum(1).max())
    E, apply, applyT = dense_dilu(A)
    rD, w, wT = orc.dilu(s, x)
    np.testing.assert_allclose(rD, 1.0 / E, rtol=1e-12)
    np.testing.assert_allclose(w, apply(x), rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(wT, applyT(x), rtol=1e-10, atol=1e-13)


@pytest.mark.parametrize("seed", [3, 5, 8])
def test_pbicg_dilu_against_dense_bicg(seed):
    s = small_transport(seed=seed, kappa=0.05)
    A = dense(s)
    N = s.addr.nCells
    _, apply, applyT = dense_dilu(A)
    for tol in (1e-6, 1e-10):
        x, n, init, final = dense_pbicg(A, s.source, np.zeros(N), apply, applyT, tol, 1000)
        psi = np.zeros(N)
        p = orc.pbicg_solve(s, psi, "DILU", tolerance=tol, maxIter=1000)
        assert p.nIterations == n and n > 1
        assert p.initialResidual == pytest.approx(init, rel=1e-12) and p.finalResidual == pytest.approx(final, rel=1e-6)
        np.testing.assert_allclose(psi, x, rtol=0, atol=1e-9 * np.abs(x).max())
    for pre, ap in (("diagonal", lambda r: r / np.diag(A)), ("none", lambda r: r)):
        x, n, init, final = dense_pbicg(A, s.source, np.zeros(N), ap, ap, 1e-8, 1000)
        psi = np.zeros(N)
        p = orc.pbicg_solve(s, psi, pre, tolerance=1e-8, maxIter=1000)
        assert abs(p.nIterations - n) <= 1 and np.abs(psi - x).max() <= 1e-6 * np.abs(x).max()


def test_symmetric_matrix_reproduces_the_pinned_dicpcg():
    """DILU of a symmetric matrix is DIC, BiCG of a symmetric matrix is CG: on the digit-pinned steckler ph_rgh system
    (tests/golden/steckler_ph_rgh_c1.b200sys, the system behind log.fireFoam:92) PBiCG + DILU needs DICPCG's 29
    iterations and returns its residual -- the one link between this file and a reference artefact."""
    import os
    from firefoam_dev_b200 import replay
    from conftest import ROOT
    d = replay.read_dump(os.path.join(ROOT, "tests", "golden", "steckler_ph_rgh_c1.b200sys"))
    psi = d.psi0.copy()
    p = orc.pbicg_solve(d.system, psi, "DILU", d.controls["tolerance"], d.controls["relTol"], d.controls["maxIter"])
    assert p.nIterations == d.reference["nIterations"] == 29
    assert p.finalResidual == pytest.approx(d.reference["finalResidual"], rel=1e-9)
    np.testing.assert_allclose(psi, d.psi, rtol=0, atol=1e-11 * np.abs(d.psi).max())


def test_control_flow_and_the_zero_system_line():
    s = small_transport(seed=7, kappa=0.05)
    N = s.addr.nCells
    # cases/steckler/original/darwinIntel64/log.fireFoam:161 "DILUPBiCG:  Solving for H2O, Initial residual = 0,
    # Final residual = 0, No Iterations 0"
    z = meshgen.System(s.addr, s.diag, s.upper, np.zeros(N), [], None, lower=s.lower)
    psi = np.zeros(N)
    p = orc.pbicg_solve(z, psi, "DILU", tolerance=1e-8)
    assert (p.initialResidual, p.finalResidual, p.nIterations) == (0.0, 0.0, 0) and not psi.any()
    # maxIter: nIterations++ < maxIter is tested with the OLD value -> maxIter + 1 loop bodies
    psi = np.zeros(N)
    assert orc.pbicg_solve(s, psi, "DILU", tolerance=1e-30, maxIter=3).nIterations == 4
    psi = s.xstar.copy()
    assert orc.pbicg_solve(s, psi, "DILU", tolerance=1e-3, minIter=2).nIterations == 2
    psi = np.zeros(N)
    p = orc.pbicg_solve(s, psi, "DILU", tolerance=1e-30, relTol=0.01)
    assert p.finalResidual < 0.01 * p.initialResidual and p.converged
