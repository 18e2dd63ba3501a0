"""the reference's other symmetric linear solves through the same plug-in.

  G    P1 radiation: fvm::laplacian(gamma, G) - fvm::Sp(a, G) == -4 e sigma T^4 - E
       (packages/thermophysicalModels/radiation/radiationModels/P1/P1.C:238-244), `PCG` + `DIC`, tol 1e-6, relTol 0
       (cases/steckler/system/fvSolution:75-81) -- a negative-definite system with an Sp diagonal and mixed
       (Marshak) boundary coefficients on every wall, restated on the steckler topology (cases.p1_G_terms).
  rho  fvm::ddt(rho) + fvc::div(phi) == sources (solver/rhoEqn.H:33-43): a DIAGONAL lduMatrix.  Upstream's
       lduMatrix::solver::New returns diagonalSolver for it before the `solver` keyword is looked up (the log
       prints `diagonal:  Solving for rho`, cases/steckler/original/linux64/log.fireFoam:209), so it never reaches
       B200PCG; the C ABI still has to cope with nFaces == 0 (the adapter guards on matrix_.diagonal()).
"""
import os

import numpy as np
import pytest

from firefoam_dev_b200 import meshgen as mg
from firefoam_dev_b200.cases import p1_G_terms
from firefoam_dev_b200.ldu import LduAddressing
from firefoam_dev_b200.meshgen import System
from oracle import oracle as orc
from test_prgh_assembly import numpy_assembly


def g_system():
    case, t = p1_G_terms()
    up, dg, src = orc.assemble_p_rgh(case.addr.lowerAddr, case.addr.upperAddr, case.N, t)
    return case, t, System(case.addr, dg, up, src, [])


def test_G_equation_oracle_assembly_and_solve():
    case, t, s = g_system()
    # independent (vectorised) statement of the same fvMatrix algebra
    up2, dg2, src2 = numpy_assembly(case.addr, t)
    assert np.array_equal(s.upper, up2)
    np.testing.assert_allclose(s.diag, dg2, rtol=1e-13)
    np.testing.assert_allclose(s.source, src2, rtol=1e-12)
    # shape of the P1 system: laplacian NOT negated (negative definite), Sp strengthens the diagonal
    assert (s.upper > 0).all() and (s.diag < 0).all()
    rowsum = s.diag.copy()
    np.add.at(rowsum, case.addr.lowerAddr, s.upper)
    np.add.at(rowsum, case.addr.upperAddr, s.upper)
    assert (rowsum < 0).all()                      # strictly diagonally dominant: -a V - Marshak
    counts = {}
    for pre in ("DIC", "diagonal"):
        G = np.zeros(case.N)
        p = orc.pcg_solve(s, G, pre, 1e-6, 0.0, 1000)
        assert p.converged and p.initialResidual == pytest.approx(1.0)
        counts[pre] = p.nIterations
        # incident radiation between the cold-wall and the plume black-body levels
        sig = 5.670367e-08
        assert 4 * sig * 298.15 ** 4 < G.min() and G.max() < 4 * sig * 1100.0 ** 4
    # regression values of the oracle (the reference ships no P1 log: steckler runs fvDOM)
    assert counts == {"DIC": 62, "diagonal": 192}


def test_diagonal_matrix_oracle():
    """rho-shaped system: no faces at all.  PCG on a diagonal matrix is exact after one iteration."""
    N = 1000
    rng = np.random.default_rng(5)
    s = System(LduAddressing(N, [], []), rng.uniform(1, 2, N), np.zeros(0), rng.standard_normal(N), [])
    for pre in ("diagonal", "DIC", "none"):
        x = np.zeros(N)
        p = orc.pcg_solve(s, x, pre, 1e-12, 0.0, 50)
        assert p.converged and (p.nIterations == 1 or pre == "none")
        np.testing.assert_allclose(x, s.source / s.diag, rtol=1e-12 if pre != "none" else 1e-8, atol=1e-10)


@pytest.mark.gpu
def test_G_equation_on_gpu():
    from firefoam_dev_b200 import B200PCG, Context, LduMatrix
    case, t, s = g_system()
    ctx = Context(device=0)
    try:
        ctx.set_addressing(case.addr)
        up, dg, src = ctx.assemble_p_rgh(t)            # device assembly of the whole G equation
        assert np.array_equal(up, s.upper) and np.array_equal(dg, s.diag) and np.array_equal(src, s.source)
        for pre, mode, exact in (("diagonal", None, True), ("DIC", "exact", True), ("DIC", "auto", False),
                                 ("DIC", "multicolour", False)):
            for tol in ((1e-6, 1e-11) if not exact else (1e-6,)):
                ctl = {"preconditioner": pre, "tolerance": tol, "relTol": 0.0, "maxIter": 1000}
                if mode:
                    ctl["B200"] = {"dicMode": mode}
                G = np.zeros(case.N)
                perf = B200PCG("G", LduMatrix(case.addr, dg, up), [], None, [], ctl, context=ctx).solve(G, src)
                ref = np.zeros(case.N)
                pr = orc.pcg_solve(s, ref, pre, tol, 0.0, 1000)
                assert perf.converged
                if exact:
                    assert perf.nIterations == pr.nIterations, (pre, mode)
                    assert np.abs(G - ref).max() <= 1e-12 * np.abs(ref).max()
                    assert perf.finalResidual == pytest.approx(pr.finalResidual, rel=1e-9)
                elif tol == 1e-11:
                    assert np.linalg.norm(G - ref) / np.linalg.norm(ref) < 1e-8
        assert str(perf).startswith("DIC(mc)B200PCG:  Solving for G, Initial residual = 1")
    finally:
        ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("N", [1, 1000, 300000])
def test_diagonal_matrix_on_gpu(N):
    """nFaces == 0 through the C ABI (cluster kernel for the small sizes, multi-kernel loop for the large one)."""
    from firefoam_dev_b200 import B200PCG, Context, LduMatrix
    rng = np.random.default_rng(N)
    a = LduAddressing(N, [], [])
    s = System(a, rng.uniform(1, 2, N), np.zeros(0), rng.standard_normal(N), [])
    ctx = Context(device=0)
    try:
        for pre in ("diagonal", "DIC"):
            x = np.zeros(N)
            perf = B200PCG("rho", LduMatrix(a, s.diag, s.upper), [], None, [], {"preconditioner": pre, "tolerance": 1e-12},
                           context=ctx).solve(x, s.source)
            ref = np.zeros(N)
            pr = orc.pcg_solve(s, ref, pre, 1e-12, 0.0, 1000)
            assert perf.nIterations == pr.nIterations == 1 and perf.converged
            np.testing.assert_allclose(x, s.source / s.diag, rtol=1e-14)
    finally:
        ctx.close()
