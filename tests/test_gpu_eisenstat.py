"""GPU parity of the Eisenstat form of the DIC-class PCG (`preconditioner DIC; B200 { dicMode eisenstat; }`,
B200_PRECOND_DIC_MC_EIS): same preconditioner as the multicolour DIC-class mode, applied so that the two
triangular sweeps also deliver A*p (no separate Amul) and the true residual is evaluated lazily.
Bars: solution within 1e-8 relative L2 of the oracle's DIC solve at the same tolerance (the DIC-class bar
of BASELINE.json), same iteration count as the three-kernel DIC-class loop up to a skipped early dip
below the threshold (<= +2), the reported final residual is the true residual of the returned solution."""
import os

import numpy as np
import pytest

from firefoam_dev_b200 import B200PCG, B200Error, meshgen as mg
from firefoam_dev_b200.cases import StecklerHydrostatic
from oracle import oracle as orc
from helpers import hydrostatic_loop, random_ldu
from test_gpu_parity import GOLD, _ctx_with_env, cases, relmax, solve_cpu

pytestmark = pytest.mark.gpu


def solve_mode(ctx, s, mode, tol=1e-6, relTol=0.0, maxIter=5000, minIter=0, psi0=None):
    ctl = {"solver": "B200PCG", "preconditioner": "DIC", "tolerance": tol, "relTol": relTol,
           "maxIter": maxIter, "minIter": minIter, "B200": {"dicMode": mode}}
    psi = np.zeros(s.addr.nCells) if psi0 is None else psi0.copy()
    perf = B200PCG("p_rgh", s.matrix, s.bou, None, s.interfaces, ctl, context=ctx).solve(psi, s.source)
    return psi, perf


def true_residual(ctx, s, x, normFactor):
    ctx.set_addressing(s.addr)
    return np.abs(s.source - ctx.amul(s.matrix, s.bou, x)).sum() / normFactor


@pytest.mark.parametrize("name,s", cases())
def test_eisenstat_solution_1e8_and_same_iterations(ctx, name, s):
    for tol in (1e-6, 1e-11):
        xe, pe = solve_mode(ctx, s, "eisenstat", tol=tol)
        xm, pm = solve_mode(ctx, s, "multicolour", tol=tol)
        assert pe.converged and pm.converged
        assert pm.nIterations <= pe.nIterations <= pm.nIterations + 2, (tol, pe.nIterations, pm.nIterations)
        assert pe.initialResidual == pytest.approx(pm.initialResidual, rel=1e-12)
        assert pe.finalResidual < tol
        assert pe.finalResidual == pytest.approx(true_residual(ctx, s, xe, pe.normFactor), rel=1e-3, abs=1e-15)
        assert np.linalg.norm(xe - xm) / np.linalg.norm(xm) < 1e-8
    xc, pc = solve_cpu(s, "DIC", tol=1e-11, maxIter=5000)
    assert np.linalg.norm(xe - xc) / np.linalg.norm(xc) < 1e-8
    assert str(pe).startswith("DIC(mc)B200PCG:  Solving for p_rgh, Initial residual = ")


def test_eisenstat_medium_hex_multi_kernel_path(ctx):
    """~260 k cells: above the cluster-kernel limit, so every ctx variant runs the k_eis_* kernels; many
    blocks per colour, grid-stride loops wrap, ~100 iterations with lazily evaluated residuals."""
    s = mg.hex_block(64, 64, 64)
    xe, pe = solve_mode(ctx, s, "eisenstat")
    xm, pm = solve_mode(ctx, s, "multicolour")
    assert pe.converged and pm.nIterations <= pe.nIterations <= pm.nIterations + 2
    assert np.linalg.norm(xe - xm) / np.linalg.norm(xm) < 1e-8
    assert pe.finalResidual == pytest.approx(true_residual(ctx, s, xe, pe.normFactor), rel=1e-3)
    assert relmax(xe, s.xstar) < 1e-3
    # bit-reproducible re-run
    xe2, pe2 = solve_mode(ctx, s, "eisenstat")
    assert np.array_equal(xe, xe2) and pe2.nIterations == pe.nIterations


def test_eisenstat_controls_semantics(ctx):
    s = mg.hex_block(8, 6, 5)
    for maxIter in (0, 3, 7):     # nIterations++ < maxIter -> maxIter + 1 loop bodies
        _, pe = solve_mode(ctx, s, "eisenstat", tol=1e-30, maxIter=maxIter)
        assert pe.nIterations == maxIter + 1 and not pe.converged
    _, pe = solve_mode(ctx, s, "eisenstat", tol=1e-6, relTol=0.5)
    _, pm = solve_mode(ctx, s, "multicolour", tol=1e-6, relTol=0.5)
    assert pe.converged and pm.nIterations <= pe.nIterations <= pm.nIterations + 2
    x0 = s.xstar * (1 + 1e-9)
    xg, pe = solve_mode(ctx, s, "eisenstat", psi0=x0)
    assert pe.nIterations == 0 and pe.converged and np.array_equal(xg, x0)
    _, pe = solve_mode(ctx, s, "eisenstat", psi0=x0, minIter=2)
    assert pe.nIterations == 2
    # exact solution: zero residual -> singular break, not an error
    _, pe = solve_mode(ctx, s, "eisenstat", psi0=s.xstar.copy(), minIter=2)
    _, pm = solve_mode(ctx, s, "multicolour", psi0=s.xstar.copy(), minIter=2)
    assert bool(pe.singular) == bool(pm.singular) and pe.nIterations == pm.nIterations


def test_eisenstat_negative_definite_hydrostatic_loop(ctx):
    """ph_rghEqn (solver/phrghEqn.H:45, not negated): negative-definite system, D~ < 0, rho < 0; the
    converged hydrostatic variation of the reference's golden log (log.fireFoam:97)."""
    case = StecklerHydrostatic()
    ctx.set_addressing(case.addr)
    lap = lambda g, s, d, sign, d0: ctx.assemble_laplacian(g, s, d, sign, d0)
    ctl = {"preconditioner": "DIC", "tolerance": case.TOL, "relTol": case.RELTOL, "B200": {"dicMode": "eisenstat"}}
    solve = lambda m, b, psi: B200PCG("ph_rgh", m, [], None, [], ctl, context=ctx).solve(psi, b)
    res = hydrostatic_loop(case, lap, solve)
    assert res[4][3] == pytest.approx(GOLD["variation"][4]["value"], rel=1e-4)
    assert res[3][2] == 0 and res[4][2] == 0


@pytest.mark.parametrize("env", [{"B200PCG_COL16": "0"}, {"B200PCG_RENUMBER": "1"}, {"B200PCG_RENUMBER": "0"},
                                 {"B200PCG_EIS_BATCH": "0"}, {"B200PCG_EIS_BATCH": "0", "B200PCG_COL16": "0"},
                                 {"B200PCG_SWEEP_CTAS": "2"}, {"B200PCG_EIS_CTAS": "3"}, {"B200PCG_EIS_CTAS": "3", "B200PCG_COL16": "0"}, {"B200PCG_EIS_CTAS": "4"}],
                         ids=["col32", "rcm", "natural-base", "plain-loops", "plain-loops-col32", "2-ctas", "3cta-build", "3cta-build-col32", "4cta-build"])
def test_eisenstat_plan_variants(env):
    """32-bit ELL columns, RCM-renumbered and natural base orders; polyhedral mesh (>= 4 colours: several
    un-fused backward and forward launches) and a random graph, multi-kernel path."""
    c = _ctx_with_env(dict(env, B200PCG_SMALL_N="0"))
    try:
        for s in (mg.bcc_poly(12, 12, 16), random_ldu(20011, 6.0, seed=3), mg.hex_block(37, 23, 11)):
            xe, pe = solve_mode(c, s, "eisenstat", tol=1e-11)
            xm, pm = solve_mode(c, s, "multicolour", tol=1e-11)
            assert pe.converged and pm.nIterations <= pe.nIterations <= pm.nIterations + 2
            assert np.linalg.norm(xe - xm) / np.linalg.norm(xm) < 1e-8
            xc, _ = solve_cpu(s, "DIC", tol=1e-11, maxIter=5000)
            assert np.linalg.norm(xe - xc) / np.linalg.norm(xc) < 1e-8
    finally:
        c.close()


def indefinite_system():
    """a hex system with one DIC pivot of the wrong sign"""
    s = mg.hex_block(64, 64, 64)
    bad = mg.System(s.addr, s.diag.copy(), s.upper, s.source, s.bou, s.xstar)
    bad.diag[1000] = -bad.diag[1000]
    return bad


def test_eisenstat_rejects_indefinite_matrix(ctx):
    """DIC pivots of mixed sign: the symmetric scaling does not exist -> B200_EUNSUPPORTED, context still usable."""
    from firefoam_dev_b200 import B200Error
    s = mg.hex_block(64, 64, 64)
    bad = mg.System(s.addr, s.diag.copy(), s.upper, s.source, s.bou, s.xstar)
    bad.diag[1000] = -bad.diag[1000]
    with pytest.raises(B200Error) as ei:
        solve_mode(ctx, bad, "eisenstat")
    assert "mixed sign" in str(ei.value)
    xe, pe = solve_mode(ctx, s, "eisenstat")
    assert pe.converged and relmax(xe, s.xstar) < 1e-3


def test_eisenstat_rejects_tiled_plan():
    c = _ctx_with_env({"B200PCG_TILE": "64", "B200PCG_SMALL_N": "0"})
    try:
        s = mg.hex_block(24, 20, 16)
        with pytest.raises(Exception) as ei:
            solve_mode(c, s, "eisenstat")
        assert "colour-major" in str(ei.value)
    finally:
        c.close()


def test_dic_keyword_defaults_to_the_eisenstat_form():
    """Plain `preconditioner DIC` (no dicMode) takes the Eisenstat form; B200PCG_DIC=multicolour, tiled plans and
    matrices with DIC pivots of mixed sign keep / fall back to the three-kernel loop."""
    s = mg.hex_block(64, 40, 33)
    c = _ctx_with_env({"B200PCG_SMALL_N": "0"})
    try:
        c.profile(True)
        xa, pa = solve_mode(c, s, "auto", tol=1e-9)
        prof = c.profile_json()
        assert "eis_fwd_dot" in prof and "spmv_dot" not in prof
        xe, pe = solve_mode(c, s, "eisenstat", tol=1e-9)
        assert pa.nIterations == pe.nIterations and np.array_equal(xa, xe)
        assert pa.solverName == "DIC(mc)B200PCG"
        c.profile(True)
        xm, pm = solve_mode(c, s, "multicolour", tol=1e-9)       # explicit: always the three-kernel loop
        assert "spmv_dot" in c.profile_json() and "eis_fwd_dot" not in c.profile_json()
        # indefinite matrix (DIC pivots of mixed sign): explicit eisenstat is an error, auto falls back
        si = indefinite_system()
        with pytest.raises(B200Error):
            solve_mode(c, si, "eisenstat", maxIter=20)
        c.profile(True)
        xi, pi = solve_mode(c, si, "auto", maxIter=20)
        xl, pl = solve_mode(c, si, "multicolour", maxIter=20)
        assert "spmv_dot" in c.profile_json()
        assert pi.nIterations == pl.nIterations and np.array_equal(xi, xl)
    finally:
        c.close()
    for env in ({"B200PCG_DIC": "multicolour", "B200PCG_SMALL_N": "0"}, {"B200PCG_SMALL_N": "0", "B200PCG_TILE": "64"}):
        c = _ctx_with_env(env)
        try:
            c.profile(True)
            xt, pt = solve_mode(c, s, "auto", tol=1e-9)
            assert "spmv_dot" in c.profile_json() and pt.converged
        finally:
            c.close()


@pytest.mark.parametrize("mode", ["multicolour", "eisenstat"])
def test_column_sorted_multicolour_plan(mode):
    """B200PCG_SORT_COLS=1: entries of the multicolour plan ordered by column (coalesced gathers on renumbered
    meshes).  Same preconditioner, different summation order inside a row: same solution, iterations +-2."""
    base = {"B200PCG_SMALL_N": "0", "B200PCG_RENUMBER": "1"}
    c0, c1 = _ctx_with_env(dict(base, B200PCG_SORT_COLS="0")), _ctx_with_env(dict(base, B200PCG_SORT_COLS="1"))
    try:
        for s in (mg.bcc_poly(12, 12, 16), random_ldu(20011, 6.0, seed=3), mg.hex_block(37, 23, 11)):
            x0, p0 = solve_mode(c0, s, mode, tol=1e-11)
            x1, p1 = solve_mode(c1, s, mode, tol=1e-11)
            assert p0.converged and p1.converged and abs(p0.nIterations - p1.nIterations) <= 2
            assert np.linalg.norm(x1 - x0) / np.linalg.norm(x0) < 1e-8
    finally:
        c0.close()
        c1.close()
