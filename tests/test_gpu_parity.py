"""GPU parity tests proper: the CUDA path (through the C ABI / the B200PCG solver mirror) against
the CPU oracle on the same seeded inputs.  Bars (BASELINE.json north_star):
  * coefficients (assembly), Amul, flux: bit-exact (integer-like determinism of element-wise fp64),
  * PCG + diagonal / none: identical iteration counts, solution within 1e-12 relative,
  * DIC-exact (level-scheduled DIC): identical iteration counts to the oracle's DIC,
  * DIC-class (multicolour IC0): solution within 1e-8 relative L2 at the same residual tolerance.
"""
import json
import os

import numpy as np
import pytest

from firefoam_dev_b200 import B200PCG, B200Error, LduAddressing, LduMatrix, meshgen as mg
from firefoam_dev_b200.cases import StecklerHydrostatic
from firefoam_dev_b200.meshgen import System
from oracle import oracle as orc
from helpers import hydrostatic_loop, random_ldu, STECKLER_DIAG_COUNTS

pytestmark = pytest.mark.gpu

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "steckler_log.json")))


def cases():
    return [("hex", mg.hex_block(24, 20, 16)), ("hex-odd", mg.hex_block(7, 5, 3)),
            ("random", random_ldu(5001, 6.0, seed=7)), ("ragged", random_ldu(1000, 1.5, seed=9))]


def solve_gpu(ctx, s, pre, tol=1e-6, relTol=0.0, maxIter=1000, minIter=0, psi0=None, exact=False):
    ctl = {"solver": "B200PCG", "preconditioner": pre, "tolerance": tol, "relTol": relTol,
           "maxIter": maxIter, "minIter": minIter}
    if exact:
        ctl["B200"] = {"dicMode": "exact"}
    psi = np.zeros(s.addr.nCells) if psi0 is None else psi0.copy()
    solver = B200PCG("p_rgh", s.matrix, s.bou, None, s.interfaces, ctl, context=ctx)
    perf = solver.solve(psi, s.source)
    return psi, perf


def solve_cpu(s, pre, tol=1e-6, relTol=0.0, maxIter=1000, minIter=0, psi0=None):
    psi = np.zeros(s.addr.nCells) if psi0 is None else psi0.copy()
    return psi, orc.pcg_solve(s, psi, pre, tol, relTol, maxIter, minIter)


def relmax(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("name,s", cases())
def test_amul_bit_exact(ctx, name, s):
    ctx.set_addressing(s.addr)
    x = np.random.default_rng(1).standard_normal(s.addr.nCells)
    y = ctx.amul(s.matrix, s.bou, x)
    assert np.array_equal(y, orc.amul(s, x)[0])


AMUL_VARIANTS = [{"B200PCG_SPMV": "ell"}, {"B200PCG_SPMV": "sym"}, {"B200PCG_SPMV": "tma"},
                 {"B200PCG_SPMV": "tma", "B200PCG_EXACT": "0"}, {"B200PCG_SPMV": "tma", "B200PCG_STAGES": "3"}]


@pytest.mark.parametrize("env", AMUL_VARIANTS, ids=lambda e: ",".join(f"{k[8:]}={v}" for k, v in e.items()))
def test_amul_variants_bit_exact(env):
    """Every Amul kernel variant (full-row ELL, symmetric direct, bulk-copy staged, shared-memory
    window with different run lengths / grid sizes) against the oracle, on systems that span many
    256-row chunks and runs: banded hex (window hits), a mesh whose last chunk is partial, and a
    random graph (almost every neighbour outside the window).  Also the solve on top of it."""
    c = _ctx_with_env(env)
    try:
        for s in (mg.hex_block(64, 40, 33), mg.hex_block(37, 23, 11), random_ldu(20011, 6.0, seed=3),
                  mg.hex_block(5, 3, 2)):
            c.set_addressing(s.addr)
            x = np.random.default_rng(5).standard_normal(s.addr.nCells)
            assert np.array_equal(c.amul(s.matrix, s.bou, x), orc.amul(s, x)[0])
        s = mg.hex_block(64, 40, 33)
        psi, perf = solve_gpu(c, s, "diagonal", maxIter=3000)
        ref, pref = solve_cpu(s, "diagonal", maxIter=3000)
        assert perf.nIterations == pref.nIterations
        assert relmax(psi, ref) < 1e-12
    finally:
        c.close()


@pytest.mark.parametrize("sign", [-1.0, 1.0])
def test_assembly_bit_exact(ctx, sign):
    s = mg.hex_block(20, 11, 9)
    a = s.addr
    ctx.set_addressing(a)
    up, dg = ctx.assemble_laplacian(s.gamma_f, s.magSf, s.deltaCoeffs, sign, s.diag0)
    up_ref, dg_ref = orc.laplacian_assemble(a.lowerAddr, a.upperAddr, a.nCells, s.gamma_f, s.magSf,
                                            s.deltaCoeffs, sign, s.diag0)
    assert np.array_equal(up, up_ref) and np.array_equal(dg, dg_ref)
    if sign < 0:
        assert np.array_equal(up, s.upper) and np.array_equal(dg, s.diag)


def test_assembly_irregular_bit_exact(ctx):
    s = random_ldu(3000, 7.0, seed=21)
    a = s.addr
    rng = np.random.default_rng(2)
    g, sf, d = rng.uniform(0.5, 2, a.nFaces), rng.uniform(0.1, 1, a.nFaces), rng.uniform(1, 9, a.nFaces)
    d0 = rng.uniform(0, 1, a.nCells)
    ctx.set_addressing(a)
    up, dg = ctx.assemble_laplacian(g, sf, d, -1.0, d0)
    up_ref, dg_ref = orc.laplacian_assemble(a.lowerAddr, a.upperAddr, a.nCells, g, sf, d, -1.0, d0)
    assert np.array_equal(up, up_ref) and np.array_equal(dg, dg_ref)


def test_flux_bit_exact(ctx):
    s = mg.hex_block(13, 9, 8)
    ctx.set_addressing(s.addr)
    x = np.random.default_rng(3).standard_normal(s.addr.nCells)
    assert np.array_equal(ctx.flux(s.matrix, x), orc.flux(s.addr.lowerAddr, s.addr.upperAddr, s.upper, x))


@pytest.mark.parametrize("name,s", cases())
def test_pcg_diagonal_iterations_identical_solution_1e12(ctx, name, s):
    xg, pg = solve_gpu(ctx, s, "diagonal", tol=1e-6, maxIter=5000)
    xc, pc = solve_cpu(s, "diagonal", tol=1e-6, maxIter=5000)
    assert pg.nIterations == pc.nIterations
    assert pg.converged and pc.converged
    assert pg.initialResidual == pytest.approx(pc.initialResidual, rel=1e-12)
    assert pg.finalResidual == pytest.approx(pc.finalResidual, rel=1e-6)
    assert pg.normFactor == pytest.approx(pc.normFactor, rel=1e-13)
    assert relmax(xg, xc) < 1e-12
    assert str(pg).startswith("diagonalB200PCG:  Solving for p_rgh, Initial residual = ")


@pytest.mark.parametrize("name,s", cases())
def test_pcg_unpreconditioned(ctx, name, s):
    """`preconditioner none`: unpreconditioned CG on these badly scaled systems amplifies the
    1e-16 difference in summation order by ten orders of magnitude within ~50 iterations (measured:
    tools/diverge.py; any two CPU builds with different sum orders do the same), so the bar here is
    the first iterations to rounding, the count to +-3 %, and the converged solution."""
    xg, pg = solve_gpu(ctx, s, "none", tol=1e-30, maxIter=9)
    xc, pc = solve_cpu(s, "none", tol=1e-30, maxIter=9)
    assert pg.nIterations == pc.nIterations == 10
    assert pg.finalResidual == pytest.approx(pc.finalResidual, rel=1e-11)
    assert relmax(xg, xc) < 1e-12
    xg, pg = solve_gpu(ctx, s, "none", tol=1e-6, maxIter=5000)
    xc, pc = solve_cpu(s, "none", tol=1e-6, maxIter=5000)
    assert pg.converged and abs(pg.nIterations - pc.nIterations) <= max(3, 0.03 * pc.nIterations)
    assert relmax(xg, xc) < 1e-4
    assert pg.solverName == "noneB200PCG"


@pytest.mark.parametrize("name,s", cases())
def test_dic_exact_matches_oracle_dic(ctx, name, s):
    xg, pg = solve_gpu(ctx, s, "DIC", tol=1e-6, maxIter=5000, exact=True)
    xc, pc = solve_cpu(s, "DIC", tol=1e-6, maxIter=5000)
    assert pg.nIterations == pc.nIterations
    assert relmax(xg, xc) < 1e-12
    assert pg.solverName == "DICB200PCG"


@pytest.mark.parametrize("name,s", cases())
def test_dic_class_multicolour_solution_1e8(ctx, name, s):
    xg, pg = solve_gpu(ctx, s, "DIC", tol=1e-11, maxIter=5000)
    xc, pc = solve_cpu(s, "DIC", tol=1e-11, maxIter=5000)
    assert pg.converged and pc.converged
    assert np.linalg.norm(xg - xc) / np.linalg.norm(xc) < 1e-8
    # and it must actually precondition: fewer iterations than plain diagonal
    _, pd = solve_cpu(s, "diagonal", tol=1e-11, maxIter=5000)
    assert pg.nIterations < pd.nIterations


def test_steckler_kat_on_gpu_golden_log(ctx):
    """The reference's golden log (log.fireFoam:92-101: 29, 32, 7, 0, 0 DICPCG iterations, the printed
    residuals and hydrostatic variations) reproduced to the printed digits THROUGH THE CUDA PATH:
    assembly kernel + level-scheduled DIC + device-resident PCG loop."""
    case = StecklerHydrostatic()
    ctx.set_addressing(case.addr)
    lap = lambda g, s, d, sign, d0: ctx.assemble_laplacian(g, s, d, sign, d0)
    ctl = {"preconditioner": "DIC", "tolerance": case.TOL, "relTol": case.RELTOL, "B200": {"dicMode": "exact"}}
    solve = lambda m, b, psi: B200PCG("ph_rgh", m, [], None, [], ctl, context=ctx).solve(psi, b)
    res = hydrostatic_loop(case, lap, solve)
    gold = GOLD["ph_rgh"]
    assert [r[2] for r in res] == [g["iters"] for g in gold] == [29, 32, 7, 0, 0]
    for k in range(5):
        tol = 1e-7 if k < 3 else 1e-6      # same bars as tests/test_oracle_kat.py
        assert res[k][0] == pytest.approx(gold[k]["initial"], rel=tol), k
        assert res[k][1] == pytest.approx(gold[k]["final"], rel=tol), k
        assert res[k][3] == pytest.approx(GOLD["variation"][k]["value"], rel=5e-8), k
    # and identical to the CPU oracle run of the same loop
    case2 = StecklerHydrostatic()
    a = case2.addr
    lap2 = lambda g, s, d, sign, d0: orc.laplacian_assemble(a.lowerAddr, a.upperAddr, a.nCells, g, s, d, sign, d0)
    solve2 = lambda m, b, psi: orc.pcg_solve(System(a, m.diag, m.upper, b), psi, "DIC", case2.TOL, case2.RELTOL)
    ref = hydrostatic_loop(case2, lap2, solve2)
    assert [r[2] for r in res] == [r[2] for r in ref]
    for r, q in zip(res, ref):
        assert r[3] == pytest.approx(q[3], rel=1e-10)


def test_steckler_diagonal_and_multicolour(ctx):
    for pre, expect in (("diagonal", STECKLER_DIAG_COUNTS), ("DIC", None)):
        case = StecklerHydrostatic()
        ctx.set_addressing(case.addr)
        lap = lambda g, s, d, sign, d0: ctx.assemble_laplacian(g, s, d, sign, d0)
        ctl = {"preconditioner": pre, "tolerance": case.TOL, "relTol": case.RELTOL}
        solve = lambda m, b, psi: B200PCG("ph_rgh", m, [], None, [], ctl, context=ctx).solve(psi, b)
        res = hydrostatic_loop(case, lap, solve)
        if expect:
            assert [r[2] for r in res] == expect
        assert res[4][3] == pytest.approx(GOLD["variation"][4]["value"], rel=1e-4)


def test_controls_semantics(ctx):
    s = mg.hex_block(8, 6, 5)
    # maxIter: nIterations++ < maxIter -> maxIter + 1 loop bodies
    for maxIter in (0, 3, 7):
        _, pg = solve_gpu(ctx, s, "diagonal", tol=1e-30, maxIter=maxIter)
        _, pc = solve_cpu(s, "diagonal", tol=1e-30, maxIter=maxIter)
        assert pg.nIterations == pc.nIterations == maxIter + 1 and not pg.converged
    # relTol
    _, pg = solve_gpu(ctx, s, "diagonal", tol=1e-6, relTol=0.5)
    _, pc = solve_cpu(s, "diagonal", tol=1e-6, relTol=0.5)
    assert pg.nIterations == pc.nIterations and pg.converged
    # converged initial guess: 0 iterations; minIter forces them
    x0 = s.xstar * (1 + 1e-9)
    xg, pg = solve_gpu(ctx, s, "diagonal", psi0=x0)
    assert pg.nIterations == 0 and pg.converged and np.array_equal(xg, x0)
    _, pg = solve_gpu(ctx, s, "diagonal", psi0=x0, minIter=2)
    _, pc = solve_cpu(s, "diagonal", psi0=x0, minIter=2)
    assert pg.nIterations == pc.nIterations == 2
    # exact solution -> singular break like OpenFOAM (not an error)
    _, pg = solve_gpu(ctx, s, "diagonal", psi0=s.xstar.copy(), minIter=2)
    _, pc = solve_cpu(s, "diagonal", psi0=s.xstar.copy(), minIter=2)
    assert bool(pg.singular) == bool(pc.singular) and pg.nIterations == pc.nIterations


def test_edge_meshes(ctx):
    # single cell, no faces
    a = LduAddressing(1, [], [])
    m = LduMatrix(a, [2.0], [])
    psi = np.zeros(1)
    perf = B200PCG("p", m, [], None, [], {"preconditioner": "diagonal"}, context=ctx).solve(psi, np.array([3.0]))
    assert psi[0] == pytest.approx(1.5, rel=1e-15) and perf.nIterations == 1
    # two disconnected cells + one face elsewhere (rows with zero faces)
    a = LduAddressing(4, [0], [1])
    m = LduMatrix(a, [2.0, 2.0, 1.0, 4.0], [-1.0])
    s = System(a, m.diag, m.upper, np.array([1.0, 0.0, 5.0, 8.0]))
    for pre, exact in (("none", False), ("diagonal", False), ("DIC", False), ("DIC", True)):
        xg, pg = solve_gpu(ctx, s, pre, tol=1e-14, exact=exact)
        np.testing.assert_allclose(xg, np.linalg.solve(np.array([[2, -1, 0, 0], [-1, 2, 0, 0], [0, 0, 1, 0], [0, 0, 0, 4.0]]), s.source), rtol=1e-12)
    # negative-definite system (ph_rgh form, laplacian not negated): PCG still works
    s = mg.hex_block(6, 5, 4)
    sneg = System(s.addr, -s.diag, -s.upper, -s.source)
    xg, pg = solve_gpu(ctx, sneg, "DIC", tol=1e-8, exact=True)
    xc, pc = solve_cpu(sneg, "DIC", tol=1e-8)
    assert pg.nIterations == pc.nIterations and relmax(xg, xc) < 1e-12


def test_errors(ctx):
    with pytest.raises(B200Error):
        ctx.set_addressing(LduAddressing(3, [1, 0], [2, 1]))     # not upper-triangular
    s = mg.hex_block(4, 4, 4)
    ctx.set_addressing(s.addr)
    with pytest.raises(ValueError):
        B200PCG("p", s.matrix, [], None, [], {"preconditioner": "GAMG"}, context=ctx)
    bad = System(s.addr, s.diag.copy(), s.upper, s.source)
    bad.diag[5] = np.nan
    with pytest.raises(B200Error) as e:
        solve_gpu(ctx, bad, "diagonal")
    assert e.value.code == 7   # B200_ENONFINITE


def test_device_entry_points_match_host(ctx):
    import torch
    s = mg.hex_block(16, 12, 10)
    ctx.set_addressing(s.addr)
    dev = torch.device("cuda")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    g, sf, d = t(s.gamma_f), t(s.magSf), t(s.deltaCoeffs)
    upper = torch.empty(s.addr.nFaces, dtype=torch.float64, device=dev)
    diag = t(s.diag0)
    ctx.assemble_laplacian_device(g, sf, d, -1.0, upper, diag)
    torch.cuda.synchronize()
    assert np.array_equal(upper.cpu().numpy(), s.upper) and np.array_equal(diag.cpu().numpy(), s.diag)
    from firefoam_dev_b200 import make_controls
    ctl, _ = make_controls({"preconditioner": "diagonal", "tolerance": 1e-6})
    psi = torch.zeros(s.addr.nCells, dtype=torch.float64, device=dev)
    perf = ctx.solve_device(diag, upper, [], t(s.source), psi, ctl)
    xh, ph = solve_gpu(ctx, s, "diagonal")
    assert perf.nIterations == ph.nIterations
    assert np.array_equal(psi.cpu().numpy(), xh)     # same kernels, same order: bit-identical


def test_rerun_is_bitwise_reproducible(ctx):
    s = random_ldu(20000, 6.0, seed=31)
    x1, p1 = solve_gpu(ctx, s, "diagonal", tol=1e-9, maxIter=5000)
    x2, p2 = solve_gpu(ctx, s, "diagonal", tol=1e-9, maxIter=5000)
    assert p1.nIterations == p2.nIterations and np.array_equal(x1, x2)
    assert p1.finalResidual == p2.finalResidual


def test_medium_hex_all_preconditioners(ctx):
    """~260 k cells: many blocks, grid-stride loops wrap, several hundred iterations."""
    s = mg.hex_block(64, 64, 64)
    for pre, exact in (("diagonal", False), ("DIC", True)):
        xg, pg = solve_gpu(ctx, s, pre, tol=1e-6, maxIter=5000, exact=exact)
        xc, pc = solve_cpu(s, pre, tol=1e-6, maxIter=5000)
        assert pg.nIterations == pc.nIterations, (pre, pg.nIterations, pc.nIterations)
        assert relmax(xg, xc) < 1e-11, (pre, relmax(xg, xc))
    xg, pg = solve_gpu(ctx, s, "DIC", tol=1e-6, maxIter=5000)
    assert pg.converged and relmax(xg, s.xstar) < 1e-3


def test_singlebox_config1_on_gpu(ctx):
    from firefoam_dev_b200.cases import SingleBoxHydrostatic
    case = SingleBoxHydrostatic()
    ctx.set_addressing(case.addr)
    lap = lambda g, s, d, sign, d0: ctx.assemble_laplacian(g, s, d, sign, d0)
    ctl = {"preconditioner": "diagonal", "tolerance": case.TOL, "relTol": case.RELTOL}
    solve = lambda m, b, psi: B200PCG("ph_rgh", m, [], None, [], ctl, context=ctx).solve(psi, b)
    res = hydrostatic_loop(case, lap, solve)
    assert [r[2] for r in res] == [12, 8, 9, 0, 0]


def test_steckler_p_rgh_replay_config2_on_gpu(ctx):
    from firefoam_dev_b200.cases import steckler_p_rgh_system
    s = steckler_p_rgh_system()
    for pre, exact in (("diagonal", False), ("DIC", True)):
        for rt in (0.01, 0.0):
            xg, pg = solve_gpu(ctx, s, pre, tol=1e-6, relTol=rt, exact=exact)
            xc, pc = solve_cpu(s, pre, tol=1e-6, relTol=rt)
            assert pg.nIterations == pc.nIterations and relmax(xg, xc) < 1e-12


def test_polyhedral_config5_small(ctx):
    """BCC truncated-octahedron mesh (up to 14 faces per cell, irregular lowerAddr segments)."""
    s = mg.bcc_poly(12, 12, 16)
    a = s.addr
    ctx.set_addressing(a)
    up, dg = ctx.assemble_laplacian(s.gamma_f, s.magSf, s.deltaCoeffs, -1.0, s.diag0)
    assert np.array_equal(up, s.upper) and np.array_equal(dg, s.diag)
    x = np.random.default_rng(5).standard_normal(a.nCells)
    assert np.array_equal(ctx.amul(s.matrix, [], x), orc.amul(s, x)[0])
    for pre, exact in (("diagonal", False), ("DIC", True)):
        xg, pg = solve_gpu(ctx, s, pre, tol=1e-6, maxIter=5000, exact=exact)
        xc, pc = solve_cpu(s, pre, tol=1e-6, maxIter=5000)
        assert pg.nIterations == pc.nIterations and relmax(xg, xc) < 1e-12
    xg, pg = solve_gpu(ctx, s, "DIC", tol=1e-11, maxIter=5000)
    xc, pc = solve_cpu(s, "DIC", tol=1e-11, maxIter=5000)
    assert np.linalg.norm(xg - xc) / np.linalg.norm(xc) < 1e-8
    assert pg.nColours >= 4


def _ctx_with_env(env):
    from firefoam_dev_b200 import Context
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        return Context()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.parametrize("mode,small,spmv", [("0", "0", ""), ("1", "0", ""), ("1", "150000", ""), ("1", "0", "sr")])
def test_rcm_renumbering_does_not_change_results(mode, small, spmv):
    """Rows renumbered by reverse Cuthill-McKee (forced) or not at all: assembly, Amul and flux stay
    bit-identical to the oracle, PCG + diagonal keeps the oracle's iteration counts, DIC-exact is
    unaffected (never renumbered), multicolour DIC still converges to the same solution."""
    env = {"B200PCG_RENUMBER": mode, "B200PCG_SMALL_N": small}
    if spmv:
        env["B200PCG_SPMV"] = spmv     # opt-in: single-read face-ordered layout on the renumbered plan (k_spmv_sr)
    c = _ctx_with_env(env)
    try:
        for s in (mg.bcc_poly(9, 8, 10), mg.hex_block(24, 20, 16), random_ldu(5001, 6.0, seed=7)):
            a = s.addr
            c.set_addressing(a)
            assert c.describe()["renumbered_rcm"] == (mode == "1")
            if mode == "1":      # renumbered plans: full-row ELL by default, k_spmv_sr on request
                assert c.describe()["amul_natural"].startswith({"": "k_spmv<", "sr": "k_spmv_sr"}[spmv])
            if s.gamma_f is not None:
                up, dg = c.assemble_laplacian(s.gamma_f, s.magSf, s.deltaCoeffs, -1.0, s.diag0)
                up_ref, dg_ref = orc.laplacian_assemble(a.lowerAddr, a.upperAddr, a.nCells, s.gamma_f, s.magSf,
                                                        s.deltaCoeffs, -1.0, s.diag0)
                assert np.array_equal(up, up_ref) and np.array_equal(dg, dg_ref)
            x = np.random.default_rng(5).standard_normal(a.nCells)
            assert np.array_equal(c.amul(s.matrix, [], x), orc.amul(s, x)[0])
            for pre, exact in (("diagonal", False), ("DIC", True)):
                xg, pg = solve_gpu(c, s, pre, tol=1e-7, maxIter=5000, exact=exact)
                xc, pc = solve_cpu(s, pre, tol=1e-7, maxIter=5000)
                assert pg.nIterations == pc.nIterations and relmax(xg, xc) < 1e-12
            xg, pg = solve_gpu(c, s, "DIC", tol=1e-11, maxIter=5000)
            xc, pc = solve_cpu(s, "DIC", tol=1e-11, maxIter=5000)
            assert np.linalg.norm(xg - xc) / np.linalg.norm(xc) < 1e-8
    finally:
        c.close()


@pytest.mark.parametrize("ctas,fast", [("1", "0"), ("4", "0"), ("8", "0"), ("16", "0"), ("16", "1")])
def test_cluster_kernel_sizes(ctas, fast):
    """k_pcg_small with every cluster size (16 is the non-portable maximum) and k_pcg_small_fast (one and
    two rows per thread, 1..16 CTAs), on systems from a handful of cells to 64 000, all preconditioners,
    against the oracle."""
    c = _ctx_with_env({"B200PCG_SMALL_CTAS": ctas, "B200PCG_SMALL_N": "150000", "B200PCG_SMALL_FAST": fast})
    try:
        for s in (mg.hex_block(40, 40, 40), mg.hex_block(7, 5, 3), random_ldu(1000, 1.5, seed=9),
                  mg.bcc_poly(9, 8, 10), mg.hex_block(32, 31, 30), mg.hex_block(30, 20, 20), mg.bcc_poly(20, 20, 18)):
            for pre, exact in (("diagonal", False), ("none", False), ("DIC", True)):
                xg, pg = solve_gpu(c, s, pre, tol=1e-7, maxIter=4000, exact=exact)
                xc, pc = solve_cpu(s, pre, tol=1e-7, maxIter=4000)
                if pre != "none":
                    assert pg.nIterations == pc.nIterations, (pre, pg.nIterations, pc.nIterations)
                    assert relmax(xg, xc) < 1e-12
                else:
                    assert pg.converged and relmax(xg, xc) < 1e-5
            xg, pg = solve_gpu(c, s, "DIC", tol=1e-11, maxIter=5000)
            xc, pc = solve_cpu(s, "DIC", tol=1e-11, maxIter=5000)
            assert np.linalg.norm(xg - xc) / np.linalg.norm(xc) < 1e-8
        assert c.describe()["small_system_cluster_kernel"]
    finally:
        c.close()


@pytest.mark.parametrize("tile,spmv", [("64", ""), ("256", ""), ("64", "ell"), ("0", "")])
def test_tiled_multicolour_dic_class(tile, spmv):
    """DIC-class on a tiled multicolour order (rows = tile, colour, base position) with the symmetric
    single-read Amul: the preconditioner is the same IC0 as in the colour-major order, so iteration
    counts agree (to the +-1 that a different Amul summation order can cause) and the solution meets
    the 1e-8 bar; the first-colour sweep fused into k_r and the segment sweeps are exercised."""
    env = {"B200PCG_TILE": tile, "B200PCG_SMALL_N": "0"}
    if spmv:
        env["B200PCG_SPMV"] = spmv
    c = _ctx_with_env(env)
    ref_ctx = _ctx_with_env({"B200PCG_TILE": "0", "B200PCG_SMALL_N": "0", "B200PCG_FUSE_FIRST": "0",
                             "B200PCG_SPMV": "ell"})
    try:
        for s in (mg.hex_block(24, 20, 16), mg.bcc_poly(9, 8, 10), random_ldu(5001, 6.0, seed=7),
                  mg.hex_block(64, 40, 33)):
            xg, pg = solve_gpu(c, s, "DIC", tol=1e-11, maxIter=5000)
            xr, pr = solve_gpu(ref_ctx, s, "DIC", tol=1e-11, maxIter=5000)
            xc, pc = solve_cpu(s, "DIC", tol=1e-11, maxIter=5000)
            assert pg.converged and abs(pg.nIterations - pr.nIterations) <= 1, (pg.nIterations, pr.nIterations)
            assert pg.nColours == pr.nColours
            assert np.linalg.norm(xg - xc) / np.linalg.norm(xc) < 1e-8
            assert np.linalg.norm(xg - xr) / np.linalg.norm(xr) < 1e-9
            d = c.describe()
            if tile != "0" and s.addr.nCells > 4 * int(tile):
                assert d["multicolour_tiles"] > 1
                assert d["multicolour_amul"] == ("k_spmv" if spmv == "ell" else d["multicolour_amul"])
                if not spmv:
                    assert d["multicolour_amul"].startswith("k_spmv_sym")
            # diagonal / DIC-exact on the same context are untouched by the tiling
            xg, pg = solve_gpu(c, s, "diagonal", tol=1e-7, maxIter=5000)
            xc, pc = solve_cpu(s, "diagonal", tol=1e-7, maxIter=5000)
            assert pg.nIterations == pc.nIterations and relmax(xg, xc) < 1e-12
    finally:
        c.close()
        ref_ctx.close()


def test_16bit_ell_columns_change_nothing():
    """The full-row ELL kernels (Amul on permuted orders, DIC-class sweeps) with 16-bit column offsets
    give bit-identical results to the 32-bit form on banded meshes (every slice entry fits); on a
    150 000-cell random graph kept in its natural numbering most slice entries are too wide and the plan
    keeps 32-bit columns throughout."""
    base = {"B200PCG_SMALL_N": "0", "B200PCG_SPMV": "ell", "B200PCG_RENUMBER": "0"}
    c16 = _ctx_with_env(dict(base, B200PCG_COL16="1"))
    c32 = _ctx_with_env(dict(base, B200PCG_COL16="0"))
    try:
        for s, mixed in ((mg.hex_block(24, 20, 16), False), (mg.bcc_poly(9, 8, 10), False),
                         (random_ldu(150000, 4.0, seed=11), True)):
            x = np.random.default_rng(3).standard_normal(s.addr.nCells)
            for c in (c16, c32):
                c.set_addressing(s.addr)
            d = c16.describe()
            assert c32.describe()["ell_col16_fraction_natural"] == 0
            assert d["ell_col16_fraction_natural"] == (0.0 if mixed else 1.0)
            y16, y32 = c16.amul(s.matrix, [], x), c32.amul(s.matrix, [], x)
            assert np.array_equal(y16, y32)
            if not mixed:
                assert np.array_equal(y16, orc.amul(s, x)[0])
            for pre, exact in (("DIC", False), ("DIC", True), ("diagonal", False)):
                if mixed and exact:
                    continue       # thousands of levels on a random graph: slow, nothing new
                x16, p16 = solve_gpu(c16, s, pre, tol=1e-8, maxIter=3000, exact=exact)
                x32, p32 = solve_gpu(c32, s, pre, tol=1e-8, maxIter=3000, exact=exact)
                assert p16.nIterations == p32.nIterations and np.array_equal(x16, x32), pre
    finally:
        c16.close()
        c32.close()


def test_staged_copies_of_pageable_memory_change_nothing():
    """B200PCG_STAGED_COPY (default 1): pageable caller arrays travel through two page-locked 16 MB pieces (OpenMP copy
    overlapped with the DMA); several pieces per array, a last partial piece, identical results."""
    s = mg.hex_block(128, 96, 90)          # 1.1 M cells: upper is 26 MB (2 pieces), the vectors 8.8 MB (1 partial)
    base = {"B200PCG_SMALL_N": "0"}
    c0, c1 = _ctx_with_env(dict(base, B200PCG_STAGED_COPY="0")), _ctx_with_env(dict(base, B200PCG_STAGED_COPY="1"))
    try:
        x0, p0 = solve_gpu(c0, s, "diagonal", maxIter=5000)
        x1, p1 = solve_gpu(c1, s, "diagonal", maxIter=5000)
        assert p0.nIterations == p1.nIterations and np.array_equal(x0, x1)
        psi0 = np.random.default_rng(4).standard_normal(s.addr.nCells)     # the initial guess travels too
        x0, p0 = solve_gpu(c0, s, "diagonal", maxIter=20, psi0=psi0)
        x1, p1 = solve_gpu(c1, s, "diagonal", maxIter=20, psi0=psi0)
        assert np.array_equal(x0, x1)
    finally:
        c0.close()
        c1.close()


def test_stale_plan_is_not_reused_for_a_different_mesh_at_the_same_key(ctx):
    """The adapter's mesh key is the ADDRESS of the lduAddressing object: after fvMesh::clearOut / updateMesh a
    different mesh with the same sizes can live there.  b200_set_addressing also compares a fingerprint of sampled
    lowerAddr / upperAddr / faceCells entries: the stale plan must be rebuilt, results must stay the oracle's."""
    n = 6000
    s1 = random_ldu(n, 6.0, seed=21)
    # a second system with the SAME cell and face counts but different connectivity
    rng = np.random.default_rng(22)
    perm = rng.permutation(n).astype(np.int32)
    l2, u2 = perm[s1.addr.lowerAddr], perm[s1.addr.upperAddr]
    lo, hi = np.minimum(l2, u2), np.maximum(l2, u2)
    order = np.lexsort((hi, lo))
    a2 = LduAddressing(n, lo[order].astype(np.int32), hi[order].astype(np.int32))
    up2 = s1.upper[order]
    dg2 = np.zeros(n)
    np.add.at(dg2, a2.lowerAddr, -up2)
    np.add.at(dg2, a2.upperAddr, -up2)
    dg2 += 0.05
    s2 = System(a2, dg2, up2, np.zeros(n), [], rng.standard_normal(n))
    s2.source = orc.amul(s2, s2.xstar)[0]
    assert (a2.nCells, a2.nFaces) == (s1.addr.nCells, s1.addr.nFaces)
    a2.mesh_key = s1.addr.mesh_key                       # same "address", same sizes
    for s in (s1, s2, s1):
        x = np.random.default_rng(3).standard_normal(n)
        ctx.set_addressing(s.addr)
        assert np.array_equal(ctx.amul(s.matrix, [], x), orc.amul(s, x)[0])
        xg, pg = solve_gpu(ctx, s, "diagonal", tol=1e-8, maxIter=3000)
        xc, pc = solve_cpu(s, "diagonal", tol=1e-8, maxIter=3000)
        assert pg.nIterations == pc.nIterations and relmax(xg, xc) < 1e-12
