"""CPU tests of the data structure + operation order the PBiCG / DILU kernels use (numpy transliteration in
tests/helpers.py) against oracle/bicg_oracle.c: on the Levels plan (rows grouped by dependency level of the cell order,
entries [lower | upper] each in ascending face order) the row-form triangular sweeps reproduce DILUPreconditioner's
face loops -- losort order included -- BIT FOR BIT; the multicolour plan is the DILU-class stand-in."""
import numpy as np
import pytest

import helpers
from firefoam_dev_b200 import cases, meshgen
from oracle import oracle as orc

LEVELS, MULTICOLOUR = 2, 1


def systems():
    yield "random", cases.transport_system(helpers.random_ldu(120, 6, 17), seed=5, kappa=0.05)
    yield "hex", cases.transport_system(meshgen.hex_block(7, 5, 6), seed=6, kappa=0.05)
    yield "poly", cases.transport_system(meshgen.bcc_poly(4, 3, 3), seed=7, kappa=0.1)
    b = helpers.random_ldu(90, 5, 23)
    yield "symmetric", meshgen.System(b.addr, b.diag, b.upper, b.source, [], b.xstar)


SYS = list(systems())


@pytest.mark.parametrize("name,s", SYS, ids=[n for n, _ in SYS])
def test_level_scheduled_dilu_is_bit_identical(name, s):
    pv = helpers.PlanView(LEVELS, s.addr)
    low = s.upper if s.lower is None else s.lower
    val = pv.values_asym(s.upper, low, s.addr.lowerAddr)
    valT = pv.values_asym(low, s.upper, s.addr.lowerAddr)
    r = np.random.default_rng(4).standard_normal(s.addr.nCells)
    rD_ref, w_ref, wT_ref = orc.dilu(s, r)
    rD = helpers.dilu_calc_rd_emulated(pv, pv.to_internal(s.diag), val, valT)
    assert np.array_equal(pv.to_natural(rD), rD_ref)
    assert np.array_equal(pv.to_natural(pv.dic_precondition(rD, val, pv.to_internal(r))), w_ref)
    assert np.array_equal(pv.to_natural(pv.dic_precondition(rD, valT, pv.to_internal(r))), wT_ref)
    # Amul / Tmul row sums in face order
    x = np.random.default_rng(5).standard_normal(s.addr.nCells)
    d = pv.to_internal(s.diag)
    assert np.array_equal(pv.to_natural(pv.spmv(d, val, pv.to_internal(x))), orc.amul_asym(s, x)[0])
    assert np.array_equal(pv.to_natural(pv.spmv(d, valT, pv.to_internal(x))), orc.tmul_asym(s, x))


@pytest.mark.parametrize("name,s", SYS, ids=[n for n, _ in SYS])
@pytest.mark.parametrize("pre", ["DILU", "diagonal", "none"])
def test_pbicg_on_the_level_plan_matches_the_oracle(name, s, pre):
    pv = helpers.PlanView(LEVELS, s.addr)
    N = s.addr.nCells
    for ctl in (dict(tol=1e-6, maxIter=1000), dict(tol=1e-10, maxIter=1000), dict(tol=1e-30, maxIter=5), dict(tol=1e-3, minIter=3)):
        psi = np.zeros(N)
        p = orc.pbicg_solve(s, psi, pre, tolerance=ctl["tol"], maxIter=ctl.get("maxIter", 1000), minIter=ctl.get("minIter", 0))
        got, n, init, final = helpers.pbicg_emulated(pv, s, np.zeros(N), precond=pre, **ctl)
        # (un-preconditioned BiCG amplifies the 1e-16 difference in the order of the dot-product sums, like `none` PCG)
        assert abs(n - p.nIterations) <= (1 if pre == "none" else 0), (ctl, n, p.nIterations)
        assert init == pytest.approx(p.initialResidual, rel=1e-12)
        assert np.abs(got - psi).max() <= (1e-5 if pre == "none" else 1e-10) * np.abs(psi).max()


@pytest.mark.parametrize("name,s", SYS, ids=[n for n, _ in SYS])
def test_multicolour_dilu_class_converges_to_the_same_solution(name, s):
    pv = helpers.PlanView(MULTICOLOUR, s.addr)
    N = s.addr.nCells
    psi = np.zeros(N)
    p = orc.pbicg_solve(s, psi, "DILU", tolerance=1e-12, maxIter=2000)
    got, n, init, final = helpers.pbicg_emulated(pv, s, np.zeros(N), precond="DILU", tol=1e-12, maxIter=2000)
    assert p.finalResidual < 1e-12 and final < 1e-12
    assert np.linalg.norm(got - psi) <= 1e-8 * np.linalg.norm(psi)
