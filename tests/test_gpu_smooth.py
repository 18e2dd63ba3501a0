"""GPU parity of the smoothSolver path (U, Yi, h, k -- cases/steckler/system/fvSolution:48-61)
through the C ABI (b200_smooth_solve / b200_amul_asym) against oracle/smooth_oracle.c on the same seeded inputs.

Bars (written out below):
  * asymmetric Amul                             bit-identical to the CPU loop
  * sweepMode exact (level-scheduled sweeps)    psi after a fixed number of sweeps BIT-identical; identical sweep
                                                counts, identical initial residual to 1e-12, final to 1e-9 relative
                                                (only the order of the global |r| sum differs from the CPU)
  * sweepMode multicolour (GS-class)            same solution within 1e-8 relative L2 at a tight tolerance; at the
                                                reference's own controls the residual it reports is the true one
                                                (1e-6 relative); bit-identical to the numpy transliteration of the
                                                kernels on the same plan (tests/helpers.py smooth_solve_emulated)
"""
import numpy as np
import pytest

import helpers
from firefoam_dev_b200 import cases, meshgen
from firefoam_dev_b200.meshgen import System
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def systems():
    yield "random", cases.transport_system(helpers.random_ldu(3000, 6, 17), seed=5)
    yield "hex", cases.transport_system(meshgen.hex_block(23, 17, 19), seed=6)
    yield "hex-odd", cases.transport_system(meshgen.hex_block(33, 1, 31), seed=8)
    yield "poly", cases.transport_system(meshgen.bcc_poly(9, 8, 7), seed=7, kappa=0.3)
    b = helpers.random_ldu(2500, 5, 23)
    yield "symmetric", System(b.addr, b.diag, b.upper, b.source, [], b.xstar)
    st = cases.StecklerHydrostatic()
    m, src = st.assemble(lambda g, ms, dc, sign, d0: orc.laplacian_assemble(st.addr.lowerAddr, st.addr.upperAddr, st.N, g, ms, dc,
                                                                            sign, d0))
    yield "steckler-topology", cases.transport_system(System(st.addr, m.diag, m.upper, src, [], None), seed=9)


SYSTEMS = list(systems())
IDS = [n for n, _ in SYSTEMS]


@pytest.fixture(scope="module")
def gctx():
    from firefoam_dev_b200 import Context
    c = Context(device=0)
    yield c
    c.close()


def solver(gctx, s, **ctl):
    from firefoam_dev_b200 import B200smoothSolver
    return B200smoothSolver("U", s.matrix, [], None, [], ctl, context=gctx)


@pytest.mark.parametrize("name,s", SYSTEMS, ids=IDS)
def test_asymmetric_amul_bit_exact(gctx, name, s):
    x = np.random.default_rng(1).standard_normal(s.addr.nCells)
    gctx.set_addressing(s.addr)
    got = gctx.amul_asym(s.matrix, [], x)
    assert np.array_equal(got, orc.amul_asym(s, x)[0])


@pytest.mark.parametrize("name,s", SYSTEMS, ids=IDS)
@pytest.mark.parametrize("smoother", ["GaussSeidel", "symGaussSeidel"])
def test_exact_sweeps_bit_identical(gctx, name, s, smoother):
    N = s.addr.nCells
    for nS in (1, 3):
        ref = np.zeros(N)
        orc.smooth_solve(s, ref, smoother=smoother, nSweeps=-nS)
        psi = np.zeros(N)
        perf = solver(gctx, s, smoother=smoother, nSweeps=-nS, B200={"sweepMode": "exact"}).solve(psi, s.source)
        assert (perf.nIterations, perf.initialResidual, perf.finalResidual) == (nS, 0.0, 0.0)
        assert np.array_equal(psi, ref), np.abs(psi - ref).max()


@pytest.mark.parametrize("name,s", SYSTEMS, ids=IDS)
@pytest.mark.parametrize("ctl", [dict(tolerance=1e-6, relTol=0.0, maxIter=10),          # fvSolution:48-55 (U)
                                 dict(tolerance=1e-8, relTol=0.0, maxIter=10),          # fvSolution:57-61 (Yi|h|k)
                                 dict(tolerance=1e-10, relTol=0.0, maxIter=1000),
                                 dict(tolerance=1e-30, relTol=0.05, maxIter=1000, nSweeps=2),
                                 dict(tolerance=1e-3, minIter=3, maxIter=1000)],
                         ids=["U", "Yi-h-k", "tight", "relTol-nSweeps2", "minIter"])
def test_exact_solve_matches_oracle(gctx, name, s, ctl):
    N = s.addr.nCells
    ref = np.zeros(N)
    pr = orc.smooth_solve(s, ref, smoother="symGaussSeidel", **ctl)
    psi = np.zeros(N)
    perf = solver(gctx, s, smoother="symGaussSeidel", B200={"sweepMode": "exact"}, **ctl).solve(psi, s.source)
    assert perf.nIterations == pr.nIterations
    assert perf.converged == bool(pr.converged)
    assert np.array_equal(psi, ref)
    assert perf.initialResidual == pytest.approx(pr.initialResidual, rel=1e-12)
    assert perf.finalResidual == pytest.approx(pr.finalResidual, rel=1e-9)
    assert perf.normFactor == pytest.approx(pr.normFactor, rel=1e-12)
    assert str(perf).startswith("B200smoothSolver:  Solving for U, Initial residual = ")


@pytest.mark.parametrize("name,s", SYSTEMS, ids=IDS)
@pytest.mark.parametrize("smoother", ["GaussSeidel", "symGaussSeidel"])
def test_multicolour_solution_parity(gctx, name, s, smoother):
    N = s.addr.nCells
    ref = np.zeros(N)
    pr = orc.smooth_solve(s, ref, smoother=smoother, tolerance=1e-12, maxIter=3000)
    psi = np.zeros(N)
    perf = solver(gctx, s, smoother=smoother, tolerance=1e-12, maxIter=3000).solve(psi, s.source)
    assert pr.finalResidual < 1e-12 and perf.converged and perf.finalResidual < 1e-12
    assert perf.initialResidual == pytest.approx(pr.initialResidual, rel=1e-12)
    assert np.linalg.norm(psi - ref) <= 1e-8 * np.linalg.norm(ref)
    assert perf.nIterations <= 2 * pr.nIterations + 2
    # the reference's own controls: what it reports is the true residual of what it returns
    psi = np.zeros(N)
    perf = solver(gctx, s, smoother=smoother, tolerance=1e-6, maxIter=10).solve(psi, s.source)
    r = orc.residual_asym(s, psi)[0]
    # (rel 1e-6: the rows the iteration updated last contribute their rounding-level residual as formed in-kernel)
    assert perf.finalResidual == pytest.approx(np.abs(r).sum() / perf.normFactor, rel=1e-6)
    assert 1 <= perf.nIterations <= 10 and (perf.nIterations == 10 or perf.finalResidual < 1e-6)
    assert str(perf).startswith("B200smoothSolver(mc):  Solving for U")


def test_matches_the_kernel_transliteration(gctx):
    """multicolour mode against the numpy transliteration of the same kernels on the same plan"""
    _, s = SYSTEMS[1]
    pv = helpers.PlanView(1, s.addr, renumber=-1)
    N = s.addr.nCells
    got, n, init, final = helpers.smooth_solve_emulated(pv, s, np.zeros(N), tol=1e-9, maxIter=500)
    psi = np.zeros(N)
    perf = solver(gctx, s, smoother="symGaussSeidel", tolerance=1e-9, maxIter=500).solve(psi, s.source)
    assert perf.nIterations == n
    assert np.array_equal(psi, got)
    assert perf.finalResidual == pytest.approx(final, rel=1e-9)


def test_log_line_of_a_zero_system_and_error_paths(gctx):
    from firefoam_dev_b200 import B200Error
    _, s = SYSTEMS[0]
    N = s.addr.nCells
    z = System(s.addr, s.diag, s.upper, np.zeros(N), [], None, lower=s.lower)
    psi = np.zeros(N)
    for mode in ("exact", "multicolour"):
        perf = solver(gctx, z, smoother="symGaussSeidel", tolerance=1e-8, maxIter=10, B200={"sweepMode": mode}).solve(psi, z.source)
        # cases/steckler/original/linux64/log.fireFoam:176
        assert str(perf).endswith("Solving for U, Initial residual = 0, Final residual = 0, No Iterations 0")
        assert not psi.any()
    with pytest.raises(ValueError):
        solver(gctx, s, smoother="DILU")
    with pytest.raises(B200Error):
        solver(gctx, s, smoother="symGaussSeidel", nSweeps=0).solve(np.zeros(N), s.source)


def test_device_entry_point_and_rerun_reproducible(gctx):
    import torch
    from firefoam_dev_b200.ldu import make_smooth_controls
    _, s = SYSTEMS[1]
    N = s.addr.nCells
    gctx.set_addressing(s.addr)
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    d, up, lo, b = t(s.diag), t(s.upper), t(s.lower), t(s.source)
    ctl, _, _ = make_smooth_controls(dict(smoother="symGaussSeidel", tolerance=1e-9, maxIter=500))
    outs = []
    for _ in range(2):
        x = torch.zeros(N, dtype=torch.float64, device=dev)
        perf = gctx.smooth_solve_device(d, up, lo, [], b, x, ctl)
        torch.cuda.synchronize()
        outs.append((x.cpu().numpy(), perf.nIterations, perf.finalResidual))
    assert outs[0][1] == outs[1][1] and outs[0][2] == outs[1][2] and np.array_equal(outs[0][0], outs[1][0])
    psi = np.zeros(N)
    ph = solver(gctx, s, smoother="symGaussSeidel", tolerance=1e-9, maxIter=500).solve(psi, s.source)
    assert ph.nIterations == outs[0][1] and np.array_equal(psi, outs[0][0])


def test_lagged_residual_switch_gives_identical_iterates(gctx):
    """two-colour (hex) plans complete an iteration's residual in the first pass of the next one (default); with
    B200PCG_GS_LAGGED=0 a separate kernel evaluates it: same iterates, same sweep counts"""
    import os
    from firefoam_dev_b200 import Context
    old = os.environ.get("B200PCG_GS_LAGGED")
    os.environ["B200PCG_GS_LAGGED"] = "0"
    try:
        c2 = Context(device=0)
    finally:
        if old is None:
            os.environ.pop("B200PCG_GS_LAGGED", None)
        else:
            os.environ["B200PCG_GS_LAGGED"] = old
    try:
        for name in ("hex", "hex-odd", "steckler-topology"):
            s = dict(SYSTEMS)[name]
            N = s.addr.nCells
            for smoother in ("GaussSeidel", "symGaussSeidel"):
                for ctl in (dict(tolerance=1e-9, maxIter=500), dict(tolerance=1e-6, maxIter=10), dict(tolerance=1e-30, maxIter=7),
                            dict(tolerance=1e-30, maxIter=8, nSweeps=3), dict(tolerance=1e-3, minIter=4, maxIter=100)):
                    a, b = np.zeros(N), np.zeros(N)
                    pa = solver(gctx, s, smoother=smoother, **ctl).solve(a, s.source)
                    pb = solver(c2, s, smoother=smoother, **ctl).solve(b, s.source)
                    assert pa.nColours == 2
                    assert pa.nIterations == pb.nIterations and np.array_equal(a, b), (name, smoother, ctl)
                    assert pa.finalResidual == pytest.approx(pb.finalResidual, rel=1e-6)
                    assert pa.converged == pb.converged
    finally:
        c2.close()
