"""Parity at BASELINE's full single-GPU size (configs[2]: 256 x 250 x 250 = 16 M hex cells) through
size-independent properties -- the oracle would need minutes there: symmetry and linearity of Amul,
Amul of the constant vector against an independent numpy row sum, the assembled coefficients against
numpy, the converged solution against the manufactured one and against a host-side residual, the
assembly -> solve chain kept on the device, and bitwise reproducibility of a re-run."""
import numpy as np
import pytest

from firefoam_dev_b200 import B200PCG, LduMatrix, meshgen as mg

pytestmark = pytest.mark.gpu

DIMS = (256, 250, 250)


@pytest.fixture(scope="module")
def big():
    from firefoam_dev_b200 import Context
    s = mg.hex_block(*DIMS)
    c = Context()
    c.set_addressing(s.addr)
    yield c, s
    c.close()


def host_amul(s, x):
    a = s.addr
    N = a.nCells
    y = s.diag * x
    y += np.bincount(a.upperAddr, s.upper * x[a.lowerAddr], N)
    y += np.bincount(a.lowerAddr, s.upper * x[a.upperAddr], N)
    return y


def test_amul_properties_16M(big):
    c, s = big
    N = s.addr.nCells
    rng = np.random.default_rng(16)
    x, y = rng.standard_normal(N), rng.standard_normal(N)
    Ax, Ay = c.amul(s.matrix, [], x), c.amul(s.matrix, [], y)
    # symmetry: (y, Ax) == (x, Ay)
    a, b = float(np.dot(y, Ax)), float(np.dot(x, Ay))
    assert abs(a - b) <= 1e-11 * (np.linalg.norm(y) * np.linalg.norm(Ax))
    # linearity
    Axy = c.amul(s.matrix, [], 0.5 * x - 2.0 * y)
    assert np.abs(Axy - (0.5 * Ax - 2.0 * Ay)).max() <= 1e-12 * np.abs(Axy).max()
    # against an independent host formulation (different summation order: tolerance, not bits)
    ref = host_amul(s, x)
    assert np.abs(Ax - ref).max() <= 1e-12 * np.abs(ref).max()
    # bitwise reproducible
    assert np.array_equal(Ax, c.amul(s.matrix, [], x))


def test_assembly_16M(big):
    c, s = big
    a = s.addr
    up, dg = c.assemble_laplacian(s.gamma_f, s.magSf, s.deltaCoeffs, -1.0, s.diag0)
    assert np.array_equal(up, -1.0 * (s.deltaCoeffs * (s.gamma_f * s.magSf)))     # element-wise: bit-exact
    dref = s.diag0 - (np.bincount(a.lowerAddr, up, a.nCells) + np.bincount(a.upperAddr, up, a.nCells))
    assert np.abs(dg - dref).max() <= 1e-13 * np.abs(dref).max()
    assert np.array_equal(up, s.upper) and np.array_equal(dg, s.diag)             # the generator's own (face-order) sums


def test_solve_16M(big):
    c, s = big
    ctl = {"preconditioner": "diagonal", "tolerance": 1e-6, "relTol": 0.0, "maxIter": 5000}
    psi = np.zeros(s.addr.nCells)
    perf = B200PCG("p_rgh", s.matrix, [], None, [], ctl, context=c).solve(psi, s.source)
    assert perf.converged and perf.finalResidual < 1e-6 and perf.initialResidual == 1.0
    assert np.abs(psi - s.xstar).max() < 1e-4 * np.abs(s.xstar).max()
    # OpenFOAM's residual definition recomputed on the host from the returned solution
    r = s.source - host_amul(s, psi)
    assert np.abs(r).sum() / perf.normFactor < 2e-6
    # re-run: same iteration count, same bits
    psi2 = np.zeros(s.addr.nCells)
    perf2 = B200PCG("p_rgh", s.matrix, [], None, [], ctl, context=c).solve(psi2, s.source)
    assert perf2.nIterations == perf.nIterations and np.array_equal(psi, psi2)
    # flux of the solution: sum over the faces of a cell telescopes to (A psi - diag psi)
    phi = c.flux(s.matrix, psi)
    a = s.addr
    div = np.bincount(a.lowerAddr, phi, a.nCells) - np.bincount(a.upperAddr, phi, a.nCells)
    offd = host_amul(s, psi) - s.diag * psi
    rowsum = np.bincount(a.lowerAddr, s.upper, a.nCells) + np.bincount(a.upperAddr, s.upper, a.nCells)
    assert np.abs(div - (offd - rowsum * psi)).max() <= 1e-10 * np.abs(offd).max()
