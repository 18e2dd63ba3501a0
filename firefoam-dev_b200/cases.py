"""Restated reference cases for the hot path (no OpenFOAM here, so the case files are turned into
LDU systems by hand).

HydrostaticBox        the hydrostatic-initialisation loop of solver/phrghEqn.H:19-60 on a uniform box
                      mesh with optional internal baffles.
StecklerHydrostatic   cases/steckler (30 x 15 x 20, constant/polyMesh/blockMeshDict:50; compartment
                      baffles from system/topoSetDictCompartment + system/createBafflesDict): the
                      reference's only pinned instance of PCG results,
                      cases/steckler/original/linux64/log.fireFoam:92-100.            (BASELINE config 2)
SingleBoxHydrostatic  the 7 x 5 x 7 base block of cases/singleBox/constant/polyMesh/blockMeshDict:53
                      with the BCs of cases/singleBox/0/ph_rgh.orig (snappyHexMesh refinement is not
                      reproducible here); PCG + diagonal plumbing case.                 (BASELINE config 1)
steckler_p_rgh_system synthetic p_rgh-shaped system on the steckler topology:
                      no per-time-step matrix dumps exist in the reference.
"""
import numpy as np

from .ldu import LduAddressing, LduMatrix


def _splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15))
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return x ^ (x >> np.uint64(31))


def uniform_pm1(seed, counters):
    """counter-based U(-1, 1) (splitmix64), same recipe as csrc/meshgen.cpp"""
    with np.errstate(over="ignore"):
        c = np.asarray(counters, dtype=np.uint64)
        r = _splitmix64(np.uint64(seed) ^ (c * np.uint64(0xD1342543DE82EF95) + np.uint64(0x2545F4914F6CDD1D)))
    return ((r >> np.uint64(11)).astype(np.float64) + 0.5) * (2.0 / 9007199254740992.0) - 1.0


class HydrostaticBox:
    # 0/T, 0/O2, 0/N2, constant/thermo.compressibleGas, constant/g, constant/pRef (same in both cases)
    T, Y_O2, Y_N2 = 298.15, 0.23301, 0.76699
    W_O2, W_N2, RR = 31.9988, 28.0134, 8314.47
    G, PREF = 9.81, 101325.0
    TOL, RELTOL = 1e-6, 0.01   # fvSolution ph_rgh { $p_rgh; }
    N_CORR = 5                 # nHydrostaticCorrectors

    def __init__(self, nx, ny, nz, dx, dy, dz, href, baffle=None):
        self.dims, self.d = (nx, ny, nz), (dx, dy, dz)
        self.HREF = href
        N = nx * ny * nz
        c = np.arange(N)
        i, j, k = c % nx, (c // nx) % ny, c // (nx * ny)
        self.ijk = (i, j, k)
        area = (dy * dz, dx * dz, dx * dy)
        dist = (dx, dy, dz)
        faces = []
        for d, (di, dj, dk, stride) in enumerate(((1, 0, 0, 1), (0, 1, 0, nx), (0, 0, 1, nx * ny))):
            ok = (i + di < nx) & (j + dj < ny) & (k + dk < nz)
            own = c[ok]
            nei = own + stride
            if baffle is not None:
                keep = ~baffle(d, own, nei, i, j, k)
                own, nei = own[keep], nei[keep]
            faces.append(np.stack([own, nei, np.full(own.size, d)], axis=1))
        fa = np.concatenate(faces)
        fa = fa[np.lexsort((fa[:, 1], fa[:, 0]))]     # upper-triangular order
        self.addr = LduAddressing(N, fa[:, 0].astype(np.int32), fa[:, 1].astype(np.int32))
        self.N, self.F = N, fa.shape[0]
        self.magSf = np.array(area)[fa[:, 2]]
        self.deltaCoeffs = 1.0 / np.array(dist)[fa[:, 2]]
        self.y_cell = (j + 0.5) * dy
        yl, yu = self.y_cell[fa[:, 0]], self.y_cell[fa[:, 1]]
        self.ghf = self.G * (self.HREF - 0.5 * (yl + yu))
        self.gh = self.G * (self.HREF - self.y_cell)
        self.top = np.nonzero(j == ny - 1)[0]
        W = 1.0 / (self.Y_O2 / self.W_O2 + self.Y_N2 / self.W_N2)
        self.psi_thermo = W / (self.RR * self.T)
        self.rho0 = self.psi_thermo * self.PREF
        # density on the `top` patch faces: thermo.rho() there is psi_b * p_b with p_b = pRef (gh_b = 0 at
        # y = hRef) and psi_b from the PATCH-FACE mixture (multiComponentMixture::patchFaceMixture sums
        # Y_i,b * specie_i over the boundary values of the Y fields as read from 0/)
        self.rho_top = self.top_patch_W() / (self.RR * self.T) * self.PREF
        # p = ph_rgh + rho*gh + pRef; thermo.correct(); rho = thermo.rho()   (phrghEqn.H:21-23)
        self.ph_rgh = np.zeros(N)
        p = self.ph_rgh + self.rho0 * self.gh + self.PREF
        self.rho = self.psi_thermo * p

    def top_patch_W(self):
        """molecular weight of the mixture on the `top` patch faces; cases/singleBox/0/N2:24-28 and
        0/O2:24-29 carry $internalField there: the cell mixture"""
        return 1.0 / (self.Y_O2 / self.W_O2 + self.Y_N2 / self.W_N2)

    def assemble(self, laplacian):
        """One corrector's ph_rghEqn: fvm::laplacian(rhof, ph_rgh) == fvc::div(phig)
        (phrghEqn.H:32-46).  `laplacian(gamma_f, magSf, deltaCoeffs, sign, diag0) -> (upper, diag)`
        is the assembly under test (oracle or CUDA)."""
        dy = self.d[1]
        l, u = self.addr.lowerAddr, self.addr.upperAddr
        rhof = 0.5 * (self.rho[l] + self.rho[u])
        snGrad = (self.rho[u] - self.rho[l]) * self.deltaCoeffs
        phig = -rhof * self.ghf * snGrad * self.magSf
        # top patch fixedValue 0: internalCoeffs = -rho_b*magSf_b*deltaCoeffs_b, deltaCoeffs_b = 2/dy,
        # rho_b = rho_top (patch-face mixture, see __init__);
        # the fixedFluxPressure patches contribute nothing (their gradient cancels fvc::div's
        # boundary flux)
        diag0 = np.zeros(self.N)
        diag0[self.top] += -self.rho_top * (self.d[0] * self.d[2]) * (2.0 / dy)
        upper, diag = laplacian(rhof, self.magSf, self.deltaCoeffs, 1.0, diag0)
        source = np.zeros(self.N)
        np.add.at(source, l, phig)
        np.subtract.at(source, u, phig)
        return LduMatrix(self.addr, diag, upper), source

    def update(self, ph_rgh):
        """p = ph_rgh + rho*gh + pRef; thermo.correct(); rho = thermo.rho()  (phrghEqn.H:50-52);
        returns gMax-gMin of ph_rgh (the log's 'Hydrostatic pressure variation')."""
        self.ph_rgh = ph_rgh
        p = ph_rgh + self.rho * self.gh + self.PREF
        self.rho = self.psi_thermo * p
        return float(ph_rgh.max() - ph_rgh.min())


class StecklerHydrostatic(HydrostaticBox):
    NX, NY, NZ, H = 30, 15, 20, 0.2

    def __init__(self):
        def baffle(d, own, nei, i, j, k):
            inside = (i >= 3) & (i <= 16) & (j >= 0) & (j <= 10) & (k >= 3) & (k <= 16)
            b = inside[own] != inside[nei]
            if d == 0:  # doorway: faces on the x = 1.4 plane (between i=16 and i=17)
                b &= ~((i[own] == 16) & (j[own] <= 4) & (k[own] >= 7) & (k[own] <= 12))
            return b
        super().__init__(self.NX, self.NY, self.NZ, self.H, self.H, self.H, 3.0, baffle)

    def top_patch_W(self):
        """cases/steckler/0/N2:24-28: the inert specie's `top` patch is `calculated; value uniform 0`
        (N2 is only re-derived as 1 - sum(Y_i) by the first YEEqn, after the hydrostatic loop), while
        0/O2:24-29 has 0.23301 there: during solver/phrghEqn.H:19-60 the patch-face mixture is O2 alone,
        W_b = W_O2, so rho_b(top) = 1.1091 rho_0.  This is the detail that makes the restated system
        reproduce log.fireFoam:92-100 to the printed digits (tools/kat/steckler_kat_candidates.py)."""
        return self.W_O2


class SingleBoxHydrostatic(HydrostaticBox):
    def __init__(self):
        inch = 0.0254   # convertToMeters, blockMeshDict:18; box 120 x 80 x 120 in, hRef 2.032 = 80 in
        super().__init__(7, 5, 7, 120 * inch / 7, 80 * inch / 5, 120 * inch / 7, 2.032)


def steckler_p_rgh_system(seed=1711, psi=1.17e-5, dt=0.0667):
    """A = diag(psi V/dt) - L(gamma) on the steckler topology,
    gamma_f = linear interpolate(rho*rAU), rho*rAU = dt (1 + 0.3 xi_c); `top` + `sides` fixed value,
    other patches zero flux; source = A x*, x* smooth + 1 % noise.  Returns meshgen.System."""
    from .meshgen import System
    case = StecklerHydrostatic()
    a, h = case.addr, case.H
    i, j, k = case.ijk
    nx, ny, nz = case.dims
    N = case.N
    g_c = dt * (1.0 + 0.3 * uniform_pm1(seed, np.arange(N)))
    l, u = a.lowerAddr, a.upperAddr
    gamma_f = 0.5 * (g_c[l] + g_c[u])
    magSf, delta = np.full(case.F, h * h), np.full(case.F, 1.0 / h)
    diag0 = np.full(N, psi * h ** 3 / dt)
    nb = (j == ny - 1).astype(float) + (i == 0) + (i == nx - 1) + (k == 0) + (k == nz - 1)
    diag0 += nb * g_c * (h * h) * (2.0 / h)
    upper = -1.0 * (delta * (gamma_f * magSf))
    # negSumDiag in face order (owner and neighbour interleaved, like OpenFOAM's loop)
    diag = np.zeros(N)
    np.subtract.at(diag, np.stack([l, u], 1).ravel(), np.repeat(upper, 2))
    diag = diag0 + diag
    x, y, z = (i + 0.5) * h - 2.0, (j + 0.5) * h, (k + 0.5) * h - 2.0
    xstar = np.sin(0.7 * x) * np.cos(1.1 * y) * np.sin(0.9 * z) + 0.01 * uniform_pm1(seed ^ 0x5EED, np.arange(N))
    source = diag * xstar
    np.add.at(source, u, upper * xstar[l])
    np.add.at(source, l, upper * xstar[u])
    return System(a, diag, upper, source, [], xstar, gamma_f, magSf, delta, diag0, -1.0)


def boundary_faces(case):
    """faceCells of every boundary face of a HydrostaticBox mesh (domain boundary + both sides of the
    baffles), as (cells, direction 0..2, deltaCoeffs_b, magSf_b), grouped direction by direction like
    patches."""
    nx, ny, nz = case.dims
    i, j, k = case.ijk
    a = case.addr
    N = case.N
    strides = (1, nx, nx * ny)
    have = set(zip(a.lowerAddr.tolist(), a.upperAddr.tolist()))
    cells, dirs = [], []
    for d, st in enumerate(strides):
        c = np.arange(N)
        up_missing = np.array([(cc, cc + st) not in have for cc in c.tolist()])
        lo_missing = np.array([(cc - st, cc) not in have for cc in c.tolist()])
        for miss in (lo_missing, up_missing):
            cs = c[miss]
            cells.append(cs)
            dirs.append(np.full(cs.size, d))
    cells, dirs = np.concatenate(cells).astype(np.int32), np.concatenate(dirs)
    dist = np.array(case.d)[dirs]
    area = np.array((case.d[1] * case.d[2], case.d[0] * case.d[2], case.d[0] * case.d[1]))[dirs]
    return cells, dirs, 2.0 / dist, area


def p1_G_terms(seed=238, case=None):
    """The P1 radiation model's G equation on the steckler topology, as the reference assembles it at
    packages/thermophysicalModels/radiation/radiationModels/P1/P1.C:238-244
        fvm::laplacian(gamma, G) - fvm::Sp(a, G) == - 4 e sigma T^4 - E,   gamma = 1/(3 a + 3 sigmaEff + a0)
    and solves it with `G { solver PCG; preconditioner DIC; tolerance 1e-06; relTol 0; }`
    (cases/steckler/system/fvSolution:75-81): the reference's OTHER symmetric lduMatrix system besides p_rgh
    (SURVEY.md 8f-4).  Expressed in the term slots of b200_assemble_p_rgh / orc_assemble_p_rgh -- the fvMatrix
    algebra is the same: fvm::Sp(a, G) puts V*a on the diagonal exactly where EulerDdtScheme puts
    rDeltaT*psi*V, so `- fvm::Sp(a, G)` is the ddt slot with rDeltaT = -1, psi = a, psi0 = 0; the volScalarField
    right-hand side is the Su slot; fvm::laplacian(volScalarField gamma, G) interpolates gamma linearly to the
    faces (0.5/0.5 on this uniform mesh); every wall carries a MarshakRadiation (mixed) condition:
        valueFraction = 1/(1 + gamma_b deltaCoeffs_b / Ep),  Ep = emissivity/(2 (2 - emissivity)),  refValue = 4 sigma T_w^4
        internalCoeffs = -gamma_b magSf valueFraction deltaCoeffs_b,  boundaryCoeffs = internalCoeffs * refValue.
    Synthetic but physically scaled fields: a = e in [0.05, 0.6] 1/m (sooty layer under the ceiling), T between 298
    and 1100 K (plume over the burner).  Returns (case, terms)."""
    case = case or StecklerHydrostatic()
    a_ = case.addr
    N, F = case.N, case.F
    i, j, k = case.ijk
    h = case.H if hasattr(case, "H") else case.d[0]
    x, y, z = (i + 0.5) * h - 2.0, (j + 0.5) * h, (k + 0.5) * h - 2.0
    r2 = x * x + z * z
    T = 298.15 + 800.0 * np.exp(-r2 / 0.18) * np.exp(-y / 1.6) + 150.0 * (y > 1.6) * (np.abs(x) < 1.4) * (np.abs(z) < 1.4)
    absorb = 0.05 + 0.55 * np.clip((T - 298.15) / 800.0, 0.0, 1.0) + 0.01 * uniform_pm1(seed, np.arange(N))
    sigmaSB = 5.670367e-08
    gamma_c = 1.0 / (3.0 * absorb)
    l, u = a_.lowerAddr, a_.upperAddr
    gamma_f = 0.5 * (gamma_c[l] - gamma_c[u]) + gamma_c[u]          # surfaceInterpolationScheme::interpolate, lambda = 0.5
    bCells, bDir, bDelta, bArea = boundary_faces(case)
    emissivity = 0.9
    Ep = emissivity / (2.0 * (2.0 - emissivity))
    gamma_b = gamma_c[bCells]                                       # zeroGradient-like patch value of gamma
    vf = 1.0 / (1.0 + gamma_b * bDelta / Ep)
    bInt = -(gamma_b * bArea) * (vf * bDelta)
    refValue = 4.0 * sigmaSB * 298.15 ** 4
    terms = {"rDeltaT": -1.0, "V": np.full(N, h ** 3), "psi": absorb, "psi0": np.zeros(N), "p0": np.zeros(N),
             "explicit": [], "gamma_f": gamma_f, "magSf": np.full(F, h * h), "deltaCoeffs": np.full(F, 1.0 / h),
             "lapSign": 1.0, "Su": -4.0 * (absorb * sigmaSB * T ** 4), "bCells": bCells,
             "bInternal": bInt, "bBoundary": bInt * refValue}
    return case, terms


def transport_system(base, peclet=2.0, kappa=0.15, upwind=0.9, seed=1930):
    """A U / Yi / h -shaped ASYMMETRIC lduMatrix on the topology of `base` (any meshgen.System): what
    fvMatrix::solveSegregated hands to `smoothSolver` for
        fvm::ddt(rho, U) + fvm::div(phi, U) - fvm::laplacian(muEff, U)        (solver/UEqn.H:19-30)
    restated term by term (OF-dev EulerDdtScheme.C, gaussConvectionScheme.C fvmDiv, gaussLaplacianScheme.C):
        div:        lower = -w phi;  upper = lower + phi;  negSumDiag         (w: owner-side interpolation weight)
        laplacian:  upper -= D;  lower -= D;  diag[l] += D;  diag[u] += D      (`- fvm::laplacian`)
        ddt:        diag += rho V / deltaT
    No per-time-step U / Yi / h matrices exist in the reference (they need the whole solver), so the fields are
    synthetic: D_f = |upper_f| of `base` (its diffusion coefficient), phi_f = peclet * D_f * U(-1, 1) (counter-based),
    w = 0.5 +- (upwind - 0.5) on the upwind side, rho V / deltaT = kappa * sum over the cell's faces of (D_f + |phi_f|)
    (kappa ~ 1 / Courant number: 0.15 makes symGaussSeidel need 10-30 sweeps, the shipped cases' time steps 1-4).
    Returns a System with .lower set, source = A xstar (xstar of `base`, or a seeded smooth + noise field), x0 = 0."""
    from .meshgen import System
    a = base.addr
    l, u = a.lowerAddr, a.upperAddr
    N, F = a.nCells, a.nFaces
    D = np.abs(np.asarray(base.upper, dtype=np.float64))
    phi = peclet * D * uniform_pm1(seed, np.arange(F))
    w = np.where(phi >= 0.0, upwind, 1.0 - upwind)
    lower = -w * phi
    upper = lower + phi
    # (np.bincount adds in index order of appearance, like the face loops; much faster than np.add.at at 48 M faces)
    acc = lambda idx, wts: np.bincount(idx, weights=wts, minlength=N)
    diag = -(acc(l, lower) + acc(u, upper))          # negSumDiag
    upper = upper - D
    lower = lower - D
    diag = diag + (acc(l, D) + acc(u, D))
    aphi = D + np.abs(phi)
    mass = acc(l, aphi) + acc(u, aphi)
    diag = diag + kappa * mass + 1e-300
    xstar = base.xstar if base.xstar is not None else 1.0 + 0.3 * uniform_pm1(seed + 1, np.arange(N))
    y = diag * xstar + acc(u, lower * xstar[l]) + acc(l, upper * xstar[u])
    return System(a, diag, upper, y, [], xstar, lower=lower)
