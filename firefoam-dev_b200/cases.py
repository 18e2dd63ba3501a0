"""Restated reference cases for the hot path (no OpenFOAM here, so the case files are turned into
LDU systems by hand).

StecklerHydrostatic: the hydrostatic-initialisation loop of solver/phrghEqn.H:19-60 on
cases/steckler (30 x 15 x 20 box, constant/polyMesh/blockMeshDict:50; compartment baffles from
system/topoSetDictCompartment + system/createBafflesDict), which is the reference's only pinned
instance of PCG results: cases/steckler/original/linux64/log.fireFoam:92-100.
"""
import numpy as np

from .ldu import LduAddressing, LduMatrix


class StecklerHydrostatic:
    NX, NY, NZ = 30, 15, 20
    H = 0.2
    # 0/T, 0/O2, 0/N2, constant/thermo.compressibleGas, constant/g, constant/hRef, constant/pRef
    T, Y_O2, Y_N2 = 298.15, 0.23301, 0.76699
    W_O2, W_N2, RR = 31.9988, 28.0134, 8314.47
    G, HREF, PREF = 9.81, 3.0, 101325.0
    # cases/steckler/system/fvSolution:43-46 (ph_rgh: $p_rgh -> tolerance 1e-6, relTol 0.01)
    TOL, RELTOL = 1e-6, 0.01
    N_CORR = 5  # nHydrostaticCorrectors, fvSolution:92

    def __init__(self):
        nx, ny, nz, h = self.NX, self.NY, self.NZ, self.H
        N = nx * ny * nz
        c = np.arange(N)
        i, j, k = c % nx, (c // nx) % ny, c // (nx * ny)
        inside = (i >= 3) & (i <= 16) & (j >= 0) & (j <= 10) & (k >= 3) & (k <= 16)
        faces = []
        for d, (di, dj, dk, stride) in enumerate(((1, 0, 0, 1), (0, 1, 0, nx), (0, 0, 1, nx * ny))):
            ok = (i + di < nx) & (j + dj < ny) & (k + dk < nz)
            own = c[ok]
            nei = own + stride
            baffle = inside[own] != inside[nei]
            if d == 0:  # doorway: faces on the x = 1.4 plane (between i=16 and i=17)
                door = (i[own] == 16) & (j[own] <= 4) & (k[own] >= 7) & (k[own] <= 12)
                baffle &= ~door
            own, nei = own[~baffle], nei[~baffle]
            faces.append(np.stack([own, nei, np.full(own.size, d)], axis=1))
        fa = np.concatenate(faces)
        order = np.lexsort((fa[:, 1], fa[:, 0]))   # upper-triangular order
        fa = fa[order]
        self.addr = LduAddressing(N, fa[:, 0].astype(np.int32), fa[:, 1].astype(np.int32))
        self.N, self.F = N, fa.shape[0]
        self.y_cell = (j + 0.5) * h
        yl, yu = self.y_cell[fa[:, 0]], self.y_cell[fa[:, 1]]
        self.ghf = self.G * (self.HREF - 0.5 * (yl + yu))
        self.gh = self.G * (self.HREF - self.y_cell)
        self.top = np.nonzero(j == ny - 1)[0]
        W = 1.0 / (self.Y_O2 / self.W_O2 + self.Y_N2 / self.W_N2)
        self.psi_thermo = W / (self.RR * self.T)
        self.rho0 = self.psi_thermo * self.PREF
        # p = ph_rgh + rho*gh + pRef; thermo.correct(); rho = thermo.rho()   (phrghEqn.H:21-23)
        self.ph_rgh = np.zeros(N)
        p = self.ph_rgh + self.rho0 * self.gh + self.PREF
        self.rho = self.psi_thermo * p

    def assemble(self, laplacian):
        """One corrector's ph_rghEqn: fvm::laplacian(rhof, ph_rgh) == fvc::div(phig)
        (phrghEqn.H:32-46).  `laplacian(gamma_f, magSf, deltaCoeffs, sign, diag0) -> (upper, diag)`
        is the assembly under test (oracle or CUDA)."""
        h = self.H
        l, u = self.addr.lowerAddr, self.addr.upperAddr
        rhof = 0.5 * (self.rho[l] + self.rho[u])
        snGrad = (self.rho[u] - self.rho[l]) / h
        phig = -rhof * self.ghf * snGrad * (h * h)
        # top patch fixedValue 0: internalCoeffs = -rho_b*magSf*deltaCoeffs_b, deltaCoeffs_b = 2/h
        diag0 = np.zeros(self.N)
        diag0[self.top] += -self.rho0 * (h * h) * (2.0 / h)
        upper, diag = laplacian(rhof, np.full(self.F, h * h), np.full(self.F, 1.0 / h), 1.0, diag0)
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
