"""Synthetic workloads and decomposePar-style partitioning: thin numpy wrappers over
libb200mesh.so (csrc/meshgen.cpp, csrc/decompose.cpp).  Harness code -- produces inputs only."""
import ctypes as C
from dataclasses import dataclass, field
from typing import List

import numpy as np

from . import _lib
from .ldu import LduAddressing, LduMatrix, ProcessorLduInterface

SEED = 20261018


@dataclass
class System:
    """What fvMatrix::solveSegregated hands to lduMatrix::solver on ONE rank."""
    addr: LduAddressing
    diag: np.ndarray          # with boundary internalCoeffs
    upper: np.ndarray
    source: np.ndarray        # totalSource
    bou: List[np.ndarray] = field(default_factory=list)   # interfaceBouCoeffs per coupled patch
    xstar: np.ndarray = None  # manufactured solution (source = A xstar), if any
    # inputs of fvm::laplacian for the same matrix
    gamma_f: np.ndarray = None
    magSf: np.ndarray = None
    deltaCoeffs: np.ndarray = None
    diag0: np.ndarray = None  # diagonal before the laplacian is added (ddt + boundary coeffs)
    sign: float = -1.0
    lower: np.ndarray = None  # asymmetric matrices only (lduMatrix::lower()); None: lower aliases upper

    @property
    def matrix(self):
        return LduMatrix(self.addr, self.diag, self.upper, self.lower)

    @property
    def interfaces(self):
        return list(self.addr.interfaces)


def hex_sizes(NX, NY, NZ, PX=1, PY=1, PZ=1, rank=0):
    L = _lib.load_mesh()
    sz = np.zeros(15, dtype=np.int64)
    if L.b200mesh_hex_sizes(NX, NY, NZ, PX, PY, PZ, rank, sz.ctypes.data) != 0:
        raise ValueError("bad hex block specification")
    return sz


def hex_block(NX, NY, NZ, PX=1, PY=1, PZ=1, rank=0, seed=SEED, h=None, gamma0=1e-3,
              psi=1.17e-5, dt=1e-3, pinned=None):
    """p_rgh-shaped system on rank `rank`'s block of the NX x NY x NZ box decomposed
    hierarchical/simple (PX PY PZ).  `pinned`: optional allocator f(n, dtype) -> ndarray used for
    the arrays the solver reads (so the bench can hand over page-locked host memory)."""
    L = _lib.load_mesh()
    sz = hex_sizes(NX, NY, NZ, PX, PY, PZ, rank)
    N, F, nIf = int(sz[0]), int(sz[1]), int(sz[2])
    if h is None:
        h = 1.0 / NX
    alloc = pinned or (lambda n, dt_: np.empty(n, dtype=dt_))
    lower = np.empty(F, dtype=np.int32)
    upper_addr = np.empty(F, dtype=np.int32)
    gamma_f = alloc(F, np.float64)
    magSf = alloc(F, np.float64)
    delta = alloc(F, np.float64)
    diag0 = alloc(N, np.float64)
    diag = alloc(N, np.float64)
    upper = alloc(F, np.float64)
    source = alloc(N, np.float64)
    xstar = np.empty(N, dtype=np.float64)
    fcs = [np.empty(int(sz[3 + k]), dtype=np.int32) for k in range(nIf)]
    bcs = [alloc(int(sz[3 + k]), np.float64) for k in range(nIf)]
    fc_arr = (C.c_void_p * 6)(*[a.ctypes.data for a in fcs] + [None] * (6 - nIf))
    bc_arr = (C.c_void_p * 6)(*[a.ctypes.data for a in bcs] + [None] * (6 - nIf))
    rc = L.b200mesh_hex_fill(NX, NY, NZ, PX, PY, PZ, rank, seed, float(h), float(gamma0),
                             float(psi / dt), lower.ctypes.data, upper_addr.ctypes.data,
                             gamma_f.ctypes.data, magSf.ctypes.data, delta.ctypes.data,
                             diag0.ctypes.data, diag.ctypes.data, upper.ctypes.data,
                             source.ctypes.data, xstar.ctypes.data, C.cast(fc_arr, C.c_void_p),
                             C.cast(bc_arr, C.c_void_p))
    if rc != 0:
        raise ValueError("hex_fill failed")
    ifs = [ProcessorLduInterface(int(sz[9 + k]), fcs[k], myProcNo=rank) for k in range(nIf)]
    addr = LduAddressing(N, lower, upper_addr, ifs)
    return System(addr, diag, upper, source, bcs, xstar, gamma_f, magSf, delta, diag0, -1.0)


def bcc_poly(nx, ny, nz, seed=SEED, a=None, gamma0=1e-3, psi=1.17e-5, dt=1e-3, shuffle_block=4096):
    """Polyhedral workload: BCC-lattice Voronoi mesh of truncated octahedra
    (2*nx*ny*nz cells, up to 14 faces per cell), Morton + block-shuffled numbering.  Returns a
    single-rank System; `.xyz` holds the cell centres (for the partitioners)."""
    L = _lib.load_mesh()
    sz = np.zeros(2, dtype=np.int64)
    if L.b200mesh_bcc_sizes(nx, ny, nz, seed, shuffle_block, sz.ctypes.data) != 0:
        raise ValueError("bad BCC specification")
    N, F = int(sz[0]), int(sz[1])
    if a is None:
        a = 1.0 / nx
    i32 = lambda n: np.empty(n, dtype=np.int32)
    f64 = lambda n: np.empty(n, dtype=np.float64)
    lower, upper_addr = i32(F), i32(F)
    gamma_f, magSf, delta, upper = f64(F), f64(F), f64(F), f64(F)
    diag0, diag, source, xstar, xyz = f64(N), f64(N), f64(N), f64(N), f64(3 * N)
    rc = L.b200mesh_bcc_fill(seed, float(a), float(gamma0), float(psi / dt), lower.ctypes.data,
                             upper_addr.ctypes.data, gamma_f.ctypes.data, magSf.ctypes.data,
                             delta.ctypes.data, diag0.ctypes.data, diag.ctypes.data, upper.ctypes.data,
                             source.ctypes.data, xstar.ctypes.data, xyz.ctypes.data)
    if rc != 0:
        raise ValueError("bcc_fill failed")
    s = System(LduAddressing(N, lower, upper_addr), diag, upper, source, [], xstar, gamma_f, magSf,
               delta, diag0, -1.0)
    s.xyz = xyz.reshape(N, 3)
    return s


# ---- partitioners + decomposition ----------------------------------------------------------------
def partition_simple(xyz, n):
    L = _lib.load_mesh()
    xyz = np.ascontiguousarray(xyz, dtype=np.float64)
    out = np.empty(xyz.shape[0], dtype=np.int32)
    L.b200mesh_partition_simple(xyz.shape[0], xyz.ctypes.data, n[0], n[1], n[2], out.ctypes.data)
    return out


def partition_hierarchical(xyz, n, order="xyz"):
    L = _lib.load_mesh()
    xyz = np.ascontiguousarray(xyz, dtype=np.float64)
    out = np.empty(xyz.shape[0], dtype=np.int32)
    o = np.array(["xyz".index(ch) for ch in order], dtype=np.int32)
    L.b200mesh_partition_hierarchical(xyz.shape[0], xyz.ctypes.data, n[0], n[1], n[2],
                                      o.ctypes.data, out.ctypes.data)
    return out


def partition_rcb(xyz, nProcs):
    L = _lib.load_mesh()
    xyz = np.ascontiguousarray(xyz, dtype=np.float64)
    out = np.empty(xyz.shape[0], dtype=np.int32)
    L.b200mesh_partition_rcb(xyz.shape[0], xyz.ctypes.data, int(nProcs), out.ctypes.data)
    return out


def decompose(system: System, cellToProc, nProcs):
    """Split a single-rank System into per-rank Systems with processor interfaces, as
    decomposePar + the processor fvPatchField coefficients would:
    interfaceBouCoeffs = -upper[cut face], interfaceIntCoeffs already inside diag."""
    L = _lib.load_mesh()
    a = system.addr
    if a.interfaces:
        raise ValueError("decompose expects a single-rank system")
    c2p = np.ascontiguousarray(cellToProc, dtype=np.int32)
    h = L.b200mesh_decompose(a.nCells, a.nFaces, a.lowerAddr.ctypes.data, a.upperAddr.ctypes.data,
                             c2p.ctypes.data, int(nProcs))
    if not h:
        raise ValueError("decompose failed (cellToProc out of range?)")

    def get(p, name):
        ptr = C.c_void_p()
        n = L.b200mesh_decompose_get(h, p, name.encode(), C.byref(ptr))
        if n <= 0:
            return np.empty(0, dtype=np.int32)
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_int32)), shape=(n,)).copy()

    out = []
    try:
        for p in range(nProcs):
            cells, faces = get(p, "cells"), get(p, "faces")
            ifNbr, ifStart = get(p, "ifNbr"), get(p, "ifStart")
            ifFC, ifGF = get(p, "ifFaceCells"), get(p, "ifGlobalFace")
            ifs, bou = [], []
            for k in range(ifNbr.size):
                s, e = int(ifStart[k]), int(ifStart[k + 1])
                ifs.append(ProcessorLduInterface(int(ifNbr[k]), ifFC[s:e], myProcNo=p))
                if system.lower is None:
                    bou.append(-system.upper[ifGF[s:e]])
                else:
                    # asymmetric: the row of the local cell carries upper[f] when that cell owns the cut face
                    # (A[l][u] = upper), lower[f] when it is the face's neighbour (A[u][l] = lower)
                    gf = ifGF[s:e]
                    owns = a.lowerAddr[gf] == cells[ifFC[s:e]]
                    bou.append(-np.where(owns, system.upper[gf], system.lower[gf]))
            addr = LduAddressing(cells.size, get(p, "lower"), get(p, "upper"), ifs)
            sub = System(addr, system.diag[cells].copy(), system.upper[faces].copy(),
                         system.source[cells].copy(), bou,
                         None if system.xstar is None else system.xstar[cells].copy())
            sub.cells, sub.faces = cells, faces
            if system.lower is not None:
                sub.lower = system.lower[faces].copy()
            if system.gamma_f is not None:
                # inputs of fvm::laplacian on the sub-mesh; the cut faces enter the sub-mesh diagonal
                # through the processor patches' internalCoeffs (= -upper of the cut face), which
                # solveSegregated adds to diag before the solver sees it
                sub.gamma_f, sub.magSf = system.gamma_f[faces].copy(), system.magSf[faces].copy()
                sub.deltaCoeffs, sub.sign = system.deltaCoeffs[faces].copy(), system.sign
                sub.diag0 = system.diag0[cells].copy()
                if ifFC.size:
                    np.subtract.at(sub.diag0, ifFC, system.upper[ifGF])
            out.append(sub)
    finally:
        L.b200mesh_decompose_free(h)
    return out
