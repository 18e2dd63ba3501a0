"""TEST INFRASTRUCTURE: ctypes binding of oracle/liboracle.so (pcg_oracle.c), the CPU restatement
of OpenFOAM-dev's PCG / DIC / Amul / laplacian path the reference calls (see pcg_oracle.c header
for the citations and the pinning status).  Import ONLY from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "liboracle.so")


def build():
    r = subprocess.run(["make", "-C", _HERE], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)


class _Iface(C.Structure):
    _fields_ = [("nbrRank", C.c_int32), ("nFaces", C.c_int32), ("faceCells", C.c_void_p),
                ("bouCoeffs", C.c_void_p)]


class _Rank(C.Structure):
    _fields_ = [("nCells", C.c_int32), ("nFaces", C.c_int32), ("lower", C.c_void_p),
                ("upper", C.c_void_p), ("diag", C.c_void_p), ("upperCoeffs", C.c_void_p),
                ("source", C.c_void_p), ("psi", C.c_void_p), ("nIfaces", C.c_int32),
                ("ifaces", C.c_void_p)]


class _Controls(C.Structure):
    _fields_ = [("tolerance", C.c_double), ("relTol", C.c_double), ("maxIter", C.c_int32),
                ("minIter", C.c_int32), ("precond", C.c_int32), ("unused", C.c_int32)]


class OrcPerf(C.Structure):
    _fields_ = [("initialResidual", C.c_double), ("finalResidual", C.c_double),
                ("normFactor", C.c_double), ("nIterations", C.c_int32), ("converged", C.c_int32),
                ("singular", C.c_int32), ("pad", C.c_int32)]


_L = None


def lib():
    global _L
    if _L is None:
        if not os.path.exists(SO):
            build()
        _L = C.CDLL(SO)
        _L.orc_laplacian_assemble.argtypes = [C.c_int32, C.c_int32] + [C.c_void_p] * 5 + [C.c_double] + [C.c_void_p] * 2
        _L.orc_laplacian_assemble.restype = None
        _L.orc_flux.argtypes = [C.c_int32] + [C.c_void_p] * 5
        _L.orc_flux.restype = None
        _L.orc_pcg_solve.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        _L.orc_amul.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        _L.orc_sumA.argtypes = [C.c_void_p, C.c_void_p]
        _L.orc_sumA.restype = None
        _L.orc_dic_calc_rd.argtypes = [C.c_void_p, C.c_void_p]
        _L.orc_dic_calc_rd.restype = None
        _L.orc_dic_precondition.argtypes = [C.c_void_p] * 4
        _L.orc_dic_precondition.restype = None
    return _L


PRECOND = {"none": 0, "diagonal": 1, "DIC": 2}


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def laplacian_assemble(lowerAddr, upperAddr, nCells, gamma_f, magSf, deltaCoeffs, sign, diag0=None):
    l, u = _i32(lowerAddr), _i32(upperAddr)
    g, s, d = _f64(gamma_f), _f64(magSf), _f64(deltaCoeffs)
    upper = np.empty(l.size, dtype=np.float64)
    diag = np.zeros(nCells, dtype=np.float64) if diag0 is None else _f64(diag0).copy()
    lib().orc_laplacian_assemble(nCells, l.size, l.ctypes.data, u.ctypes.data, g.ctypes.data,
                                 s.ctypes.data, d.ctypes.data, float(sign), upper.ctypes.data,
                                 diag.ctypes.data)
    return upper, diag


class PrghTerms(C.Structure):
    """orc_prgh_terms (pcg_oracle.c) == b200_prgh_terms (include/b200pcg.h): same layout."""
    _fields_ = [("rDeltaT", C.c_double), ("V", C.c_void_p), ("psi", C.c_void_p), ("psi0", C.c_void_p),
                ("p0", C.c_void_p), ("nExplicit", C.c_int32), ("pad0", C.c_int32),
                ("explicitFields", C.c_void_p), ("phi", C.c_void_p), ("divSign", C.c_double),
                ("gamma_f", C.c_void_p), ("magSf", C.c_void_p), ("deltaCoeffs", C.c_void_p),
                ("lapSign", C.c_double), ("Su", C.c_void_p), ("nB", C.c_int32), ("pad1", C.c_int32),
                ("bCells", C.c_void_p), ("bPhi", C.c_void_p), ("bInternal", C.c_void_p),
                ("bBoundary", C.c_void_p)]


def pack_prgh_terms(terms, ptr=lambda a: a.ctypes.data):
    """dict -> (PrghTerms, keep-alive list).  Keys: rDeltaT, V, psi, psi0, p0, explicit (list), phi,
    divSign, gamma_f, magSf, deltaCoeffs, lapSign, Su, bCells, bPhi, bInternal, bBoundary (absent = NULL).
    `ptr` maps an array to its address (device tensors: lambda t: t.data_ptr())."""
    t = PrghTerms()
    keep = []

    def put(name, key, dtype=np.float64):
        v = terms.get(key)
        if v is None:
            return
        if isinstance(v, np.ndarray) or isinstance(v, (list, tuple)):
            v = np.ascontiguousarray(v, dtype=dtype)
        keep.append(v)
        setattr(t, name, ptr(v))
    t.rDeltaT = float(terms.get("rDeltaT", 0.0))
    for n in ("V", "psi", "psi0", "p0", "phi", "gamma_f", "magSf", "deltaCoeffs", "Su", "bPhi", "bInternal", "bBoundary"):
        put(n, n)
    put("bCells", "bCells", np.int32)
    ex = terms.get("explicit") or []
    ex = [np.ascontiguousarray(e, dtype=np.float64) if isinstance(e, (np.ndarray, list, tuple)) else e for e in ex]
    arr = (C.c_void_p * max(1, len(ex)))(*[ptr(e) for e in ex])
    keep += [ex, arr]
    t.nExplicit = len(ex)
    t.explicitFields = C.cast(arr, C.c_void_p)
    t.divSign = float(terms.get("divSign", -1.0))
    t.lapSign = float(terms.get("lapSign", -1.0))
    bc = terms.get("bCells")
    t.nB = 0 if bc is None else int(len(bc))
    return t, keep


def assemble_p_rgh(lowerAddr, upperAddr, nCells, terms):
    """Restatement of the p_rghEqn assembly + solveSegregated boundary fold; returns (upper, diag, source)."""
    l, u = _i32(lowerAddr), _i32(upperAddr)
    t, keep = pack_prgh_terms(terms)
    upper, diag, src = np.empty(l.size), np.empty(nCells), np.empty(nCells)
    L = lib()
    L.orc_assemble_p_rgh.argtypes = [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p]
    L.orc_assemble_p_rgh.restype = None
    L.orc_assemble_p_rgh(nCells, l.size, l.ctypes.data, u.ctypes.data, C.addressof(t), upper.ctypes.data,
                         diag.ctypes.data, src.ctypes.data)
    return upper, diag, src


def flux(lowerAddr, upperAddr, upper, psi):
    l, u, a, x = _i32(lowerAddr), _i32(upperAddr), _f64(upper), _f64(psi)
    out = np.empty(l.size, dtype=np.float64)
    lib().orc_flux(l.size, l.ctypes.data, u.ctypes.data, a.ctypes.data, x.ctypes.data, out.ctypes.data)
    return out


class _Packed:
    """keeps the numpy arrays alive while the C structs point at them"""

    def __init__(self, systems, psis):
        self.keep = []
        R = len(systems)
        self.ranks = (_Rank * R)()
        for r, (s, psi) in enumerate(zip(systems, psis)):
            a = s.addr
            arrs = dict(l=_i32(a.lowerAddr), u=_i32(a.upperAddr), d=_f64(s.diag), up=_f64(s.upper),
                        b=_f64(s.source))
            ifs = (_Iface * max(1, len(a.interfaces)))()
            for k, itf in enumerate(a.interfaces):
                fc, bc = _i32(itf.faceCells), _f64(s.bou[k])
                self.keep += [fc, bc]
                ifs[k].nbrRank, ifs[k].nFaces = itf.neighbProcNo, fc.size
                ifs[k].faceCells, ifs[k].bouCoeffs = fc.ctypes.data, bc.ctypes.data
            self.keep += [arrs, ifs, psi]
            R_ = self.ranks[r]
            R_.nCells, R_.nFaces = a.nCells, a.nFaces
            R_.lower, R_.upper = arrs["l"].ctypes.data, arrs["u"].ctypes.data
            R_.diag, R_.upperCoeffs = arrs["d"].ctypes.data, arrs["up"].ctypes.data
            R_.source, R_.psi = arrs["b"].ctypes.data, psi.ctypes.data
            R_.nIfaces, R_.ifaces = len(a.interfaces), C.cast(ifs, C.c_void_p)


def pcg_solve(systems, psis, preconditioner="diagonal", tolerance=1e-6, relTol=0.0, maxIter=1000,
              minIter=0):
    """PCG on len(systems) emulated ranks.  systems: objects with .addr (LduAddressing-like with
    .interfaces), .diag, .upper, .source, .bou; psis: list of float64 arrays (updated in place).
    A single system may be passed bare."""
    if not isinstance(systems, (list, tuple)):
        systems, psis = [systems], [psis]
    for p in psis:
        assert p.dtype == np.float64 and p.flags.c_contiguous
    pk = _Packed(systems, psis)
    ctl = _Controls(tolerance, relTol, maxIter, minIter, PRECOND[preconditioner], 0)
    perf = OrcPerf()
    rc = lib().orc_pcg_solve(len(systems), C.cast(pk.ranks, C.c_void_p), C.byref(ctl), C.byref(perf))
    if rc != 0:
        raise RuntimeError(f"oracle: interfaces do not pair up (rc={rc})")
    return perf


def amul(systems, xs):
    if not isinstance(systems, (list, tuple)):
        systems, xs = [systems], [xs]
    xs = [_f64(x) for x in xs]
    ys = [np.empty_like(x) for x in xs]
    pk = _Packed(systems, [x.copy() for x in xs])
    xa = (C.c_void_p * len(xs))(*[x.ctypes.data for x in xs])
    ya = (C.c_void_p * len(ys))(*[y.ctypes.data for y in ys])
    rc = lib().orc_amul(len(systems), C.cast(pk.ranks, C.c_void_p), C.cast(xa, C.c_void_p),
                        C.cast(ya, C.c_void_p))
    if rc != 0:
        raise RuntimeError(f"oracle: interfaces do not pair up (rc={rc})")
    return ys


def dic(system, r):
    """(rD, w) of DICPreconditioner for one rank: calcReciprocalD and precondition(w, r)."""
    pk = _Packed([system], [np.zeros(system.addr.nCells)])
    rD = np.empty(system.addr.nCells)
    w = np.empty(system.addr.nCells)
    r = _f64(r)
    lib().orc_dic_calc_rd(C.byref(pk.ranks[0]), rD.ctypes.data)
    lib().orc_dic_precondition(C.byref(pk.ranks[0]), rD.ctypes.data, w.ctypes.data, r.ctypes.data)
    return rD, w


def sumA(system):
    pk = _Packed([system], [np.zeros(system.addr.nCells)])
    s = np.empty(system.addr.nCells)
    lib().orc_sumA(C.byref(pk.ranks[0]), s.ctypes.data)
    return s


# ---- smooth_oracle.c: smoothSolver + GaussSeidel / symGaussSeidel on asymmetric lduMatrices ----
class _SmRank(C.Structure):
    _fields_ = [("nCells", C.c_int32), ("nFaces", C.c_int32), ("lower", C.c_void_p), ("upper", C.c_void_p),
                ("diag", C.c_void_p), ("upperCoeffs", C.c_void_p), ("lowerCoeffs", C.c_void_p),
                ("source", C.c_void_p), ("psi", C.c_void_p), ("nIfaces", C.c_int32), ("ifaces", C.c_void_p)]


class _SmControls(C.Structure):
    _fields_ = [("tolerance", C.c_double), ("relTol", C.c_double), ("maxIter", C.c_int32),
                ("minIter", C.c_int32), ("nSweeps", C.c_int32), ("smoother", C.c_int32)]


SMOOTHER = {"GaussSeidel": 0, "symGaussSeidel": 1}


class _SmPacked:
    """systems: objects with .addr, .diag, .upper, .source, .bou and (optionally) .lower -- None or absent
    means a symmetric matrix (lower aliases upper, as in lduMatrix::lower())."""

    def __init__(self, systems, psis):
        self.keep = []
        R = len(systems)
        self.ranks = (_SmRank * R)()
        for r, (s, psi) in enumerate(zip(systems, psis)):
            a = s.addr
            low = getattr(s, "lower", None)
            arrs = dict(l=_i32(a.lowerAddr), u=_i32(a.upperAddr), d=_f64(s.diag), up=_f64(s.upper),
                        b=_f64(s.source))
            arrs["lo"] = arrs["up"] if low is None else _f64(low)
            ifs = (_Iface * max(1, len(a.interfaces)))()
            for k, itf in enumerate(a.interfaces):
                fc, bc = _i32(itf.faceCells), _f64(s.bou[k])
                self.keep += [fc, bc]
                ifs[k].nbrRank, ifs[k].nFaces = itf.neighbProcNo, fc.size
                ifs[k].faceCells, ifs[k].bouCoeffs = fc.ctypes.data, bc.ctypes.data
            self.keep += [arrs, ifs, psi]
            R_ = self.ranks[r]
            R_.nCells, R_.nFaces = a.nCells, a.nFaces
            R_.lower, R_.upper = arrs["l"].ctypes.data, arrs["u"].ctypes.data
            R_.diag, R_.upperCoeffs, R_.lowerCoeffs = arrs["d"].ctypes.data, arrs["up"].ctypes.data, arrs["lo"].ctypes.data
            R_.source, R_.psi = arrs["b"].ctypes.data, psi.ctypes.data
            R_.nIfaces, R_.ifaces = len(a.interfaces), C.cast(ifs, C.c_void_p)


def smooth_solve(systems, psis, smoother="symGaussSeidel", tolerance=1e-6, relTol=0.0, maxIter=1000, minIter=0,
                 nSweeps=1):
    """smoothSolver::solve on len(systems) emulated ranks; psis updated in place.  Returns OrcPerf."""
    if not isinstance(systems, (list, tuple)):
        systems, psis = [systems], [psis]
    for p in psis:
        assert p.dtype == np.float64 and p.flags.c_contiguous
    pk = _SmPacked(systems, psis)
    ctl = _SmControls(tolerance, relTol, maxIter, minIter, nSweeps, SMOOTHER[smoother])
    perf = OrcPerf()
    L = lib()
    L.orc_smooth_solve.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    rc = L.orc_smooth_solve(len(systems), C.cast(pk.ranks, C.c_void_p), C.byref(ctl), C.byref(perf))
    if rc != 0:
        raise RuntimeError(f"oracle: interfaces do not pair up (rc={rc})")
    return perf


def _asym_apply(systems, xs, what):
    if not isinstance(systems, (list, tuple)):
        systems, xs = [systems], [xs]
    xs = [_f64(x) for x in xs]
    ys = [np.empty_like(x) for x in xs]
    pk = _SmPacked(systems, [x.copy() for x in xs])
    xa = (C.c_void_p * len(xs))(*[x.ctypes.data for x in xs])
    ya = (C.c_void_p * len(ys))(*[y.ctypes.data for y in ys])
    L = lib()
    L.orc_asym_apply.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    rc = L.orc_asym_apply(len(systems), C.cast(pk.ranks, C.c_void_p), C.cast(xa, C.c_void_p),
                          C.cast(ya, C.c_void_p), what)
    if rc != 0:
        raise RuntimeError(f"oracle: interfaces do not pair up (rc={rc})")
    return ys


def amul_asym(systems, xs):
    """lduMatrix::Amul with lower != upper (list per rank)."""
    return _asym_apply(systems, xs, 0)


def residual_asym(systems, xs):
    """lduMatrix::residual: source - A x with the coupled coefficients negated (list per rank)."""
    return _asym_apply(systems, xs, 1)


def sumA_asym(system):
    pk = _SmPacked([system], [np.zeros(system.addr.nCells)])
    s = np.empty(system.addr.nCells)
    L = lib()
    L.orc_asym_sumA.argtypes = [C.c_void_p, C.c_void_p]
    L.orc_asym_sumA.restype = None
    L.orc_asym_sumA(C.byref(pk.ranks[0]), s.ctypes.data)
    return s


# ---- bicg_oracle.c: PBiCG + DILU / diagonal / none on asymmetric lduMatrices, one rank ----
class _BiSys(C.Structure):
    _fields_ = [("nCells", C.c_int32), ("nFaces", C.c_int32), ("lower", C.c_void_p), ("upper", C.c_void_p),
                ("diag", C.c_void_p), ("upperCoeffs", C.c_void_p), ("lowerCoeffs", C.c_void_p),
                ("source", C.c_void_p), ("psi", C.c_void_p)]


BICG_PRECOND = {"none": 0, "diagonal": 1, "DILU": 2}


def _bi_pack(system, psi):
    a = system.addr
    low = getattr(system, "lower", None)
    arrs = dict(l=_i32(a.lowerAddr), u=_i32(a.upperAddr), d=_f64(system.diag), up=_f64(system.upper), b=_f64(system.source))
    arrs["lo"] = arrs["up"] if low is None else _f64(low)
    s = _BiSys(a.nCells, a.nFaces, arrs["l"].ctypes.data, arrs["u"].ctypes.data, arrs["d"].ctypes.data,
               arrs["up"].ctypes.data, arrs["lo"].ctypes.data, arrs["b"].ctypes.data, psi.ctypes.data)
    return s, (arrs, psi)


def pbicg_solve(system, psi, preconditioner="DILU", tolerance=1e-6, relTol=0.0, maxIter=1000, minIter=0):
    """PBiCG::solve on one rank; psi (float64, contiguous) updated in place.  Returns OrcPerf."""
    assert psi.dtype == np.float64 and psi.flags.c_contiguous
    if getattr(system.addr, "interfaces", None):
        raise ValueError("bicg_oracle.c is a one-rank restatement")
    s, keep = _bi_pack(system, psi)
    ctl = _Controls(tolerance, relTol, maxIter, minIter, BICG_PRECOND[preconditioner], 0)
    perf = OrcPerf()
    L = lib()
    L.orc_pbicg_solve.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.orc_pbicg_solve(C.byref(s), C.byref(ctl), C.byref(perf))
    return perf


def tmul_asym(system, x):
    x = _f64(x)
    y = np.empty_like(x)
    s, keep = _bi_pack(system, np.zeros(system.addr.nCells))
    L = lib()
    L.orc_asym_tmul.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.orc_asym_tmul.restype = None
    L.orc_asym_tmul(C.byref(s), x.ctypes.data, y.ctypes.data)
    return y


def dilu(system, r):
    """(rD, w, wT) of DILUPreconditioner: calcReciprocalD, precondition(w, r), preconditionT(wT, r)."""
    r = _f64(r)
    n = system.addr.nCells
    rD, w, wT = np.empty(n), np.empty(n), np.empty(n)
    s, keep = _bi_pack(system, np.zeros(n))
    L = lib()
    L.orc_dilu.argtypes = [C.c_void_p] * 5
    L.orc_dilu.restype = None
    L.orc_dilu(C.byref(s), r.ctypes.data, rD.ctypes.data, w.ctypes.data, wT.ctypes.data)
    return rD, w, wT
