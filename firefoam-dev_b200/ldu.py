"""Host-side mirror of the OpenFOAM interfaces the hot path sits behind.

Names and argument meaning follow the un-vendored OpenFOAM-dev classes the reference calls
through `p_rghEqn.solve(...)` (solver/pEqn.H:39, solver/phrghEqn.H:48):
lduAddressing, lduMatrix, processorLduInterface, lduMatrix::solver(::New), SolverPerformance
(SURVEY.md 8a/8b).  Everything numerical is delegated to libb200pcg.so through the C ABI of
include/b200pcg.h -- this file holds no arithmetic.
"""
import ctypes as C
import itertools

import numpy as np

from . import _lib
from ._lib import B200Error, BICG_PRECOND, Controls, Iface, Perf, PRECOND, SMOOTHER, SWEEP_MODE, SmoothControls

_mesh_keys = itertools.count(1)


def _f64(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class ProcessorLduInterface:
    """processorLduInterface: a coupled patch to one neighbour rank."""

    def __init__(self, neighbProcNo, faceCells, myProcNo=0, tag=0):
        self.neighbProcNo = int(neighbProcNo)
        self.myProcNo = int(myProcNo)
        self.faceCells = _i32(faceCells)
        self.tag = int(tag)


class LduAddressing:
    """lduAddressing: lowerAddr (owner), upperAddr (neighbour) in upper-triangular order,
    patchAddr(k) = faceCells of coupled patch k."""

    def __init__(self, nCells, lowerAddr, upperAddr, interfaces=()):
        self.nCells = int(nCells)
        self.lowerAddr = _i32(lowerAddr)
        self.upperAddr = _i32(upperAddr)
        if self.lowerAddr.shape != self.upperAddr.shape:
            raise ValueError("lowerAddr and upperAddr differ in size")
        self.interfaces = list(interfaces)
        self.mesh_key = next(_mesh_keys)

    @property
    def nFaces(self):
        return self.lowerAddr.size

    def patchAddr(self, k):
        return self.interfaces[k].faceCells


class LduMatrix:
    """lduMatrix: diag [nCells], upper [nFaces] and, for an asymmetric matrix, lower [nFaces]
    (lower is None: symmetric, lower aliases upper -- lduMatrix::lower())."""

    def __init__(self, lduAddr, diag, upper, lower=None):
        self.lduAddr = lduAddr
        self.diag = _f64(diag)
        self.upper = _f64(upper)
        self.lower = None if lower is None else _f64(lower)
        if self.diag.size != lduAddr.nCells or self.upper.size != lduAddr.nFaces:
            raise ValueError("diag/upper size does not match the addressing")
        if self.lower is not None and self.lower.size != lduAddr.nFaces:
            raise ValueError("lower size does not match the addressing")

    def symmetric(self):
        return self.lower is None

    def asymmetric(self):
        return self.lower is not None


class SolverPerformance:
    """SolverPerformance<scalar>; str() is the log line of
    cases/steckler/original/linux64/log.fireFoam:92."""

    def __init__(self, solverName, fieldName, perf):
        self.solverName = solverName
        self.fieldName = fieldName
        self.initialResidual = perf.initialResidual
        self.finalResidual = perf.finalResidual
        self.nIterations = perf.nIterations
        self.converged = bool(perf.converged)
        self.singular = bool(perf.singular)
        self.normFactor = getattr(perf, "normFactor", float("nan"))
        self.nColours = getattr(perf, "nColours", 0)
        self.solveMs = getattr(perf, "solveMs", 0.0)
        self.setupMs = getattr(perf, "setupMs", 0.0)
        self.h2dMs = getattr(perf, "h2dMs", 0.0)
        self.d2hMs = getattr(perf, "d2hMs", 0.0)

    def __str__(self):
        s = (f"{self.solverName}:  Solving for {self.fieldName}, Initial residual = "
             f"{self.initialResidual:.8g}, Final residual = {self.finalResidual:.8g}, "
             f"No Iterations {self.nIterations}")
        if self.singular:
            s = f"{self.solverName}:  Solving for {self.fieldName}:  solution singularity"
        return s


class Context:
    """One b200_ctx: one rank on one GPU (streams, NCCL communicator, cached device state per
    mesh).  OpenFOAM constructs a new lduMatrix::solver per solve, so the context -- not the
    solver object -- owns everything that must persist."""

    def __init__(self, device=-1, rank=0, nranks=1, nccl_uid=None):
        if nranks > 1:
            # make sure the process-wide NCCL is the one PyTorch bundles (the library binds
            # NCCL lazily with dlopen and picks up whichever libnccl.so.2 is already loaded)
            import torch  # noqa: F401
        self.lib = _lib.load_pcg()
        self.handle = C.c_void_p()
        uid = None
        if nccl_uid is not None:
            uid = (C.c_char * 128).from_buffer_copy(bytes(nccl_uid))
        rc = self.lib.b200_ctx_create(device, rank, nranks, uid, C.byref(self.handle))
        if rc != 0:
            raise B200Error(rc, self.lib.b200_last_error(None).decode())
        self.rank, self.nranks = rank, nranks
        self._addr = None

    @staticmethod
    def unique_id():
        import torch  # noqa: F401  (see __init__)
        lib = _lib.load_pcg()
        buf = (C.c_char * 128)()
        rc = lib.b200_get_unique_id(buf)
        if rc != 0:
            raise B200Error(rc, lib.b200_last_error(None).decode())
        return bytes(buf)

    def _check(self, rc):
        if rc != 0:
            raise B200Error(rc, self.lib.b200_last_error(self.handle).decode())

    def close(self):
        if self.handle:
            self.lib.b200_ctx_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- lduAddressing ---------------------------------------------------------------------
    def set_addressing(self, addr):
        ifs = (Iface * max(1, len(addr.interfaces)))()
        for k, itf in enumerate(addr.interfaces):
            ifs[k].nbrRank = itf.neighbProcNo
            ifs[k].nFaces = itf.faceCells.size
            ifs[k].faceCells = itf.faceCells.ctypes.data_as(C.POINTER(C.c_int32))
            ifs[k].tag = itf.tag
        self._check(self.lib.b200_set_addressing(
            self.handle, addr.mesh_key, addr.nCells, addr.nFaces, addr.lowerAddr.ctypes.data,
            addr.upperAddr.ctypes.data, len(addr.interfaces), C.cast(ifs, C.c_void_p)))
        self._addr = addr

    # ---- fvm::laplacian --------------------------------------------------------------------
    def assemble_laplacian(self, gamma_f, magSf, deltaCoeffs, sign, diag_inout=None):
        addr = self._addr
        g, s, d = _f64(gamma_f), _f64(magSf), _f64(deltaCoeffs)
        upper = np.empty(addr.nFaces, dtype=np.float64)
        diag = np.zeros(addr.nCells, dtype=np.float64) if diag_inout is None else _f64(diag_inout).copy()
        self._check(self.lib.b200_assemble_laplacian(self.handle, g.ctypes.data, s.ctypes.data,
                                                     d.ctypes.data, float(sign), upper.ctypes.data,
                                                     diag.ctypes.data))
        return upper, diag

    def assemble_laplacian_device(self, d_gamma, d_magSf, d_delta, sign, d_upper_out, d_diag_inout):
        self._check(self.lib.b200_assemble_laplacian_device(
            self.handle, _ptr(d_gamma), _ptr(d_magSf), _ptr(d_delta), float(sign), _ptr(d_upper_out),
            _ptr(d_diag_inout)))

    # ---- the whole p_rghEqn (solver/pEqn.H:26-37) + solveSegregated boundary fold ---------------
    def set_boundary_faces(self, bCells):
        b = _i32(bCells)
        self._check(self.lib.b200_set_boundary_faces(self.handle, b.size, b.ctypes.data))

    def _prgh_terms(self, terms, device):
        """dict -> b200_prgh_terms.  Keys: rDeltaT, V, psi, psi0, p0, explicit (list of cell fields),
        phi, divSign, gamma_f, magSf, deltaCoeffs, lapSign, Su, bCells, bPhi, bInternal, bBoundary."""
        t = _lib.PrghTerms()
        keep = []

        def addr(v, dtype=np.float64):
            if v is None:
                return None
            if device:
                keep.append(v)
                return _ptr(v)
            a = np.ascontiguousarray(v, dtype=dtype)
            keep.append(a)
            return a.ctypes.data
        t.rDeltaT = float(terms.get("rDeltaT", 0.0))
        for n in ("V", "psi", "psi0", "p0", "phi", "gamma_f", "magSf", "deltaCoeffs", "Su", "bPhi",
                  "bInternal", "bBoundary"):
            setattr(t, n, addr(terms.get(n)))
        ex = list(terms.get("explicit") or [])
        arr = (C.c_void_p * max(1, len(ex)))(*[addr(e) for e in ex])
        keep.append(arr)
        t.nExplicit, t.explicitFields = len(ex), C.cast(arr, C.c_void_p)
        t.divSign, t.lapSign = float(terms.get("divSign", -1.0)), float(terms.get("lapSign", -1.0))
        bc = terms.get("bCells")
        if bc is not None:
            bc = _i32(bc)
            keep.append(bc)
            t.nB, t.bCells = bc.size, bc.ctypes.data
        return t, keep

    def assemble_p_rgh(self, terms):
        """Host arrays in, (upper, diag with internalCoeffs, totalSource) out."""
        addr = self._addr
        t, keep = self._prgh_terms(terms, device=False)
        upper, diag, src = np.empty(addr.nFaces), np.empty(addr.nCells), np.empty(addr.nCells)
        self._check(self.lib.b200_assemble_p_rgh(self.handle, C.addressof(t), upper.ctypes.data,
                                                 diag.ctypes.data, src.ctypes.data))
        return upper, diag, src

    def assemble_p_rgh_device(self, terms, d_upper_out, d_diag_out, d_source_out):
        """Device tensors in `terms` (bCells stays a host array, set once with set_boundary_faces)."""
        if terms.get("bCells") is not None:
            self.set_boundary_faces(terms["bCells"])
        t, keep = self._prgh_terms(terms, device=True)
        self._check(self.lib.b200_assemble_p_rgh_device(self.handle, C.addressof(t), _ptr(d_upper_out),
                                                        _ptr(d_diag_out), _ptr(d_source_out)))

    # ---- lduMatrix::Amul / fvMatrix::flux ----------------------------------------------------
    def amul(self, matrix, interfaceBouCoeffs, psi):
        psi = _f64(psi)
        out = np.empty_like(psi)
        bou, keep = _bou_array(interfaceBouCoeffs)
        self._check(self.lib.b200_amul(self.handle, matrix.diag.ctypes.data, matrix.upper.ctypes.data,
                                       bou, psi.ctypes.data, out.ctypes.data))
        return out

    def flux(self, matrix, psi):
        psi = _f64(psi)
        out = np.empty(matrix.upper.size, dtype=np.float64)
        self._check(self.lib.b200_flux(self.handle, matrix.upper.ctypes.data, psi.ctypes.data,
                                       out.ctypes.data))
        return out

    # ---- lduMatrix::solver::solve ------------------------------------------------------------
    def solve(self, diag, upper, interfaceBouCoeffs, source, psi, controls):
        perf = Perf()
        bou, keep = _bou_array(interfaceBouCoeffs)
        rc = self.lib.b200_solve(self.handle, diag.ctypes.data, upper.ctypes.data, bou,
                                 source.ctypes.data, psi.ctypes.data, C.byref(controls), C.byref(perf))
        self._check(rc)
        return perf

    def solve_device(self, d_diag, d_upper, d_bou_list, d_source, d_psi, controls):
        perf = Perf()
        n = len(d_bou_list) if d_bou_list else 0
        arr = (C.c_void_p * max(1, n))()
        for k in range(n):
            arr[k] = _ptr(d_bou_list[k])
        rc = self.lib.b200_solve_device(self.handle, _ptr(d_diag), _ptr(d_upper),
                                        C.cast(arr, C.c_void_p), _ptr(d_source), _ptr(d_psi),
                                        C.byref(controls), C.byref(perf))
        self._check(rc)
        return perf

    # ---- smoothSolver on (a)symmetric matrices -----------------------------------
    def amul_asym(self, matrix, interfaceBouCoeffs, psi):
        psi = _f64(psi)
        out = np.empty_like(psi)
        bou, keep = _bou_array(interfaceBouCoeffs)
        low = None if matrix.lower is None else matrix.lower.ctypes.data
        self._check(self.lib.b200_amul_asym(self.handle, matrix.diag.ctypes.data, matrix.upper.ctypes.data, low,
                                            bou, psi.ctypes.data, out.ctypes.data))
        return out

    def smooth_solve(self, diag, upper, lower, interfaceBouCoeffs, source, psi, controls):
        perf = Perf()
        bou, keep = _bou_array(interfaceBouCoeffs)
        rc = self.lib.b200_smooth_solve(self.handle, diag.ctypes.data, upper.ctypes.data,
                                        None if lower is None else lower.ctypes.data, bou,
                                        source.ctypes.data, psi.ctypes.data, C.byref(controls), C.byref(perf))
        self._check(rc)
        return perf

    def smooth_solve_device(self, d_diag, d_upper, d_lower, d_bou_list, d_source, d_psi, controls):
        perf = Perf()
        n = len(d_bou_list) if d_bou_list else 0
        arr = (C.c_void_p * max(1, n))()
        for k in range(n):
            arr[k] = _ptr(d_bou_list[k])
        rc = self.lib.b200_smooth_solve_device(self.handle, _ptr(d_diag), _ptr(d_upper), _ptr(d_lower),
                                               C.cast(arr, C.c_void_p), _ptr(d_source), _ptr(d_psi),
                                               C.byref(controls), C.byref(perf))
        self._check(rc)
        return perf

    # ---- PBiCG on asymmetric matrices ---------------------------------------------
    def bicg_solve(self, diag, upper, lower, interfaceBouCoeffs, interfaceIntCoeffs, source, psi, controls):
        perf = Perf()
        bou, keep = _bou_array(interfaceBouCoeffs)
        intc, keep2 = _bou_array(interfaceIntCoeffs)
        rc = self.lib.b200_bicg_solve(self.handle, diag.ctypes.data, upper.ctypes.data,
                                      None if lower is None else lower.ctypes.data, bou, intc,
                                      source.ctypes.data, psi.ctypes.data, C.byref(controls), C.byref(perf))
        self._check(rc)
        return perf

    def bicg_solve_device(self, d_diag, d_upper, d_lower, d_source, d_psi, controls):
        perf = Perf()
        rc = self.lib.b200_bicg_solve_device(self.handle, _ptr(d_diag), _ptr(d_upper), _ptr(d_lower), None, None,
                                             _ptr(d_source), _ptr(d_psi), C.byref(controls), C.byref(perf))
        self._check(rc)
        return perf

    # ---- harness helpers -----------------------------------------------------------------------
    def launch_count(self):
        return int(self.lib.b200_launch_count(self.handle))

    def force_iterations(self, n):
        self._check(self.lib.b200_debug_force_iterations(self.handle, int(n)))

    def profile(self, on=True):
        self._check(self.lib.b200_profile_enable(self.handle, 1 if on else 0))

    def profile_json(self):
        import json
        return json.loads(self.lib.b200_profile_json(self.handle).decode())

    def describe(self):
        import json
        return json.loads(self.lib.b200_describe(self.handle).decode())


def _ptr(x):
    if x is None:
        return None
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    if isinstance(x, np.ndarray):
        return x.ctypes.data
    return int(x)


def _bou_array(interfaceBouCoeffs):
    n = len(interfaceBouCoeffs) if interfaceBouCoeffs else 0
    keep = [_f64(b) for b in (interfaceBouCoeffs or [])]
    arr = (C.c_void_p * max(1, n))()
    for k, b in enumerate(keep):
        arr[k] = b.ctypes.data
    return C.cast(arr, C.c_void_p), (keep, arr)


def make_controls(solverControls):
    """lduMatrix::solver::readControls: maxIter 1000, minIter 0, tolerance 1e-6, relTol 0."""
    d = dict(solverControls or {})
    pre = d.get("preconditioner", "none")
    if isinstance(pre, dict):
        pre = pre.get("preconditioner", "none")
    # `preconditioner DIC` is the DIC-CLASS multicolour IC0 (the library picks its form); OpenFOAM's own DIC
    # (same elimination order, same iteration counts) is `B200 { dicMode exact; }`
    mode = d.get("B200", {}).get("dicMode", "auto")
    if pre == "DIC" and mode in ("exact", "eisenstat", "multicolour"):
        pre = "DIC-" + mode
    elif pre == "DIC" and mode != "auto":
        raise ValueError(f"Unknown dicMode {mode}; valid: auto exact eisenstat multicolour")
    if pre not in PRECOND:
        raise ValueError(f"Unknown symmetric matrix preconditioner {pre}; valid: {sorted(PRECOND)}")
    c = Controls()
    c.tolerance = float(d.get("tolerance", 1e-6))
    c.relTol = float(d.get("relTol", 0.0))
    c.maxIter = int(d.get("maxIter", 1000))
    c.minIter = int(d.get("minIter", 0))
    c.precond = PRECOND[pre]
    c.reserved = 0
    return c, pre


_default_ctx = None


def default_context():
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context()
    return _default_ctx


class B200PCG:
    """lduMatrix::solver selected by `solver B200PCG;` in system/fvSolution.

    Same constructor signature as every OpenFOAM lduMatrix::solver:
        B200PCG(fieldName, matrix, interfaceBouCoeffs, interfaceIntCoeffs, interfaces, solverControls)
    and `solve(psi, source, cmpt=0) -> SolverPerformance`, psi updated in place.
    `interfaces[k]` is None for non-coupled patches (UPtrList null slot) and a
    ProcessorLduInterface otherwise; anything else is rejected like an unsupported interface."""

    typeName = "B200PCG"

    def __init__(self, fieldName, matrix, interfaceBouCoeffs, interfaceIntCoeffs, interfaces,
                 solverControls, context=None):
        self.fieldName = fieldName
        self.matrix = matrix
        self.controls, self.preconditionerName = make_controls(solverControls)
        self.ctx = context or default_context()
        coupled = [k for k, itf in enumerate(interfaces or []) if itf is not None]
        for k in coupled:
            if not isinstance(interfaces[k], ProcessorLduInterface):
                raise B200Error(_lib.B200_EUNSUPPORTED, f"unsupported interface type on patch {k}")
        addr = matrix.lduAddr
        if [interfaces[k] for k in coupled] != list(addr.interfaces):
            raise ValueError("interfaces do not match lduAddr.interfaces")
        self.bou = [_f64(interfaceBouCoeffs[k]) for k in coupled]

    def solve(self, psi, source, cmpt=0):
        if not (isinstance(psi, np.ndarray) and psi.dtype == np.float64 and psi.flags.c_contiguous):
            raise TypeError("psi must be a contiguous float64 array (updated in place)")
        m = self.matrix
        self.ctx.set_addressing(m.lduAddr)
        perf = self.ctx.solve(m.diag, m.upper, self.bou, _f64(source), psi, self.controls)
        # the log line names what ran: only `dicMode exact` is OpenFOAM's DIC; the multicolour IC0 stand-in
        # (different iteration counts) prints as DIC(mc)B200PCG
        pre = {"none": "none", "diagonal": "diagonal", "DIC": "DIC(mc)", "DIC-exact": "DIC", "DIC-eisenstat": "DIC(mc)",
               "DIC-multicolour": "DIC(mc)"}[self.preconditionerName]
        return SolverPerformance(pre + self.typeName, self.fieldName, perf)


def make_smooth_controls(solverControls):
    """smoothSolver::readControls (OF-dev smoothSolver.C): lduMatrix::solver's maxIter 1000, minIter 0,
    tolerance 1e-6, relTol 0, plus nSweeps 1; `smoother` is mandatory as upstream."""
    d = dict(solverControls or {})
    sm = d.get("smoother")
    if isinstance(sm, dict):
        sm = sm.get("smoother")
    if sm not in SMOOTHER:
        raise ValueError(f"Unknown smoother {sm}; valid: {sorted(SMOOTHER)}")
    mode = d.get("B200", {}).get("sweepMode", "multicolour")
    if mode not in SWEEP_MODE:
        raise ValueError(f"Unknown sweepMode {mode}; valid: {sorted(SWEEP_MODE)}")
    c = SmoothControls()
    c.tolerance = float(d.get("tolerance", 1e-6))
    c.relTol = float(d.get("relTol", 0.0))
    c.maxIter = int(d.get("maxIter", 1000))
    c.minIter = int(d.get("minIter", 0))
    c.nSweeps = int(d.get("nSweeps", 1))
    c.smoother = SMOOTHER[sm]
    c.sweepMode = SWEEP_MODE[mode]
    c.reserved = 0
    return c, sm, mode


class B200smoothSolver:
    """lduMatrix::solver selected by `solver B200smoothSolver;` where the reference's fvSolution says
    `solver smoothSolver; smoother symGaussSeidel;` (cases/steckler/system/fvSolution:48-61: U, Yi, h, k).
    Registered upstream in BOTH run-time tables (symmetric and asymmetric matrices); same constructor signature and
    `solve(psi, source, cmpt=0) -> SolverPerformance` as B200PCG.  `B200 { sweepMode exact; }` visits the cells in
    OpenFOAM's own order (level-scheduled: identical sweep counts and psi); the default multicolour order is a
    GS-class stand-in and prints as `B200smoothSolver(mc)`."""

    typeName = "B200smoothSolver"

    def __init__(self, fieldName, matrix, interfaceBouCoeffs, interfaceIntCoeffs, interfaces,
                 solverControls, context=None):
        self.fieldName = fieldName
        self.matrix = matrix
        self.controls, self.smootherName, self.sweepMode = make_smooth_controls(solverControls)
        self.ctx = context or default_context()
        coupled = [k for k, itf in enumerate(interfaces or []) if itf is not None]
        for k in coupled:
            if not isinstance(interfaces[k], ProcessorLduInterface):
                raise B200Error(_lib.B200_EUNSUPPORTED, f"unsupported interface type on patch {k}")
        addr = matrix.lduAddr
        if [interfaces[k] for k in coupled] != list(addr.interfaces):
            raise ValueError("interfaces do not match lduAddr.interfaces")
        self.bou = [_f64(interfaceBouCoeffs[k]) for k in coupled]

    def solve(self, psi, source, cmpt=0):
        if not (isinstance(psi, np.ndarray) and psi.dtype == np.float64 and psi.flags.c_contiguous):
            raise TypeError("psi must be a contiguous float64 array (updated in place)")
        m = self.matrix
        self.ctx.set_addressing(m.lduAddr)
        perf = self.ctx.smooth_solve(m.diag, m.upper, m.lower, self.bou, _f64(source), psi, self.controls)
        name = self.typeName + ("" if self.sweepMode == "exact" else "(mc)")
        return SolverPerformance(name, self.fieldName, perf)


def make_bicg_controls(solverControls):
    """lduMatrix::solver::readControls + the asymmetric preconditioner keyword (`DILU`, `diagonal`, `none`);
    `B200 { diluMode exact; }` selects OpenFOAM's own DILU (level-scheduled), the default is the DILU-class stand-in."""
    d = dict(solverControls or {})
    pre = d.get("preconditioner", "none")
    if isinstance(pre, dict):
        pre = pre.get("preconditioner", "none")
    mode = d.get("B200", {}).get("diluMode", "multicolour")
    if mode not in ("multicolour", "exact"):
        raise ValueError(f"Unknown diluMode {mode}; valid: multicolour exact")
    if pre == "DILU" and mode == "exact":
        pre = "DILU-exact"
    if pre not in BICG_PRECOND:
        raise ValueError(f"Unknown asymmetric matrix preconditioner {pre}; valid: {sorted(BICG_PRECOND)}")
    c = Controls()
    c.tolerance = float(d.get("tolerance", 1e-6))
    c.relTol = float(d.get("relTol", 0.0))
    c.maxIter = int(d.get("maxIter", 1000))
    c.minIter = int(d.get("minIter", 0))
    c.precond = BICG_PRECOND[pre]
    c.reserved = 0
    return c, pre


class B200PBiCG:
    """lduMatrix::solver selected by `solver B200PBiCG;` where the reference's fvSolution says
    `solver PBiCG; preconditioner DILU;` (cases/wallFireSpread2D/system/fvSolution:66-73, cases/pyrolysis1D, the
    pyrolysis / panel regions; the 2.4.x golden logs of steckler).  Registered upstream in the asymmetric-matrix
    table; same constructor signature and `solve(psi, source, cmpt=0) -> SolverPerformance` as B200PCG.  One rank."""

    typeName = "B200PBiCG"

    def __init__(self, fieldName, matrix, interfaceBouCoeffs, interfaceIntCoeffs, interfaces,
                 solverControls, context=None):
        self.fieldName = fieldName
        self.matrix = matrix
        self.controls, self.preconditionerName = make_bicg_controls(solverControls)
        self.ctx = context or default_context()
        if any(itf is not None for itf in (interfaces or [])):
            raise B200Error(_lib.B200_EUNSUPPORTED, "B200PBiCG: coupled patches are not supported yet (one rank)")

    def solve(self, psi, source, cmpt=0):
        if not (isinstance(psi, np.ndarray) and psi.dtype == np.float64 and psi.flags.c_contiguous):
            raise TypeError("psi must be a contiguous float64 array (updated in place)")
        m = self.matrix
        self.ctx.set_addressing(m.lduAddr)
        perf = self.ctx.bicg_solve(m.diag, m.upper, m.lower, [], [], _f64(source), psi, self.controls)
        pre = {"none": "none", "diagonal": "diagonal", "DILU": "DILU(mc)", "DILU-exact": "DILU"}[self.preconditionerName]
        return SolverPerformance(pre + self.typeName, self.fieldName, perf)
