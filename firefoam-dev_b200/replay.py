"""Matrix dump / replay: the `.b200sys` files written by the OpenFOAM adapter under
B200PCG_DUMP=<dir> (adapter/B200PCG.C) -- what `lduMatrix::solver::solve` receives on one rank at
solver/pEqn.H:39 / solver/phrghEqn.H:48 plus the SolverPerformance it reported -- read into the
host-side mirror types, and written from them.  Layout: csrc/dump.cpp.

`read_dump` is a pure-numpy reader (no library needed, so a dump can be inspected anywhere);
`write_dump` goes through the C ABI (`b200_dump_write`), i.e. through the same writer the adapter
uses, so that the files the tests round-trip are the files OpenFOAM would produce."""
import ctypes as C
import json
import struct

import numpy as np

from . import _lib
from .ldu import LduAddressing, ProcessorLduInterface
from .meshgen import System

MAGIC = b"B200LDU\x01"
PRECOND_NAMES = {0: "none", 1: "diagonal", 2: "DIC", 3: "DIC", 4: "DIC", 5: "DIC"}


class DumpedSolve:
    """One dumped solve: `.system` (meshgen.System: addressing, diag, upper, source, interface
    coefficients), `.psi0`, `.psi` (solution or None), `.controls` (fvSolution-style dict),
    `.reference` (what the dumping solver reported, or None), `.header` (raw JSON)."""

    def __init__(self, header, arrays):
        self.header = header
        h = header
        ifs, bou = [], []
        for k, it in enumerate(h.get("interfaces", [])):
            ifs.append(ProcessorLduInterface(it["nbrRank"], arrays[f"iface{k}.faceCells"], myProcNo=h.get("rank", 0),
                                             tag=it.get("tag", 0)))
            bou.append(arrays[f"iface{k}.bouCoeffs"])
        addr = LduAddressing(h["nCells"], arrays["lowerAddr"], arrays["upperAddr"], ifs)
        self.system = System(addr, arrays["diag"], arrays["upper"], arrays["source"], bou, lower=arrays.get("lower"))
        self.psi0 = arrays["psi0"]
        self.psi = arrays.get("psi")
        c = h.get("controls", {})
        self.controls = {"preconditioner": c.get("preconditioner", "none"), "tolerance": c.get("tolerance", 1e-6),
                         "relTol": c.get("relTol", 0.0), "maxIter": c.get("maxIter", 1000),
                         "minIter": c.get("minIter", 0)}
        if c.get("precondCode") in (3, 4):
            self.controls["B200"] = {"dicMode": {3: "exact", 4: "eisenstat"}[c["precondCode"]]}
        # a smoothSolver solve: fvSolution-style dict for B200smoothSolver
        self.smooth = None
        if c.get("solver") == "smoothSolver":
            self.smooth = {"smoother": c.get("smoother", "symGaussSeidel"), "tolerance": c.get("tolerance", 1e-6),
                           "relTol": c.get("relTol", 0.0), "maxIter": c.get("maxIter", 1000), "minIter": c.get("minIter", 0),
                           "nSweeps": c.get("nSweeps", 1), "B200": {"sweepMode": c.get("sweepMode", "multicolour")}}
            self.controls = dict(self.smooth)
        # a PBiCG solve: fvSolution-style dict for B200PBiCG
        self.bicg = c.get("solver") == "PBiCG"
        if self.bicg:
            self.controls = {"preconditioner": c.get("preconditioner", "none"), "tolerance": c.get("tolerance", 1e-6),
                             "relTol": c.get("relTol", 0.0), "maxIter": c.get("maxIter", 1000), "minIter": c.get("minIter", 0)}
            if c.get("precondCode") == 3:
                self.controls["B200"] = {"diluMode": "exact"}
        self.reference = h.get("reference")
        self.fieldName = h.get("fieldName", "")
        self.rank, self.nranks = h.get("rank", 0), h.get("nranks", 1)


def read_dump(path):
    with open(path, "rb") as f:
        blob = f.read()
    if len(blob) < 16 or blob[:8] != MAGIC:
        raise ValueError(f"{path}: not a b200 system dump")
    (hlen,) = struct.unpack_from("<Q", blob, 8)
    header = json.loads(blob[16:16 + hlen].decode("utf-8"))
    if header.get("version") != 1:
        raise ValueError(f"{path}: unsupported dump version {header.get('version')}")
    arrays = {}
    for a in header["arrays"]:
        dt = {"i4": "<i4", "f8": "<f8"}[a["dtype"]]
        off, cnt = int(a["offset"]), int(a["count"])
        if off + cnt * int(dt[2]) > len(blob):
            raise ValueError(f"{path}: array {a['name']} out of bounds")
        arrays[a["name"]] = np.frombuffer(blob, dtype=dt, count=cnt, offset=off).copy()
    for name, n in (("lowerAddr", "nFaces"), ("upperAddr", "nFaces"), ("upper", "nFaces"), ("diag", "nCells"),
                    ("source", "nCells"), ("psi0", "nCells")):
        if name not in arrays or arrays[name].size != header[n]:
            raise ValueError(f"{path}: array {name} missing or wrong size")
    return DumpedSolve(header, arrays)


def write_dump(path, system, psi0, controls, fieldName="p_rgh", psi=None, reference=None, solverName=None,
               rank=0, nranks=1, solveIndex=0, time=0.0):
    """Write one solve through the C-ABI writer (the adapter's code path).  `controls`: fvSolution-style
    dict; `reference`: object/dict with initialResidual, finalResidual, nIterations[, converged, singular]."""
    from .ldu import make_controls
    L = _lib.load_pcg()
    a = system.addr
    f64 = lambda x: np.ascontiguousarray(x, dtype=np.float64)
    keep = [a.lowerAddr, a.upperAddr, f64(system.diag), f64(system.upper), f64(system.source), f64(psi0)]
    d = _lib.Dump()
    d.fieldName = fieldName.encode()
    d.rank, d.nranks, d.nCells, d.nFaces = rank, nranks, a.nCells, a.nFaces
    d.lowerAddr, d.upperAddr = keep[0].ctypes.data, keep[1].ctypes.data
    d.diag, d.upper, d.source, d.psi0 = (k.ctypes.data for k in keep[2:6])
    if psi is not None:
        keep.append(f64(psi))
        d.psiSolution = keep[-1].ctypes.data
    n = len(a.interfaces)
    ifs = (_lib.Iface * max(1, n))()
    bous = (C.c_void_p * max(1, n))()
    for k, itf in enumerate(a.interfaces):
        ifs[k].nbrRank, ifs[k].nFaces, ifs[k].tag = itf.neighbProcNo, itf.faceCells.size, itf.tag
        ifs[k].faceCells = itf.faceCells.ctypes.data_as(C.POINTER(C.c_int32))
        keep.append(f64(system.bou[k]))
        bous[k] = keep[-1].ctypes.data
    d.nIfaces, d.ifaces, d.ifaceBouCoeffs = n, ifs, bous
    if system.lower is not None:
        keep.append(f64(system.lower))
        d.lower = keep[-1].ctypes.data
    if controls.get("smoother") is not None:
        from .ldu import make_smooth_controls
        d.smooth, _, _ = make_smooth_controls(controls)
        d.haveSmooth = 1
    elif controls.get("solver") == "PBiCG":
        from .ldu import make_bicg_controls
        ctl, _ = make_bicg_controls({k: v for k, v in controls.items() if k != "solver"})
        d.controls = ctl
        d.havePBiCG = 1
    else:
        ctl, _ = make_controls(controls)
        d.controls = ctl
    if reference is not None:
        get = (lambda k, dflt=0: reference.get(k, dflt)) if isinstance(reference, dict) else \
              (lambda k, dflt=0: getattr(reference, k, dflt))
        d.havePerf = 1
        d.perf.initialResidual, d.perf.finalResidual = float(get("initialResidual")), float(get("finalResidual"))
        d.perf.nIterations = int(get("nIterations"))
        d.perf.converged, d.perf.singular = int(bool(get("converged", 1))), int(bool(get("singular", 0)))
        d.solverName = (solverName or (get("solverName", "") or "")).encode()
    d.solveIndex, d.time = int(solveIndex), float(time)
    rc = L.b200_dump_write(str(path).encode(), C.byref(d))
    if rc != 0:
        raise _lib.B200Error(rc, L.b200_dump_last_error().decode())


def replay(path_or_dump, context=None, preconditioner=None):
    """Re-solve a dumped system through the CUDA path; returns (psi, SolverPerformance, DumpedSolve)."""
    from .ldu import B200PCG
    d = path_or_dump if isinstance(path_or_dump, DumpedSolve) else read_dump(path_or_dump)
    if d.nranks != 1:
        raise ValueError("replay() handles single-rank dumps; multi-rank dumps are replayed one rank per GPU "
                         "(tests/mgpu_worker.py shows the pattern)")
    ctl = dict(d.controls)
    s = d.system
    psi = d.psi0.copy()
    if d.smooth is not None:
        from .ldu import B200smoothSolver
        perf = B200smoothSolver(d.fieldName, s.matrix, s.bou, None, s.interfaces, ctl, context=context).solve(psi, s.source)
        return psi, perf, d
    if d.bicg:
        from .ldu import B200PBiCG
        perf = B200PBiCG(d.fieldName, s.matrix, s.bou, None, s.interfaces, ctl, context=context).solve(psi, s.source)
        return psi, perf, d
    if preconditioner:
        ctl["preconditioner"] = preconditioner
    perf = B200PCG(d.fieldName, s.matrix, s.bou, None, s.interfaces, ctl, context=context).solve(psi, s.source)
    return psi, perf, d
