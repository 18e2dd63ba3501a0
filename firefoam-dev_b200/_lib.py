"""ctypes bindings of the C ABI (include/b200pcg.h) and of the harness library (libb200mesh.so).

The CUDA library is mandatory: there is no CPU fallback.  `load_pcg()` raises if
libb200pcg.so has not been built; every compute entry point raises `B200Error` when the
device is unusable (B200_ENODEVICE)."""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
PCG_SO = os.path.join(_HERE, "libb200pcg.so")
MESH_SO = os.path.join(_HERE, "libb200mesh.so")

B200_OK, B200_EINVAL, B200_ECUDA, B200_ENCCL, B200_ENODEVICE, B200_ESTATE, B200_EUNSUPPORTED, \
    B200_ENONFINITE, B200_ENOMEM = range(9)
ERRNAMES = {1: "EINVAL", 2: "ECUDA", 3: "ENCCL", 4: "ENODEVICE", 5: "ESTATE", 6: "EUNSUPPORTED",
            7: "ENONFINITE", 8: "ENOMEM"}

PRECOND = {"none": 0, "diagonal": 1, "DIC": 2, "DIC-exact": 3, "DIC-eisenstat": 4, "DIC-multicolour": 5}

# every symbol include/b200pcg.h declares (tests check that the .so exports them all)
ABI_SYMBOLS = [
    "b200_ctx_create", "b200_get_unique_id", "b200_ctx_destroy", "b200_last_error",
    "b200_abi_version", "b200_device_count", "b200_set_addressing", "b200_assemble_laplacian",
    "b200_assemble_laplacian_device", "b200_set_boundary_faces", "b200_assemble_p_rgh",
    "b200_assemble_p_rgh_device", "b200_solve", "b200_solve_device", "b200_amul", "b200_flux",
    "b200_smooth_solve", "b200_smooth_solve_device", "b200_amul_asym", "b200_bicg_solve", "b200_bicg_solve_device",
    "b200_host_alloc", "b200_host_free", "b200_launch_count", "b200_debug_force_iterations",
    "b200_profile_enable", "b200_profile_json", "b200_describe",
    "b200_dump_write", "b200_dump_read", "b200_dump_get", "b200_dump_header_json", "b200_dump_free",
    "b200_dump_last_error",
]


class B200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"b200pcg error {code} ({ERRNAMES.get(code, '?')}): {msg}")
        self.code = code


class Iface(C.Structure):
    _fields_ = [("nbrRank", C.c_int32), ("nFaces", C.c_int32),
                ("faceCells", C.POINTER(C.c_int32)), ("tag", C.c_int32)]


class Controls(C.Structure):
    _fields_ = [("tolerance", C.c_double), ("relTol", C.c_double), ("maxIter", C.c_int32),
                ("minIter", C.c_int32), ("precond", C.c_int32), ("reserved", C.c_int32)]


BICG_PRECOND = {"none": 0, "diagonal": 1, "DILU": 2, "DILU-exact": 3}
SMOOTHER = {"GaussSeidel": 0, "symGaussSeidel": 1}
SWEEP_MODE = {"multicolour": 0, "exact": 1}


class SmoothControls(C.Structure):
    """b200_smooth_controls (include/b200pcg.h)"""
    _fields_ = [("tolerance", C.c_double), ("relTol", C.c_double), ("maxIter", C.c_int32),
                ("minIter", C.c_int32), ("nSweeps", C.c_int32), ("smoother", C.c_int32),
                ("sweepMode", C.c_int32), ("reserved", C.c_int32)]


class Perf(C.Structure):
    _fields_ = [("initialResidual", C.c_double), ("finalResidual", C.c_double),
                ("normFactor", C.c_double), ("nIterations", C.c_int32), ("converged", C.c_int32),
                ("singular", C.c_int32), ("nColours", C.c_int32), ("solveMs", C.c_double),
                ("setupMs", C.c_double), ("h2dMs", C.c_double), ("d2hMs", C.c_double)]


class PrghTerms(C.Structure):
    """b200_prgh_terms (include/b200pcg.h)"""
    _fields_ = [("rDeltaT", C.c_double), ("V", C.c_void_p), ("psi", C.c_void_p), ("psi0", C.c_void_p),
                ("p0", C.c_void_p), ("nExplicit", C.c_int32), ("pad0", C.c_int32),
                ("explicitFields", C.c_void_p), ("phi", C.c_void_p), ("divSign", C.c_double),
                ("gamma_f", C.c_void_p), ("magSf", C.c_void_p), ("deltaCoeffs", C.c_void_p),
                ("lapSign", C.c_double), ("Su", C.c_void_p), ("nB", C.c_int32), ("pad1", C.c_int32),
                ("bCells", C.c_void_p), ("bPhi", C.c_void_p), ("bInternal", C.c_void_p),
                ("bBoundary", C.c_void_p)]


class Dump(C.Structure):
    """b200_dump (include/b200pcg.h)"""
    _fields_ = [("fieldName", C.c_char_p), ("rank", C.c_int32), ("nranks", C.c_int32),
                ("nCells", C.c_int32), ("nFaces", C.c_int32),
                ("lowerAddr", C.c_void_p), ("upperAddr", C.c_void_p), ("diag", C.c_void_p),
                ("upper", C.c_void_p), ("source", C.c_void_p), ("psi0", C.c_void_p),
                ("psiSolution", C.c_void_p), ("nIfaces", C.c_int32), ("ifaces", C.POINTER(Iface)),
                ("ifaceBouCoeffs", C.POINTER(C.c_void_p)), ("controls", Controls),
                ("havePerf", C.c_int32), ("perf", Perf), ("solverName", C.c_char_p),
                ("solveIndex", C.c_int32), ("time", C.c_double),
                ("lower", C.c_void_p), ("haveSmooth", C.c_int32), ("havePBiCG", C.c_int32), ("smooth", SmoothControls)]


def build(verbose=False):
    """Compile both shared libraries in-tree (nvcc -gencode arch=compute_100a,code=sm_100a)."""
    r = subprocess.run(["make", "-C", CSRC, "all"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("build failed:\n" + r.stdout[-4000:] + r.stderr[-4000:])
    if verbose:
        print(r.stdout[-2000:])


_pcg = None
_mesh = None


def load_pcg():
    global _pcg
    if _pcg is not None:
        return _pcg
    if not os.path.exists(PCG_SO):
        raise RuntimeError(f"{PCG_SO} is missing: build the CUDA extension first "
                           "(python -c 'import __graft_entry__ as g; g.build()'). "
                           "There is no CPU fallback.")
    L = C.CDLL(PCG_SO)
    vp, i32p, f64p = C.c_void_p, C.POINTER(C.c_int32), C.c_void_p  # doubles passed as raw addresses
    L.b200_ctx_create.argtypes = [C.c_int, C.c_int, C.c_int, vp, C.POINTER(vp)]
    L.b200_get_unique_id.argtypes = [vp]
    L.b200_ctx_destroy.argtypes = [vp]
    L.b200_ctx_destroy.restype = None
    L.b200_last_error.argtypes = [vp]
    L.b200_last_error.restype = C.c_char_p
    L.b200_abi_version.restype = C.c_int
    L.b200_device_count.restype = C.c_int
    L.b200_set_addressing.argtypes = [vp, C.c_uint64, C.c_int32, C.c_int32, vp, vp, C.c_int32, vp]
    L.b200_assemble_laplacian.argtypes = [vp, f64p, f64p, f64p, C.c_double, f64p, f64p]
    L.b200_assemble_laplacian_device.argtypes = [vp, f64p, f64p, f64p, C.c_double, f64p, f64p]
    L.b200_set_boundary_faces.argtypes = [vp, C.c_int32, vp]
    L.b200_assemble_p_rgh.argtypes = [vp, vp, f64p, f64p, f64p]
    L.b200_assemble_p_rgh_device.argtypes = [vp, vp, f64p, f64p, f64p]
    L.b200_solve.argtypes = [vp, f64p, f64p, vp, f64p, f64p, C.POINTER(Controls), C.POINTER(Perf)]
    L.b200_solve_device.argtypes = L.b200_solve.argtypes
    L.b200_amul.argtypes = [vp, f64p, f64p, vp, f64p, f64p]
    L.b200_flux.argtypes = [vp, f64p, f64p, f64p]
    L.b200_smooth_solve.argtypes = [vp, f64p, f64p, f64p, vp, f64p, f64p, C.POINTER(SmoothControls), C.POINTER(Perf)]
    L.b200_smooth_solve_device.argtypes = L.b200_smooth_solve.argtypes
    L.b200_amul_asym.argtypes = [vp, f64p, f64p, f64p, vp, f64p, f64p]
    L.b200_bicg_solve.argtypes = [vp, f64p, f64p, f64p, vp, vp, f64p, f64p, C.POINTER(Controls), C.POINTER(Perf)]
    L.b200_bicg_solve_device.argtypes = L.b200_bicg_solve.argtypes
    L.b200_host_alloc.argtypes = [C.POINTER(vp), C.c_size_t]
    L.b200_host_free.argtypes = [vp]
    L.b200_host_free.restype = None
    L.b200_launch_count.argtypes = [vp]
    L.b200_launch_count.restype = C.c_uint64
    L.b200_debug_force_iterations.argtypes = [vp, C.c_int32]
    L.b200_profile_enable.argtypes = [vp, C.c_int]
    L.b200_profile_json.argtypes = [vp]
    L.b200_profile_json.restype = C.c_char_p
    L.b200_describe.argtypes = [vp]
    L.b200_describe.restype = C.c_char_p
    L.b200_dump_write.argtypes = [C.c_char_p, C.POINTER(Dump)]
    L.b200_dump_read.argtypes = [C.c_char_p, C.POINTER(vp)]
    L.b200_dump_get.argtypes = [vp]
    L.b200_dump_get.restype = C.POINTER(Dump)
    L.b200_dump_header_json.argtypes = [vp]
    L.b200_dump_header_json.restype = C.c_char_p
    L.b200_dump_free.argtypes = [vp]
    L.b200_dump_free.restype = None
    L.b200_dump_last_error.restype = C.c_char_p
    # host-only plan inspection (plan_debug.cpp)
    L.b200_debug_plan_build.argtypes = [C.c_int, C.c_int32, C.c_int32, vp, vp, C.c_int32, vp]
    L.b200_debug_plan_build.restype = vp
    L.b200_debug_plan_error.restype = C.c_char_p
    L.b200_debug_plan_free.argtypes = [vp]
    L.b200_debug_plan_free.restype = None
    L.b200_debug_plan_get.argtypes = [vp, C.c_char_p, C.POINTER(vp), C.POINTER(C.c_int32)]
    L.b200_debug_plan_get.restype = C.c_int64
    L.b200_debug_plan_ncolours.argtypes = [vp]
    L.b200_debug_plan_ncolours.restype = C.c_int32
    L.b200_debug_plan_nentries.argtypes = [vp]
    L.b200_debug_plan_nentries.restype = C.c_int64
    _pcg = L
    return L


def load_mesh():
    global _mesh
    if _mesh is not None:
        return _mesh
    if not os.path.exists(MESH_SO):
        raise RuntimeError(f"{MESH_SO} is missing: run __graft_entry__.build()")
    L = C.CDLL(MESH_SO)
    vp = C.c_void_p
    L.b200mesh_hex_sizes.argtypes = [C.c_int] * 7 + [vp]
    L.b200mesh_hex_fill.argtypes = [C.c_int] * 7 + [C.c_uint64, C.c_double, C.c_double, C.c_double] + [vp] * 12
    L.b200mesh_bcc_sizes.argtypes = [C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_int, vp]
    L.b200mesh_bcc_fill.argtypes = [C.c_uint64, C.c_double, C.c_double, C.c_double] + [vp] * 11
    L.b200mesh_partition_simple.argtypes = [C.c_int32, vp, C.c_int, C.c_int, C.c_int, vp]
    L.b200mesh_partition_hierarchical.argtypes = [C.c_int32, vp, C.c_int, C.c_int, C.c_int, vp, vp]
    L.b200mesh_partition_rcb.argtypes = [C.c_int32, vp, C.c_int, vp]
    L.b200mesh_decompose.argtypes = [C.c_int32, C.c_int32, vp, vp, vp, C.c_int]
    L.b200mesh_decompose.restype = vp
    L.b200mesh_decompose_free.argtypes = [vp]
    L.b200mesh_decompose_free.restype = None
    L.b200mesh_decompose_get.argtypes = [vp, C.c_int, C.c_char_p, C.POINTER(vp)]
    L.b200mesh_decompose_get.restype = C.c_int64
    _mesh = L
    return L
