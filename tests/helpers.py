"""Shared test helpers: numpy emulation of the device row structure (so the data structure the
kernels consume can be validated on a CPU-only box), random LDU meshes, hydrostatic KAT loop."""
import ctypes as C

import numpy as np

from firefoam_dev_b200 import _lib
from firefoam_dev_b200.ldu import LduAddressing, ProcessorLduInterface
from firefoam_dev_b200.meshgen import System


class PlanView:
    """HostPlan (csrc/plan.cpp) pulled out through the host-only debug ABI."""

    NAMES = ["perm", "iperm", "colourStart", "sliceBase", "rowLen", "col", "faceOf", "nbrRank",
             "patchStart", "slotRow", "bRow", "bStart", "bSlot", "segStart", "rowColour", "colBase", "col16"]

    def __init__(self, ordering, addr, renumber=0, tileRows=0):
        """renumber: 0 off (the default of the host-only debug ABI), -1 auto, 1 force RCM;
        tileRows > 0: tiled multicolour order (rows ordered by tile, colour, base position)."""
        L = _lib.load_pcg()
        ifs = (_lib.Iface * max(1, len(addr.interfaces)))()
        # b200_dbg_iface = {nbrRank, nFaces, faceCells} -- same leading layout, pack explicitly
        class DbgIface(C.Structure):
            _fields_ = [("nbrRank", C.c_int32), ("nFaces", C.c_int32), ("faceCells", C.c_void_p)]
        dif = (DbgIface * max(1, len(addr.interfaces)))()
        for k, itf in enumerate(addr.interfaces):
            dif[k].nbrRank, dif[k].nFaces = itf.neighbProcNo, itf.faceCells.size
            dif[k].faceCells = itf.faceCells.ctypes.data
        L.b200_debug_plan_build2.argtypes = [C.c_int, C.c_int, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                             C.c_int32, C.c_void_p, C.c_int32]
        L.b200_debug_plan_build2.restype = C.c_void_p
        h = L.b200_debug_plan_build2(ordering, renumber, addr.nCells, addr.nFaces,
                                     addr.lowerAddr.ctypes.data, addr.upperAddr.ctypes.data,
                                     len(addr.interfaces), C.cast(dif, C.c_void_p), tileRows)
        if not h:
            raise ValueError(L.b200_debug_plan_error().decode())
        try:
            L.b200_debug_plan_sym_valid.argtypes = [C.c_void_p]
            self.symValid = bool(L.b200_debug_plan_sym_valid(h))
            self.sym = {}
            L.b200_debug_plan_sym_wu.argtypes = [C.c_void_p]
            L.b200_debug_plan_sym_wl.argtypes = [C.c_void_p]
            self.symWU, self.symWL = L.b200_debug_plan_sym_wu(h), L.b200_debug_plan_sym_wl(h)
            L.b200_debug_plan_ntiles.argtypes = [C.c_void_p]
            self.nTiles = L.b200_debug_plan_ntiles(h)
            for nm, dt in (("uCol", np.int32), ("uFace", np.int32), ("lRef", np.uint32), ("rowLen", np.uint32)):
                ptr, eb = C.c_void_p(), C.c_int32()
                n = L.b200_debug_plan_get(h, ("sym." + nm).encode(), C.byref(ptr), C.byref(eb))
                self.sym[nm] = (np.empty(0, dtype=dt) if n <= 0 else
                                np.frombuffer((C.c_char * (n * eb.value)).from_address(ptr.value), dtype=dt).copy())
            L.b200_debug_plan_sr_valid.argtypes = [C.c_void_p]
            self.srValid = bool(L.b200_debug_plan_sr_valid(h))
            self.sr = {}
            for nm, dt in (("meta", np.uint32), ("ownBase", np.int64), ("ownFace", np.int32)):
                ptr, eb = C.c_void_p(), C.c_int32()
                n = L.b200_debug_plan_get(h, ("sr." + nm).encode(), C.byref(ptr), C.byref(eb))
                self.sr[nm] = (np.empty(0, dtype=dt) if n <= 0 else
                               np.frombuffer((C.c_char * (n * eb.value)).from_address(ptr.value), dtype=dt).copy())
            L.b200_debug_plan_renumbered.argtypes = [C.c_void_p]
            L.b200_debug_plan_span.argtypes = [C.c_void_p, C.c_int]
            L.b200_debug_plan_span.restype = C.c_double
            self.renumbered = bool(L.b200_debug_plan_renumbered(h))
            self.spanNatural, self.spanUsed = L.b200_debug_plan_span(h, 0), L.b200_debug_plan_span(h, 1)
            self.nColours = L.b200_debug_plan_ncolours(h)
            self.nEntries = L.b200_debug_plan_nentries(h)
            for nm in self.NAMES:
                ptr, eb = C.c_void_p(), C.c_int32()
                n = L.b200_debug_plan_get(h, nm.encode(), C.byref(ptr), C.byref(eb))
                assert n >= 0, nm
                dt = {2: np.uint16, 4: np.int32, 8: np.int64}[eb.value] if nm != "rowLen" else np.uint32
                if n == 0:
                    arr = np.empty(0, dtype=dt)
                else:
                    arr = np.frombuffer((C.c_char * (n * eb.value)).from_address(ptr.value), dtype=dt).copy()
                setattr(self, nm, arr)
        finally:
            L.b200_debug_plan_free(h)
        self.N, self.F = addr.nCells, addr.nFaces
        self.nLower = (self.rowLen & 0xFFFF).astype(np.int64)
        self.nTotal = (self.rowLen >> 16).astype(np.int64)
        self.symNLower = (self.sym["rowLen"] & 0xFFFF).astype(np.int64)
        self.symNTotal = (self.sym["rowLen"] >> 16).astype(np.int64)

    def rows_of_colour(self, k):
        """internal rows of colour k, ascending (all tiles)"""
        C = self.nColours
        out = []
        for t in range(self.nTiles):
            out.extend(range(int(self.segStart[t * C + k]), int(self.segStart[t * C + k + 1])))
        return out

    def entry(self, r, j):
        return int(self.sliceBase[r // 32]) + 32 * j + (r % 32)

    def to_internal(self, v):
        return v[self.perm] if self.perm.size else v.copy()

    def to_natural(self, v):
        if not self.perm.size:
            return v.copy()
        out = np.empty_like(v)
        out[self.perm] = v
        return out

    def values(self, upper):
        val = np.zeros(self.nEntries)
        m = self.faceOf >= 0
        val[m] = upper[self.faceOf[m]]
        return val

    def values_asym(self, upper, lower, lowerAddr):
        """k_fill_values_asym: the entry of row r on face f carries upper[f] when r's cell OWNS the face
        (A[l][u] = upper), lower[f] when it is the face's neighbour (A[u][l] = lower)."""
        val = np.zeros(self.nEntries)
        for r in range(self.N):
            c = int(self.perm[r]) if self.perm.size else r
            for j in range(self.nTotal[r]):
                e = self.entry(r, j)
                f = int(self.faceOf[e])
                val[e] = upper[f] if lowerAddr[f] == c else lower[f]
        return val

    def gs_rows(self, k, diag_i, val, b_i, x_i, res=False):
        """k_gs_rows over colour / level k, in place: x[r] = (b[r] - sum_j val*x[col]) / diag[r], entries in the
        plan's order [earlier | later], each ascending natural face order.  res: returns the rows' share of
        sum |residual| as the kernel forms it, |w - diag * x_new|."""
        tot = 0.0
        for r in self.rows_of_colour(k):
            w = b_i[r]
            for j in range(self.nTotal[r]):
                e = self.entry(r, j)
                w = w - val[e] * x_i[self.col[e]]
            x_i[r] = w / diag_i[r]
            if res:
                tot += abs(w - diag_i[r] * x_i[r])
        return tot

    def gs_residual(self, diag_i, val, b_i, x_i, r0=0, r1=None):
        """k_gs_resid: sum |b - A x| over the rows [r0, r1) with the row sum in lduMatrix::residual's order."""
        tot = 0.0
        for r in range(r0, self.N if r1 is None else r1):
            w = b_i[r] - diag_i[r] * x_i[r]
            for j in range(self.nTotal[r]):
                e = self.entry(r, j)
                w = w - val[e] * x_i[self.col[e]]
            tot += abs(w)
        return tot

    # --- numpy emulation of the kernels' row loops (same operation order) ------------------
    def spmv(self, diag_i, val, x_i):
        y = np.empty(self.N)
        for r in range(self.N):
            acc = diag_i[r] * x_i[r]
            for j in range(self.nTotal[r]):
                e = self.entry(r, j)
                acc = acc + val[e] * x_i[self.col[e]]
            y[r] = acc
        return y

    def spmv_sym(self, diag_i, upper, x_i):
        """numpy emulation of k_spmv_sym: upper entries streamed, lower entries by reference."""
        y = self.sym
        uv = np.zeros(y["uCol"].size)
        m = y["uFace"] >= 0
        uv[m] = upper[y["uFace"][m]]
        out = np.empty(self.N)
        for r in range(self.N):
            nL, nU = int(self.symNLower[r]), int(self.symNTotal[r] - self.symNLower[r])
            acc = diag_i[r] * x_i[r]
            lb = (r // 32) * 32 * self.symWL + (r % 32)
            ub = (r // 32) * 32 * self.symWU + (r % 32)
            for j in range(nL):
                pk = int(y["lRef"][lb + 32 * j])
                a, q = pk >> 5, pk & 31
                assert a < r
                acc = acc + uv[(a // 32) * 32 * self.symWU + 32 * q + (a % 32)] * x_i[a]
            for j in range(nU):
                c = int(y["uCol"][ub + 32 * j])
                assert c > r
                acc = acc + uv[ub + 32 * j] * x_i[c]
            out[r] = acc
        return out

    def spmv_sr(self, diag_i, upper, x_i):
        """numpy emulation of k_spmv_sr (renumbered natural plans): one meta word per entry, in ascending natural
        face order; q == 31: the row's next own value, else the q-th own value of row `column`."""
        meta, ownBase, ownFace = self.sr["meta"], self.sr["ownBase"], self.sr["ownFace"]
        ov = np.zeros(ownFace.size)
        m = ownFace >= 0
        ov[m] = upper[ownFace[m]]
        out = np.empty(self.N)
        for r in range(self.N):
            acc = diag_i[r] * x_i[r]
            jown = 0
            for j in range(self.nTotal[r]):
                w = int(meta[self.entry(r, j)])
                a, q = w >> 5, w & 31
                assert a == self.col[self.entry(r, j)]
                if q == 31:
                    assert a > r
                    pos = int(ownBase[r // 32]) + 32 * jown + (r % 32)
                    jown += 1
                else:
                    assert a < r
                    pos = int(ownBase[a // 32]) + 32 * q + (a % 32)
                assert ownFace[pos] == self.faceOf[self.entry(r, j)]     # the SAME coefficient, stored once
                acc = acc + ov[pos] * x_i[a]
            out[r] = acc
        return out

    def dic_calc_rd(self, diag_i, val):
        rD = np.empty(self.N)
        for k in range(self.nColours):
            for r in self.rows_of_colour(k):
                d = diag_i[r]
                for j in range(self.nLower[r]):
                    e = self.entry(r, j)
                    d = d - (val[e] * val[e]) / rD[self.col[e]]
                rD[r] = d
        return 1.0 / rD

    def dic_precondition(self, rD, val, r_i):
        w = np.empty(self.N)
        for k in range(self.nColours):
            for r in self.rows_of_colour(k):
                acc = rD[r] * r_i[r]
                for j in range(self.nLower[r]):
                    e = self.entry(r, j)
                    acc = acc - (rD[r] * val[e]) * w[self.col[e]]
                w[r] = acc
        for k in range(self.nColours - 2, -1, -1):
            for r in self.rows_of_colour(k):
                acc = w[r]
                for j in range(self.nTotal[r] - 1, self.nLower[r] - 1, -1):
                    e = self.entry(r, j)
                    acc = acc - (rD[r] * val[e]) * w[self.col[e]]
                w[r] = acc
        return w


def random_ldu(N, avg_deg, seed, spd=True):
    """Random symmetric LDU system (irregular row lengths, including empty rows)."""
    rng = np.random.default_rng(seed)
    nF = int(N * avg_deg / 2)
    a = rng.integers(0, N, size=nF)
    b = rng.integers(0, N, size=nF)
    keep = a != b
    lo, hi = np.minimum(a, b)[keep], np.maximum(a, b)[keep]
    pairs = np.unique(np.stack([lo, hi], 1), axis=0)   # sorted lexicographically = upper-triangular
    l, u = pairs[:, 0].astype(np.int32), pairs[:, 1].astype(np.int32)
    upper = -rng.uniform(0.1, 1.0, size=l.size)
    diag = np.zeros(N)
    np.add.at(diag, l, -upper)
    np.add.at(diag, u, -upper)
    diag += rng.uniform(0.01, 0.1, size=N)
    addr = LduAddressing(N, l, u)
    xstar = rng.standard_normal(N)
    s = System(addr, diag, upper, np.zeros(N), [], xstar)
    from oracle import oracle as orc
    s.source = orc.amul(s, xstar)[0]
    return s


# PCG + diagonal on the (digit-pinned) steckler hydrostatic loop: self-derived by the oracle, no shipped
# case or log runs this mode
STECKLER_DIAG_COUNTS = [87, 86, 21, 0, 0]
STECKLER_DIC_COUNTS = [29, 32, 7, 0, 0]      # cases/steckler/original/linux64/log.fireFoam:92-100


def hydrostatic_loop(case, laplacian, solve):
    """solver/phrghEqn.H:30-56: returns [(initial, final, iters, variation)] per corrector.
    solve(matrix, source, psi) -> object with initialResidual/finalResidual/nIterations."""
    out = []
    psi = case.ph_rgh.copy()
    for _ in range(case.N_CORR):
        m, src = case.assemble(laplacian)
        perf = solve(m, src, psi)
        var = case.update(psi)
        out.append((perf.initialResidual, perf.finalResidual, perf.nIterations, var))
    return out


# ---- numpy transliteration of the Eisenstat-form DIC-class loop (kernels.cuh k_eis_*) --------------
def eis_check_interval(q):          # kernels.cuh eis_check_interval
    return 1 if q < 1.5 else (2 if q < 4.0 else (8 if q < 32.0 else 32))


class SingleRank:
    """Communicator of the emulations below on one rank; tests/gloo_eis_worker.py supplies the gloo twin
    (allsum = all-reduce of one double, exchange = grouped send/recv of the patch-face values in slot order)."""
    nGlobal = None

    def allsum(self, v):
        return float(v)

    def exchange(self, send):
        return np.empty(0)


def _amul(pv, comm, diag_i, val, bou, x):
    """lduMatrix::Amul on the plan's row structure: k_pack -> exchange -> k_spmv -> k_iface_fix"""
    recv = comm.exchange(x[pv.slotRow]) if pv.slotRow.size else np.empty(0)
    y = pv.spmv(diag_i, val, x)
    for b in range(pv.bRow.size):
        acc = y[pv.bRow[b]]
        for e in range(pv.bStart[b], pv.bStart[b + 1]):
            acc = acc - bou[pv.bSlot[e]] * recv[pv.bSlot[e]]
        y[pv.bRow[b]] = acc
    return y


def _norm_factor(pv, comm, diag_i, val, bou, psi_i, src_i):
    """lduMatrix::solver::normFactor on the plan's row structure."""
    wA = _amul(pv, comm, diag_i, val, bou, psi_i)
    sumA = pv.spmv(diag_i, val, np.ones(pv.N))
    if pv.slotRow.size:
        sumA = sumA - np.bincount(pv.slotRow, weights=bou, minlength=pv.N)
    nGlobal = comm.allsum(float(pv.N))
    xRef = comm.allsum(psi_i.sum()) / nGlobal
    nf = comm.allsum((np.abs(wA - sumA * xRef) + np.abs(src_i - sumA * xRef)).sum()) + 1e-20
    return wA, nf


def pcg_multicolour_reference(pv, diag, upper, source, psi0, tol=1e-6, relTol=0.0, maxIter=1000, minIter=0,
                              comm=None, bou=None):
    """The three-kernel DIC-class loop (PCG.C control flow, rank-local multicolour IC0 preconditioner) on
    the plan's row structure: what B200_PRECOND_DIC_MC computes.  Returns (psi natural, nIter, finalRes)."""
    comm = comm or SingleRank()
    bou = np.zeros(0) if bou is None else bou
    val = pv.values(upper)
    d, b, x = pv.to_internal(diag), pv.to_internal(source), pv.to_internal(psi0)
    wA, nf = _norm_factor(pv, comm, d, val, bou, x, b)
    r = b - wA
    init = final = comm.allsum(np.abs(r).sum()) / nf
    conv = lambda: final < tol or (relTol > 1e-20 and final < relTol * init)
    n = 0
    if minIter > 0 or not conv():
        rD = pv.dic_calc_rd(d, val)
        rho = 1e20
        p = None
        while True:
            rho_old = rho
            w = pv.dic_precondition(rD, val, r)
            rho = comm.allsum(w @ r)
            p = w.copy() if n == 0 else w + (rho / rho_old) * p
            w = _amul(pv, comm, d, val, bou, p)
            alpha = rho / comm.allsum(w @ p)
            x = x + alpha * p
            r = r - alpha * w
            final = comm.allsum(np.abs(r).sum()) / nf
            n += 1
            if not ((n - 1 < maxIter and not conv()) or n < minIter):
                break
    return pv.to_natural(x), n, final


def pcg_eisenstat_emulated(pv, diag, upper, source, psi0, tol=1e-6, relTol=0.0, maxIter=1000, minIter=0,
                           halo=False, comm=None, bou=None, overlap=False):
    """Kernel-by-kernel transliteration of B200_PRECOND_DIC_MC_EIS (solver.cu eis_setup /
    enqueue_eis_iteration / launch_eis_sweeps, kernels.cuh k_eis_* and the STEP_EIS_* scalar steps).
    One rank by default; `comm` + `bou` (interface coefficients in slot order) run it across ranks.
    halo=True takes the multi-rank kernel selection even on one rank (empty halo term);
    overlap=True the overlapped form of it (B200PCG_EIS_OVERLAP=1: first colour's interface rows first,
    the first colour's forward sweep fused into its backward sweep on the rows without a processor face).
    Returns (psi natural, nIter, finalRes, number of true-residual evaluations)."""
    comm = comm or SingleRank()
    bou = np.zeros(0) if bou is None else bou
    halo = halo or pv.slotRow.size > 0
    N, C = pv.N, pv.nColours
    val = pv.values(upper)
    diag_i, src, psi = pv.to_internal(diag), pv.to_internal(source), pv.to_internal(psi0)
    lastStart = int(pv.colourStart[C - 1])
    S = dict(done=0, nIter=0, converged=0, singular=0, pendingPsi=0, wArA=1e20, wArAold=1e20, beta=0.0,
             alpha=0.0, cRatio=0.0, sinceCheck=0, needCheck=0, sigma=1.0)
    conv = lambda: S["finalRes"] < tol or (relTol > 1e-20 and S["finalRes"] < relTol * S["initRes"])
    # spmv_full<INIT> + k_sum + k_norm_resid (STEP_NORM)
    wA, nf = _norm_factor(pv, comm, diag_i, val, bou, psi, src)
    nGlobal = comm.allsum(float(N))
    rh = src - wA
    S["normFactor"] = nf
    S["initRes"] = S["finalRes"] = comm.allsum(np.abs(rh).sum()) / nf
    S["converged"] = int(conv())
    S["done"] = 0 if (minIter > 0 or not S["converged"]) else 1
    # eis_setup: k_dic_calc_rd per colour -> dT
    dT = np.empty(N)
    for k in range(C):
        for r in pv.rows_of_colour(k):
            d = diag_i[r]
            for j in range(pv.nLower[r]):
                e = pv.entry(r, j)
                d = d - (val[e] * val[e]) / dT[pv.col[e]]
            dT[r] = d
    # k_eis_sign + STEP_EIS_SIGN (global counts)
    neg = comm.allsum(float((dT < 0).sum()))
    bad = comm.allsum(float((~((np.abs(dT) > 0) & (np.abs(dT) < 1.7e308))).sum()))
    if not S["done"]:
        if bad > 0 or (0 < neg < nGlobal):
            raise ValueError("DIC pivots are zero or of mixed sign")
        S["sigma"] = -1.0 if neg > 0 else 1.0
    sv = eb = xa = None
    rowB = np.full(N, -1, dtype=np.int64)
    rowB[pv.bRow] = np.arange(pv.bRow.size)
    nB0 = int((pv.bRow < pv.colourStart[1]).sum()) if C >= 2 else 0
    assert np.all(np.diff(pv.bRow) > 0)                       # ascending: the first colour's rows are a prefix
    # k_eis_setup (gated by done on the device; the exchange of s is issued by every rank regardless)
    sigma = S["sigma"]
    with np.errstate(all="ignore"):
        sv_all = 1.0 / np.sqrt(np.abs(dT))
    recv_s = comm.exchange(sv_all[pv.slotRow]) if pv.slotRow.size else np.empty(0)
    if not S["done"]:
        sv = sv_all
        eb = diag_i / dT - 2.0
        rh = (sigma * sv) * rh
        xa = np.zeros(N)
        # k_eis_scale_bou, k_eis_scale_vals (both triangles of the plan's coefficient copy)
        bou = bou * (sigma * (sv[pv.slotRow] * recv_s)) if pv.slotRow.size else bou
        val = val.copy()
        for r in range(N):
            for j in range(pv.nTotal[r]):
                e = pv.entry(r, j)
                val[e] = val[e] * (sigma * (sv[r] * sv[pv.col[e]]))
        # k_eis_init_fwd per colour, in place
        for k in range(C):
            for r in pv.rows_of_colour(k):
                w = rh[r]
                for j in range(pv.nLower[r]):
                    e = pv.entry(r, j)
                    w = w - val[e] * rh[pv.col[e]]
                rh[r] = w
    g = comm.allsum((rh * rh).sum())                        # k_eis_rho0, STEP_EIS_RHO0
    if not S["done"]:
        S["wArA"], S["beta"] = g, 0.0
        a = np.sqrt(abs(g))
        S["cRatio"] = S["finalRes"] / a if a > 0 else 0.0
    ph, t, y = np.zeros(N), np.zeros(N), np.zeros(N)
    hb = np.zeros(pv.bRow.size)
    checks = 0
    fuse0 = 0 if C < 2 else (1 if not halo else (2 if overlap else 0))

    def halo_term(tvec):                                     # k_pack -> exchange -> k_eis_halo
        recv = comm.exchange(tvec[pv.slotRow]) if pv.slotRow.size else np.empty(0)
        for b in range(pv.bRow.size):
            acc = 0.0
            for e in range(pv.bStart[b], pv.bStart[b + 1]):
                acc = acc - bou[pv.bSlot[e]] * recv[pv.bSlot[e]]
            hb[b] = acc

    def bwd_row(r):
        w = ph[r]
        for j in range(pv.nTotal[r] - 1, pv.nLower[r] - 1, -1):
            e = pv.entry(r, j)
            w = w - val[e] * t[pv.col[e]]
        return w

    enq, cap = 0, max(maxIter + 1, minIter)
    while not S["done"] and enq < cap:
        enq += 1
        # k_eis_p: rows >= lastStart keep p^ in t
        first = S["nIter"] == 0
        p = rh.copy()
        if not first:
            po = np.where(np.arange(N) >= lastStart, t, ph)
            xa = xa + S["alpha"] * t
            p = p + S["beta"] * po
        ph = ph.copy()
        t = t.copy()
        ph[:lastStart] = p[:lastStart]
        t[lastStart:] = p[lastStart:]
        dot = 0.0
        # k_eis_bwd, colours C-2 .. 0
        for k in range(C - 2, -1, -1):
            if k == 0 and fuse0 == 2:
                for b in range(nB0):                         # k_eis_bwd_rows, then the exchange starts
                    t[pv.bRow[b]] = bwd_row(int(pv.bRow[b]))
                t_sent = t.copy()
            for r in pv.rows_of_colour(k):
                assert r < lastStart
                if k == 0 and fuse0 == 2 and rowB[r] >= 0:
                    continue                                 # interface row: left alone by k_eis_bwd<2>
                w = bwd_row(r)
                t[r] = w
                if k == 0 and fuse0:
                    assert pv.nLower[r] == 0 and abs(eb[r] + 1.0) < 1e-12    # D- == 1 on first-colour rows
                    yv = ph[r] - w
                    y[r] = yv
                    dot += ph[r] * (w + yv)
        if halo and fuse0 == 2:
            assert np.array_equal(t_sent[pv.slotRow], t[pv.slotRow])          # nothing the pack reads was touched
            halo_term(t_sent)
            for b in range(nB0):                             # k_eis_fwd_rows
                r = int(pv.bRow[b])
                yv = (ph[r] - t[r]) + hb[b]
                y[r] = yv
                dot += ph[r] * (t[r] + yv)
        elif halo:
            halo_term(t)
        # k_eis_fwd
        for k in range(1 if fuse0 else 0, C):
            last = k == C - 1
            for r in pv.rows_of_colour(k):
                assert (r >= lastStart) == last
                tv = t[r]
                pp = tv if last else ph[r]
                w = pp + eb[r] * tv
                if halo and rowB[r] >= 0:
                    w = w + hb[rowB[r]]
                for j in range(pv.nLower[r]):
                    e = pv.entry(r, j)
                    assert pv.col[e] < lastStart     # a gathered y is never a stored w^
                    w = w - val[e] * y[pv.col[e]]
                wh = tv + w
                y[r] = wh if last else w
                dot += pp * wh
        # STEP_WAPA
        dot = comm.allsum(dot)
        S["wApA"] = dot
        if not (abs(dot) / nf > 1e-300):
            S["singular"], S["done"], S["pendingPsi"] = 1, 1, 0
            break
        S["alpha"] = S["wArA"] / dot
        S["pendingPsi"] = 1
        # k_eis_r + STEP_EIS_RHO
        w = y.copy()
        w[:lastStart] = t[:lastStart] + y[:lastStart]
        rh = rh - S["alpha"] * w
        g = comm.allsum((rh * rh).sum())
        S["wArAold"], S["wArA"] = S["wArA"], g
        S["beta"] = S["wArA"] / S["wArAold"]
        old = S["nIter"]
        S["nIter"] = old + 1
        S["sinceCheck"] += 1
        mustStop = not (old < maxIter)
        thr = max(tol, relTol * S["initRes"]) if relTol > 1e-20 else tol
        every = eis_check_interval(S["cRatio"] * np.sqrt(abs(g)) / thr)
        S["needCheck"] = int(mustStop or S["sinceCheck"] >= every)
        # k_eis_res + STEP_EIS_RES
        if S["needCheck"]:
            checks += 1
            tot = 0.0
            for r in range(N):
                acc = rh[r]
                for j in range(pv.nLower[r]):
                    e = pv.entry(r, j)
                    acc = acc + val[e] * rh[pv.col[e]]
                tot += abs(acc) / sv[r]
            S["finalRes"] = comm.allsum(tot) / nf
            S["converged"] = int(conv())
            cont = (S["nIter"] - 1 < maxIter and not S["converged"]) or S["nIter"] < minIter
            if not cont:
                S["done"] = 1
            a = np.sqrt(abs(S["wArA"]))
            S["cRatio"] = S["finalRes"] / a if a > 0 else 0.0
            S["sinceCheck"], S["needCheck"] = 0, 0
    # k_eis_final
    if not (S["nIter"] == 0 and not S["pendingPsi"]):
        x = xa + S["alpha"] * t if S["pendingPsi"] else xa
        psi = psi + sv * x
    return pv.to_natural(psi), S["nIter"], S["finalRes"], checks


def smooth_solve_emulated(pv, s, psi0, smoother="symGaussSeidel", tol=1e-6, relTol=0.0, maxIter=1000, minIter=0,
                          nSweeps=1, mode="multicolour", lag=True):
    """Kernel-by-kernel transliteration of b200_smooth_solve on one rank (solver.cu smooth_core, kernels.cuh
    k_fill_values_asym / k_gs_rows / k_gs_resid, STEP_NORM / STEP_GS_RES).  pv: a Levels plan with mode="exact"
    (OpenFOAM's own elimination order) or a MultiColour plan with mode="multicolour" (GS-class: the group an
    iteration updates last contributes its residual in-kernel; lag: two-colour plans take the other group's from its
    pass of the NEXT iteration, B200PCG_GS_LAGGED).  Returns (psi natural, nIter, initRes, finalRes)."""
    low = s.upper if s.lower is None else s.lower
    val = pv.values_asym(s.upper, low, s.addr.lowerAddr)
    d, b, x = pv.to_internal(s.diag), pv.to_internal(s.source), pv.to_internal(psi0)
    C = pv.nColours
    # two colours, symGaussSeidel, multicolour order: a counted sweep is executed as two red-black sweeps (solver.cu rb2)
    rb2 = smoother == "symGaussSeidel" and C == 2 and mode == "multicolour"
    inner = 2 if rb2 else 1
    back = smoother == "symGaussSeidel" and C >= 2 and not rb2
    fused = mode == "multicolour" and nSweeps > 0
    state = {"first": True}

    def sweep(res):
        k0 = 1 if (back and not state["first"]) else 0      # group 0 was the last group of the previous reverse half
        state["first"] = False
        tot = 0.0
        for k in range(k0, C):
            tot += pv.gs_rows(k, d, val, b, x, res and not back and k == C - 1)
        if back:
            for k in range(C - 2, -1, -1):          # group C-1 would be recomputed from unchanged inputs
                tot += pv.gs_rows(k, d, val, b, x, res and k == 0)
        return tot

    if nSweeps < 0:
        for _ in range(-nSweeps * inner):
            sweep(False)
        return pv.to_natural(x), -nSweeps, 0.0, 0.0
    wA = pv.spmv(d, val, x)
    sumA = pv.spmv(d, val, np.ones(pv.N))
    xRef = x.sum() / pv.N
    nf = (np.abs(wA - sumA * xRef) + np.abs(b - sumA * xRef)).sum() + 1e-20
    init = final = np.abs(b - wA).sum() / nf
    conv = lambda: final < tol or (relTol > 1e-20 and final < relTol * init)
    q0, q1 = 0, pv.N
    if fused:
        if back:
            q0 = int(pv.colourStart[1])
        else:
            q1 = int(pv.colourStart[C - 1])
        if C == 1:
            q0 = q1 = 0
    n = 0
    lagged = fused and C == 2 and lag
    if (minIter > 0 or not conv()) and not lagged:
        while True:
            tot = 0.0
            for sw in range(nSweeps * inner):
                tot += sweep(fused and sw == nSweeps * inner - 1)
            final = (tot + pv.gs_residual(d, val, b, x, q0, q1)) / nf
            n += nSweeps
            if not ((n < maxIter and not conv()) or n < minIter):
                break
    elif minIter > 0 or not conv():
        # two colours: an iteration after the first sweep is [group F, group L]; F's residual of iteration k comes out of
        # F's pass of iteration k + 1 (w - diag * x_old, k_gs_rows RES == 2), whose new values go to a shadow array
        gF = 1 if back else 0
        gL = 1 - gF
        rowsF = pv.rows_of_colour(gF)

        def f_pass_lagged():
            tot, new = 0.0, {}
            for r in rowsF:
                w = b[r]
                for j in range(pv.nTotal[r]):
                    e = pv.entry(r, j)
                    w = w - val[e] * x[pv.col[e]]
                new[r] = w / d[r]
                tot += abs(w - d[r] * x[r])
            return tot, new
        body = 0
        while True:
            tot = 0.0
            for sw in range(nSweeps * inner):
                res = sw == nSweeps * inner - 1
                if state["first"]:
                    tot += sweep(res)
                else:
                    if not (sw == 0 and body > 0):          # (the lagged pass below has already updated group F)
                        pv.gs_rows(gF, d, val, b, x)
                    tot += pv.gs_rows(gL, d, val, b, x, res)
            body += 1
            totF, newF = f_pass_lagged()                    # first pass of the next loop body
            final = (tot + totF) / nf
            n += nSweeps
            if not ((n < maxIter and not conv()) or n < minIter):
                break
            for r, v in newF.items():
                x[r] = v
    return pv.to_natural(x), n, init, final


# ---- numpy transliteration of the PBiCG + DILU path (solver.cu bicg_core, kernels.cuh k_dilu_calc_rd / k_tri_fwd /
#      k_tri_bwd / k_bicg_p / k_bicg_r / k_dot2) ---------------------------------------------------------------------
def dilu_calc_rd_emulated(pv, diag_i, val, valT):
    """k_dilu_calc_rd per colour / level + k_recip: rD[u] -= upper*lower/rD[l] over the row's earlier neighbours"""
    rD = np.empty(pv.N)
    for k in range(pv.nColours):
        for r in pv.rows_of_colour(k):
            d = diag_i[r]
            for j in range(pv.nLower[r]):
                e = pv.entry(r, j)
                d = d - (valT[e] * val[e]) / rD[pv.col[e]]
            rD[r] = d
    return 1.0 / rD


def pbicg_emulated(pv, s, psi0, precond="DILU", tol=1e-6, relTol=0.0, maxIter=1000, minIter=0):
    """PBiCG::solve on the plan's row structure: Levels plan = OpenFOAM's DILU (same elimination order, same in-row
    operation order), MultiColour plan = the DILU-class stand-in.  Returns (psi natural, nIter, initRes, finalRes)."""
    low = s.upper if s.lower is None else s.lower
    val = pv.values_asym(s.upper, low, s.addr.lowerAddr)        # A:   row entries of Amul / precondition
    valT = pv.values_asym(low, s.upper, s.addr.lowerAddr)       # A^T: row entries of Tmul / preconditionT
    d, b, x = pv.to_internal(s.diag), pv.to_internal(s.source), pv.to_internal(psi0)
    wA, wT = pv.spmv(d, val, x), pv.spmv(d, valT, x)
    sumA = pv.spmv(d, val, np.ones(pv.N))
    xRef = x.sum() / pv.N
    nf = (np.abs(wA - sumA * xRef) + np.abs(b - sumA * xRef)).sum() + 1e-20
    rA, rT = b - wA, b - wT
    init = final = np.abs(rA).sum() / nf
    conv = lambda: final < tol or (relTol > 1e-20 and final < relTol * init)
    n, singular = 0, False
    if minIter > 0 or not conv():
        if precond == "DILU":
            rD = dilu_calc_rd_emulated(pv, d, val, valT)
        elif precond == "diagonal":
            rD = 1.0 / d
        rho = 1e20
        pA = pT = None
        while True:
            rho_old = rho
            if precond == "DILU":
                wA, wT = pv.dic_precondition(rD, val, rA), pv.dic_precondition(rD, valT, rT)
            elif precond == "diagonal":
                wA, wT = rD * rA, rD * rT
            else:
                wA, wT = rA.copy(), rT.copy()
            rho = wA @ rT
            if n == 0:
                pA, pT = wA.copy(), wT.copy()
            else:
                beta = rho / rho_old
                pA, pT = wA + beta * pA, wT + beta * pT
            wA, wT = pv.spmv(d, val, pA), pv.spmv(d, valT, pT)
            wApT = wA @ pT
            if not (abs(wApT) / nf > 1e-300):
                singular = True
                break
            alpha = rho / wApT
            x, rA, rT = x + alpha * pA, rA - alpha * wA, rT - alpha * wT
            final = np.abs(rA).sum() / nf
            n += 1
            if not ((n - 1 < maxIter and not conv()) or n < minIter):
                break
    return pv.to_natural(x), n, init, final


def smooth_solve_emulated_ranks(pv, s, psi0, comm, bou, smoother="symGaussSeidel", tol=1e-6, relTol=0.0, maxIter=1000,
                                minIter=0, nSweeps=1, mode="multicolour"):
    """The multi-rank sequence of b200_smooth_solve (solver.cu smooth_core with GsHalo on; kernels.cuh k_pack ->
    exchange -> k_gs_bprime, k_gs_rows<HALO>, k_gs_resid<HALO>): processor patches are explicit contributions refreshed
    once per COUNTED sweep, the skips and the in-kernel residuals of the one-rank loop are off.  `comm`: allsum /
    exchange (tests/gloo_smooth_worker.py supplies the gloo twin), `bou`: interfaceBouCoeffs in slot order.
    Returns (psi natural, nIter, initRes, finalRes)."""
    low = s.upper if s.lower is None else s.lower
    val = pv.values_asym(s.upper, low, s.addr.lowerAddr)
    d, b, x = pv.to_internal(s.diag), pv.to_internal(s.source), pv.to_internal(psi0)
    C = pv.nColours
    rb2 = smoother == "symGaussSeidel" and C == 2 and mode == "multicolour"
    inner = 2 if rb2 else 1
    back = smoother == "symGaussSeidel" and C >= 2 and not rb2
    rowB = np.full(pv.N, -1, dtype=np.int64)
    rowB[pv.bRow] = np.arange(pv.bRow.size)

    def bprime():                                    # k_pack -> exchange -> k_gs_bprime
        recv = comm.exchange(x[pv.slotRow]) if pv.slotRow.size else np.empty(0)
        bp = b.copy()
        for i in range(pv.bRow.size):
            acc = b[pv.bRow[i]]
            for e in range(pv.bStart[i], pv.bStart[i + 1]):
                acc = acc + bou[pv.bSlot[e]] * recv[pv.bSlot[e]]
            bp[pv.bRow[i]] = acc
        return bp

    def counted_sweep():
        bp = bprime()                                # once per counted sweep, as upstream's bPrime
        for _ in range(inner):
            for k in range(C):
                pv.gs_rows(k, d, val, bp, x)
            if back:
                for k in range(C - 2, -1, -1):
                    pv.gs_rows(k, d, val, bp, x)

    def residual():                                  # k_pack -> exchange -> k_gs_resid<HALO>
        recv = comm.exchange(x[pv.slotRow]) if pv.slotRow.size else np.empty(0)
        tot = 0.0
        for r in range(pv.N):
            w = b[r] - d[r] * x[r]
            for j in range(pv.nTotal[r]):
                e = pv.entry(r, j)
                w = w - val[e] * x[pv.col[e]]
            if rowB[r] >= 0:
                for e in range(pv.bStart[rowB[r]], pv.bStart[rowB[r] + 1]):
                    w = w + bou[pv.bSlot[e]] * recv[pv.bSlot[e]]
            tot += abs(w)
        return comm.allsum(tot)

    if nSweeps < 0:
        for _ in range(-nSweeps):
            counted_sweep()
        return pv.to_natural(x), -nSweeps, 0.0, 0.0
    wA = _amul(pv, comm, d, val, bou, x)
    sumA = pv.spmv(d, val, np.ones(pv.N))
    if pv.slotRow.size:
        sumA = sumA - np.bincount(pv.slotRow, weights=bou, minlength=pv.N)
    xRef = comm.allsum(x.sum()) / comm.allsum(float(pv.N))
    nf = comm.allsum((np.abs(wA - sumA * xRef) + np.abs(b - sumA * xRef)).sum()) + 1e-20
    init = final = comm.allsum(np.abs(b - wA).sum()) / nf
    conv = lambda: final < tol or (relTol > 1e-20 and final < relTol * init)
    n = 0
    if minIter > 0 or not conv():
        while True:
            for _ in range(nSweeps):
                counted_sweep()
            final = residual() / nf
            n += nSweeps
            if not ((n < maxIter and not conv()) or n < minIter):
                break
    return pv.to_natural(x), n, init, final
