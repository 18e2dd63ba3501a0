"""Development probe: polyhedral workload performance and small-case (steckler) latency, GPU vs the
CPU oracle on the same box."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import firefoam_dev_b200 as pkg
from firefoam_dev_b200 import meshgen as mg
from firefoam_dev_b200.cases import StecklerHydrostatic, steckler_p_rgh_system
from firefoam_dev_b200.meshgen import System
from oracle import oracle as orc

ctx = pkg.Context(device=0)
what = sys.argv[1] if len(sys.argv) > 1 else "all"
if what in ("all", "poly"):
    nx, ny, nz = (int(a) for a in (sys.argv[2:5] if len(sys.argv) > 4 else (100, 100, 128)))
    t0 = time.time(); s = mg.bcc_poly(nx, ny, nz); N, F = s.addr.nCells, s.addr.nFaces
    print(f"poly N={N} F={F} F/N={F/N:.2f} gen {time.time()-t0:.1f}s", flush=True)
    t0 = time.time(); ctx.set_addressing(s.addr); print(f"set_addressing {time.time()-t0:.1f}s", flush=True)
    for pre in ("diagonal", "DIC"):
        ctl, _ = pkg.make_controls({"preconditioner": pre, "tolerance": 1e-6, "maxIter": 5000})
        for rep in range(2):
            psi = np.zeros(N)
            ctx.profile(rep == 1)
            perf = ctx.solve(s.diag, s.upper, [], s.source, psi, ctl)
        it_bytes = (136 * N + 48 * F) if pre == "DIC" else (120 * N + 16 * F)
        print(f"poly {pre}: iters={perf.nIterations} solveMs={perf.solveMs:.2f} us/iter={1e3*perf.solveMs/perf.nIterations:.1f} "
              f"colours={perf.nColours} GDOF.iter/s={N*perf.nIterations/perf.solveMs/1e6:.2f} "
              f"iter alg GB/s={it_bytes*perf.nIterations/perf.solveMs/1e6:.0f} err={np.abs(psi-s.xstar).max():.2e}", flush=True)
        for k, v in ctx.profile_json().items():
            extra = ""
            if k == "spmv_dot":
                extra = f" algGB/s={(24*N+16*F)/v['avg_us']/1e3:.0f}"
            print(f"    {k:16s} n={v['launches']:6d} avg={v['avg_us']:9.2f}us{extra}")
        ctx.profile(False)
if what in ("all", "steckler"):
    case = StecklerHydrostatic()
    s = steckler_p_rgh_system()
    ctx.set_addressing(s.addr)
    for pre, exact in (("diagonal", False), ("DIC", False), ("DIC", True)):
        for rt in (0.01, 0.0):
            sc = {"preconditioner": pre, "tolerance": 1e-6, "relTol": rt, "B200": {"dicMode": "exact" if exact else "multicolour"}}
            solver = pkg.B200PCG("p_rgh", s.matrix, [], None, [], sc, context=ctx)
            ts = []
            for rep in range(6):
                psi = np.zeros(s.addr.nCells)
                t0 = time.perf_counter(); perf = solver.solve(psi, s.source); ts.append(time.perf_counter() - t0)
            tc = []
            for rep in range(3):
                ref = np.zeros(s.addr.nCells)
                t0 = time.perf_counter(); pc = orc.pcg_solve(s, ref, pre, 1e-6, rt); tc.append(time.perf_counter() - t0)
            print(f"steckler p_rgh {pre}{'-exact' if exact else ''} relTol={rt}: gpu iters={perf.nIterations} wall={1e3*min(ts):.3f}ms "
                  f"(solve {perf.solveMs:.3f} setup {perf.setupMs:.3f} h2d {perf.h2dMs:.3f} d2h {perf.d2hMs:.3f}) | "
                  f"cpu(1 core) iters={pc.nIterations} wall={1e3*min(tc):.3f}ms", flush=True)
