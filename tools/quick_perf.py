"""Quick single-GPU performance probe (development aid): per-kernel-class CUDA-event times for one
solve of the hex workload.  usage: quick_perf.py NX NY NZ precond [iters]"""
import json
import sys
import time

import numpy as np

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import firefoam_dev_b200 as pkg
from firefoam_dev_b200 import meshgen as mg

NX, NY, NZ = (int(a) for a in sys.argv[1:4])
pre = sys.argv[4] if len(sys.argv) > 4 else "diagonal"
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 200
exact = pre == "DIC-exact"
mode = {"DIC-exact": "exact", "DIC-eisenstat": "eisenstat", "DIC-multicolour": "multicolour"}.get(pre, "auto")
t0 = time.time()
s = mg.hex_block(NX, NY, NZ)
print(f"generated N={s.addr.nCells} F={s.addr.nFaces} in {time.time()-t0:.1f}s", flush=True)
ctx = pkg.Context(device=0)
t0 = time.time()
ctx.set_addressing(s.addr)
print(f"set_addressing {time.time()-t0:.2f}s", flush=True)
ctl, _ = pkg.make_controls({"preconditioner": "DIC" if pre.startswith("DIC") else pre, "tolerance": 1e-6, "maxIter": 5000,
                            "B200": {"dicMode": mode}})
N, F = s.addr.nCells, s.addr.nFaces
for rep in range(3):
    psi = np.zeros(N)
    ctx.force_iterations(iters)
    ctx.profile(rep == 2)
    t0 = time.time()
    perf = ctx.solve(s.diag, s.upper, [], s.source, psi, ctl)
    wall = time.time() - t0
    print(f"rep{rep}: iters={perf.nIterations} solveMs={perf.solveMs:.3f} setupMs={perf.setupMs:.3f} "
          f"h2d={perf.h2dMs:.2f} d2h={perf.d2hMs:.2f} wall={wall*1e3:.1f}ms  per-iter={perf.solveMs/max(1,perf.nIterations)*1e3:.1f}us "
          f"colours={perf.nColours}", flush=True)
prof = ctx.profile_json()
bytes_per = {"spmv_dot": 24 * N + 16 * F, "precond_dot": 24 * N, "p_psi_update": 48 * N, "r_update_dots": 32 * N,
             "dic_fwd": (20 * N + 16 * F), "dic_bwd": (20 * N + 16 * F),
             # Eisenstat form: each sweep does half of Amul + half of the preconditioner apply
             "eis_bwd": 32 * N + 24 * F, "eis_fwd_dot": 32 * N + 24 * F, "eis_p_psi_update": 40 * N,
             "eis_r_update_rho": 24 * N}
for k, v in prof.items():
    line = f"  {k:16s} n={v['launches']:6d} avg={v['avg_us']:9.2f}us"
    if k in bytes_per:
        line += f"  algGB/s={bytes_per[k]/v['avg_us']/1e3:8.1f} frac={bytes_per[k]/v['avg_us']/1e3/6538:.3f}"
    print(line)
ctx.force_iterations(0)
ctx.profile(False)
if len(sys.argv) > 6 and sys.argv[6] == "noconv":
    sys.exit(0)
psi = np.zeros(N)
t0 = time.time()
perf = ctx.solve(s.diag, s.upper, [], s.source, psi, ctl)
print(f"to tolerance: {perf.nIterations} iterations, solveMs={perf.solveMs:.2f}, converged={perf.converged}, "
      f"final={perf.finalResidual:.3e}, err vs x*={np.abs(psi-s.xstar).max():.3e}, "
      f"GDOF.iter/s={N*perf.nIterations/perf.solveMs/1e6:.2f}")
