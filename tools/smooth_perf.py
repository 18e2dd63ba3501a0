"""smoothSolver (and, with `bicg`, PBiCG) path on one GPU, device-resident: time per iteration and per kernel.
usage: python tools/smooth_perf.py NX NY NZ [poly] [exact] [bicg] [iters=N]"""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from firefoam_dev_b200 import Context, cases, meshgen  # noqa: E402
from firefoam_dev_b200.ldu import make_bicg_controls, make_smooth_controls  # noqa: E402

args = sys.argv[1:]
nx, ny, nz = (int(a) for a in args[:3])
poly = "poly" in args
exact = "exact" in args
iters = next((int(a.split("=")[1]) for a in args if a.startswith("iters=")), 20)
t0 = time.time()
base = meshgen.bcc_poly(nx, ny, nz) if poly else meshgen.hex_block(nx, ny, nz)
s = cases.transport_system(base, seed=31, kappa=0.3 if poly else 0.15)
N, F = s.addr.nCells, s.addr.nFaces
print(f"system: N={N} F={F} built in {time.time() - t0:.1f}s", flush=True)
ctx = Context(device=0)
ctx.set_addressing(s.addr)
dev = torch.device("cuda:0")
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
d, up, lo, b = t(s.diag), t(s.upper), t(s.lower), t(s.source)
out = {"N": N, "F": F, "mode": "exact" if exact else "multicolour", "poly": poly}
if "bicg" in args:
    for pre in ("DILU", "diagonal"):
        ctl, _ = make_bicg_controls(dict(preconditioner=pre, tolerance=1e-8, maxIter=1000,
                                         B200={"diluMode": "exact" if exact else "multicolour"}))
        x = torch.zeros(N, dtype=torch.float64, device=dev)
        p = ctx.bicg_solve_device(d, up, lo, b, x, ctl)            # warm-up + plan build
        res = {"to_tolerance": {"iters": p.nIterations, "final": p.finalResidual, "solve_ms": p.solveMs, "setup_ms": p.setupMs,
                                "colours": p.nColours, "err": float(np.abs(x.cpu().numpy() - s.xstar).max())}}
        ctx.force_iterations(iters)
        for prof in (False, True):
            ctx.profile(prof)
            x.zero_()
            p = ctx.bicg_solve_device(d, up, lo, b, x, ctl)
            key = "profiled" if prof else "timed"
            res[key] = {"iters": p.nIterations, "solve_ms": p.solveMs, "us_per_iter": 1e3 * p.solveMs / max(1, p.nIterations)}
            if prof:
                res[key]["kernels"] = {k: v for k, v in ctx.profile_json().items()
                                       if k.startswith("bicg_") or k in ("spmv_dot", "spmv_init")}
        ctx.profile(False)
        ctx.force_iterations(0)
        out["PBiCG+" + pre] = res
    print(json.dumps(out))
    sys.exit(0)
for smoother in ("symGaussSeidel", "GaussSeidel"):
    ctl, _, _ = make_smooth_controls(dict(smoother=smoother, tolerance=1e-8, maxIter=1000,
                                          B200={"sweepMode": "exact" if exact else "multicolour"}))
    x = torch.zeros(N, dtype=torch.float64, device=dev)
    p = ctx.smooth_solve_device(d, up, lo, [], b, x, ctl)          # warm-up + plan build
    res = {"to_tolerance": {"iters": p.nIterations, "final": p.finalResidual, "solve_ms": p.solveMs, "setup_ms": p.setupMs,
                            "colours": p.nColours}}
    ctx.force_iterations(iters)
    for prof in (False, True):
        ctx.profile(prof)
        x.zero_()
        p = ctx.smooth_solve_device(d, up, lo, [], b, x, ctl)
        key = "profiled" if prof else "timed"
        res[key] = {"iters": p.nIterations, "solve_ms": p.solveMs, "us_per_iter": 1e3 * p.solveMs / max(1, p.nIterations)}
        if prof:
            res[key]["kernels"] = {k: v for k, v in ctx.profile_json().items() if k.startswith("gs_")}
    ctx.profile(False)
    ctx.force_iterations(0)
    out[smoother] = res
print(json.dumps(out))
