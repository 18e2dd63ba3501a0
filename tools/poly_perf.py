"""Development aid: N fixed iterations of PCG on the 5 M-cell polyhedral workload (ncu target).
usage: poly_perf.py <preconditioner> [iters]"""
import sys

import numpy as np

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import firefoam_dev_b200 as pkg
from firefoam_dev_b200 import meshgen as mg

s = mg.bcc_poly(125, 125, 160)
ctx = pkg.Context(device=0)
ctx.set_addressing(s.addr)
pre = sys.argv[1] if len(sys.argv) > 1 else "diagonal"
ctl, _ = pkg.make_controls({"preconditioner": pre, "tolerance": 1e-6, "maxIter": 5000})
ctx.force_iterations(int(sys.argv[2]) if len(sys.argv) > 2 else 12)
for rep in range(2):
    psi = np.zeros(s.addr.nCells)
    perf = ctx.solve(s.diag, s.upper, [], s.source, psi, ctl)
print("ok", perf.nIterations, ctx.describe()["renumbered_rcm"], ctx.describe()["amul_natural"])
