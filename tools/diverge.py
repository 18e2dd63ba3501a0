"""Development aid: residual after k iterations, GPU vs oracle, to tell rounding drift from bugs."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import firefoam_dev_b200 as pkg
from firefoam_dev_b200 import meshgen as mg
from oracle import oracle as orc
ctx = pkg.Context(device=0)
s = mg.hex_block(24, 20, 16)
for pre in ("none", "diagonal", "DIC"):
    for k in (1, 2, 5, 10, 20, 50, 100, 200, 400):
        psi = np.zeros(s.addr.nCells)
        ctl = dict(preconditioner=pre, tolerance=1e-30, maxIter=k - 1, B200={"dicMode": "exact"})
        pg = pkg.B200PCG("p", s.matrix, [], None, [], ctl, context=ctx).solve(psi, s.source)
        ref = np.zeros(s.addr.nCells)
        pc = orc.pcg_solve(s, ref, pre, 1e-30, 0.0, k - 1)
        print(f"{pre:8s} k={k:4d} it={pg.nIterations:4d}/{pc.nIterations:4d} res gpu={pg.finalResidual:.6e} cpu={pc.finalResidual:.6e} "
              f"rel={abs(pg.finalResidual-pc.finalResidual)/pc.finalResidual:.2e} xrel={np.abs(psi-ref).max()/np.abs(ref).max():.2e}")
