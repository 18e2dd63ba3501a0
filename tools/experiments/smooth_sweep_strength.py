"""CPU experiment behind a design decision of the smoothSolver path (DESIGN.md section 4 "smoothSolver", point 4):
how many sweeps does symGaussSeidel need in
  (a) upstream's natural cell order            (oracle/smooth_oracle.c: what `smoothSolver` does),
  (b) a red-black order, one red-black sweep per counted sweep   (what a literal multicolour symmetric sweep is:
      its reverse half only recomputes one colour),
  (c) a red-black order, TWO red-black sweeps per counted sweep  (what the library executes),
on U-shaped transport systems (cases.transport_system) of varying stiffness (kappa ~ 1 / Courant number) and
asymmetry (Peclet number), and on 4-colour polyhedral plans where the symmetric sweep is kept as it is.
usage: python tools/experiments/smooth_sweep_strength.py > profiles/r02_smooth_sweep_counts.txt"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers  # noqa: E402
from firefoam_dev_b200 import cases, meshgen  # noqa: E402
from oracle import oracle as orc  # noqa: E402

print("sweeps to tolerance 1e-8 (relTol 0, x0 = 0); smoother symGaussSeidel unless noted")
print(f"{'mesh':<22}{'kappa':>7}{'Pe':>5} | {'natural sym':>11} {'natural GS':>10} | {'rb x1':>6} {'rb x2 (lib)':>11} | colours")
for mesh, base in (("hex 14x12x10", meshgen.hex_block(14, 12, 10)), ("poly bcc 6x5x5", meshgen.bcc_poly(6, 5, 5))):
    for kappa in (1.0, 0.3, 0.1, 0.03):
        for pe in (0.5, 2.0, 8.0):
            s = cases.transport_system(base, peclet=pe, kappa=kappa, seed=11)
            N = s.addr.nCells
            psi = np.zeros(N)
            a = orc.smooth_solve(s, psi, smoother="symGaussSeidel", tolerance=1e-8, maxIter=5000).nIterations
            psi = np.zeros(N)
            g = orc.smooth_solve(s, psi, smoother="GaussSeidel", tolerance=1e-8, maxIter=5000).nIterations
            pv = helpers.PlanView(1, s.addr, renumber=-1)
            lib = helpers.smooth_solve_emulated(pv, s, np.zeros(N), tol=1e-8, maxIter=5000)[1]
            if pv.nColours == 2:
                # one red-black sweep per counted sweep == the multicolour GaussSeidel smoother on the same plan
                rb1 = helpers.smooth_solve_emulated(pv, s, np.zeros(N), smoother="GaussSeidel", tol=1e-8, maxIter=5000)[1]
            else:
                rb1 = "-"
            print(f"{mesh:<22}{kappa:>7}{pe:>5} | {a:>11} {g:>10} | {rb1!s:>6} {lib:>11} | {pv.nColours}")
