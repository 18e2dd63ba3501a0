#!/bin/bash
# ncu evidence for the non-headline kernels: DIC-class iteration on the 16 M hex box, Amul on 5 M polyhedra
# with and without the RCM renumbering (sectors per request).
cd "$(dirname "$0")/.."
cat > /tmp/poly_perf.py <<'PY'
import sys, numpy as np
sys.path.insert(0, ".")
import firefoam_dev_b200 as pkg
from firefoam_dev_b200 import meshgen as mg
s = mg.bcc_poly(125, 125, 160)
ctx = pkg.Context(device=0)
ctx.set_addressing(s.addr)
ctl, _ = pkg.make_controls({"preconditioner": sys.argv[1], "tolerance": 1e-6, "maxIter": 5000})
ctx.force_iterations(12)
for rep in range(2):
    psi = np.zeros(s.addr.nCells)
    perf = ctx.solve(s.diag, s.upper, [], s.source, psi, ctl)
print("ok", perf.nIterations, ctx.describe()["renumbered_rcm"])
PY
python tools/quick_perf.py 256 250 250 DIC 12 noconv > gpurun_out/ncu2_plain_hex.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_spmv|k_dic_fwd|k_dic_bwd|k_p|k_r" -s 60 -c 6 -o gpurun_out/prof_dic_hex python tools/quick_perf.py 256 250 250 DIC 12 noconv > gpurun_out/ncu2_hex.log 2>&1
python /tmp/poly_perf.py diagonal > gpurun_out/ncu2_plain_poly.log 2>&1 &&
ncu --set full --clock-control none -k regex:"k_spmv" -s 6 -c 2 -o gpurun_out/prof_poly_rcm python /tmp/poly_perf.py diagonal > gpurun_out/ncu2_poly_rcm.log 2>&1
B200PCG_RENUMBER=0 B200PCG_SPMV=ell python /tmp/poly_perf.py diagonal > gpurun_out/ncu2_plain_poly0.log 2>&1 &&
B200PCG_RENUMBER=0 B200PCG_SPMV=ell ncu --set full --clock-control none -k regex:"k_spmv" -s 6 -c 2 -o gpurun_out/prof_poly_nat python /tmp/poly_perf.py diagonal > gpurun_out/ncu2_poly_nat.log 2>&1
ls -la gpurun_out/*.ncu-rep
tail -2 gpurun_out/ncu2_hex.log gpurun_out/ncu2_poly_rcm.log gpurun_out/ncu2_poly_nat.log
