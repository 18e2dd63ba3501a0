"""Worker of tests/test_multirank_gloo.py (CPU, gloo, world_size 2): the host-side logic of the
N > 1 path exercised across REAL processes -- per-rank sub-mesh generation, processor-patch pairing,
the pack list / interface CSR the CUDA kernels consume (csrc/plan.cpp), halo exchange ordering and
the reduction protocol -- with numpy standing in for the device kernels."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from firefoam_dev_b200 import meshgen as mg  # noqa: E402
from helpers import PlanView  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
DIMS, PROCS = (8, 6, 4), (2, 1, 1)
s = mg.hex_block(*DIMS, *PROCS, rank)
P = PlanView(0, s.addr)
val = P.values(s.upper)
bou = np.concatenate(s.bou) if s.bou else np.zeros(0)


def allsum(v):
    t = torch.tensor([v], dtype=torch.float64)
    dist.all_reduce(t)
    return float(t[0])


def amul(x):
    send = x[P.slotRow].copy()                       # k_pack
    recv = np.empty_like(send)
    reqs = []
    for k in range(len(P.nbrRank)):                  # grouped send/recv per neighbour
        a, b = P.patchStart[k], P.patchStart[k + 1]
        reqs.append(dist.isend(torch.from_numpy(send[a:b]), int(P.nbrRank[k])))
        reqs.append(dist.irecv(torch.from_numpy(recv[a:b]), int(P.nbrRank[k])))
    y = P.spmv(s.diag, val, x)                       # interior rows overlap the exchange
    for r in reqs:
        r.wait()
    for b in range(P.bRow.size):                     # k_iface_fix: sorted-segment reduction
        acc = y[P.bRow[b]]
        for e in range(P.bStart[b], P.bStart[b + 1]):
            acc = acc - bou[P.bSlot[e]] * recv[P.bSlot[e]]
        y[P.bRow[b]] = acc
    return y


N = s.addr.nCells
psi = np.zeros(N)
wA = amul(psi)
rA = s.source - wA
sA = P.spmv(s.diag, val, np.ones(N)) - np.bincount(P.slotRow, weights=bou, minlength=N)
xRef = allsum(psi.sum()) / allsum(float(N))
normFactor = allsum((np.abs(wA - xRef * sA) + np.abs(s.source - xRef * sA)).sum()) + 1e-20
init = final = allsum(np.abs(rA).sum()) / normFactor
rD = 1.0 / s.diag
nIter, wArA = 0, 1e20
pA = np.zeros(N)
while True:
    wArAold = wArA
    wA = rD * rA
    wArA = allsum((wA * rA).sum())
    pA = wA.copy() if nIter == 0 else wA + (wArA / wArAold) * pA
    wA = amul(pA)
    alpha = wArA / allsum((wA * pA).sum())
    psi += alpha * pA
    rA -= alpha * wA
    final = allsum(np.abs(rA).sum()) / normFactor
    nIter += 1
    if not ((nIter - 1) < 1000 and not final < 1e-8):
        break
out = [None] * world
dist.all_gather_object(out, (psi, [int(k) for k in P.nbrRank]))
if rank == 0:
    print("GLOO_RESULT " + json.dumps({"iters": nIter, "init": init, "final": final,
                                       "psi": [o[0].tolist() for o in out], "nbr": [o[1] for o in out]}))
dist.destroy_process_group()
