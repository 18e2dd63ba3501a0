"""Worker of tests/test_multirank_gloo.py::test_multi_rank_smooth_solver_gloo (CPU, gloo): the multi-rank sequence of
the smoothSolver path across REAL processes -- the exchange of psi once per counted sweep, bPrime of the interface rows
with the negated coupled coefficients, the residual with its own exchange -- with the numpy transliteration of
tests/helpers.py standing in for the device kernels, on the plan structures the CUDA kernels consume."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from firefoam_dev_b200 import cases, meshgen as mg  # noqa: E402
from helpers import PlanView, smooth_solve_emulated_ranks  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")


class GlooComm:
    def __init__(self, pv):
        self.pv = pv

    def allsum(self, v):
        t = torch.tensor([float(v)], dtype=torch.float64)
        dist.all_reduce(t)
        return float(t[0])

    def exchange(self, send):
        P = self.pv
        send = np.ascontiguousarray(send, dtype=np.float64)
        recv = np.empty_like(send)
        reqs = []
        for k in range(len(P.nbrRank)):
            a, b = int(P.patchStart[k]), int(P.patchStart[k + 1])
            if b == a:
                continue
            reqs.append(dist.isend(torch.from_numpy(send[a:b]), int(P.nbrRank[k])))
            reqs.append(dist.irecv(torch.from_numpy(recv[a:b]), int(P.nbrRank[k])))
        for r in reqs:
            r.wait()
        return recv


def systems():
    NX, NY, NZ = 8, 6, 4
    PX, PY, PZ = {2: (2, 1, 1), 4: (2, 2, 1)}[world]
    g = cases.transport_system(mg.hex_block(NX, NY, NZ), seed=21)
    c = np.arange(NX * NY * NZ)
    ix, iy, iz = c % NX, (c // NX) % NY, c // (NX * NY)
    c2p = (ix * PX // NX) + PX * ((iy * PY // NY) + PY * (iz * PZ // NZ))
    yield "hex", mg.decompose(g, c2p.astype(np.int32), world)[rank]
    poly = mg.bcc_poly(4, 3, 3, shuffle_block=64)
    pt = cases.transport_system(poly, seed=22, kappa=0.3)
    yield "poly", mg.decompose(pt, mg.partition_rcb(poly.xyz, world), world)[rank]


out = {}
for name, s in systems():
    bou = np.concatenate(s.bou) if s.bou else np.zeros(0)
    x0 = np.zeros(s.addr.nCells)
    res = {}
    for key, ordering, mode, kw in (("exact", 2, "exact", dict(tol=1e-8, maxIter=500)),
                                    ("exact_gs2", 2, "exact", dict(tol=1e-8, maxIter=500, nSweeps=2, smoother="GaussSeidel")),
                                    ("exact_fixed3", 2, "exact", dict(nSweeps=-3)),
                                    ("mc", 1, "multicolour", dict(tol=1e-11, maxIter=3000))):
        P = PlanView(ordering, s.addr)
        x, n, init, final = smooth_solve_emulated_ranks(P, s, x0, GlooComm(P), bou, mode=mode, **kw)
        res[key] = dict(psi=x, n=n, init=init, final=final, colours=int(P.nColours),
                        multiFaceRows=int((np.diff(P.bStart) > 1).sum()))
    allr = [None] * world
    dist.all_gather_object(allr, res)
    if rank == 0:
        out[name] = {k: dict(n=allr[0][k]["n"], init=allr[0][k]["init"], final=allr[0][k]["final"],
                             colours=[r[k]["colours"] for r in allr], multiFaceRows=[r[k]["multiFaceRows"] for r in allr],
                             psi=[r[k]["psi"].tolist() for r in allr]) for k in res}
if rank == 0:
    print("GLOO_SMOOTH_RESULT " + json.dumps(out))
dist.destroy_process_group()
