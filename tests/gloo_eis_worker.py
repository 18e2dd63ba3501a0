"""Worker of tests/test_multirank_gloo.py::test_two_rank_eisenstat_gloo (CPU, gloo, world_size 2): the
Eisenstat form of the DIC-class loop across REAL processes -- the exchange of the scaling vector s, the
scaled interface coefficients, the halo term B t inside the forward sweep, the overlapped form
(B200PCG_EIS_OVERLAP=1: interface rows of the first colour swept ahead of the exchange) -- with the numpy
transliteration of tests/helpers.py standing in for the device kernels, on the plan structures the CUDA
kernels consume (csrc/plan.cpp)."""
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
from helpers import PlanView, pcg_eisenstat_emulated, pcg_multicolour_reference  # noqa: E402

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
        for k in range(len(P.nbrRank)):                  # grouped send/recv per neighbour, slot order
            a, b = int(P.patchStart[k]), int(P.patchStart[k + 1])
            if b == a:
                continue
            reqs.append(dist.isend(torch.from_numpy(send[a:b]), int(P.nbrRank[k])))
            reqs.append(dist.irecv(torch.from_numpy(recv[a:b]), int(P.nbrRank[k])))
        for r in reqs:
            r.wait()
        return recv


def systems():
    procs = {2: (2, 1, 1), 4: (2, 2, 1)}[world]
    yield "hex", mg.hex_block(8, 6, 4, *procs, rank)
    poly = mg.bcc_poly(4, 3, 3, shuffle_block=64)
    c2p = mg.partition_rcb(poly.xyz, world)
    yield "poly", mg.decompose(poly, c2p, world)[rank]


out = {}
for name, s in systems():
    P = PlanView(1, s.addr)
    comm = GlooComm(P)
    bou = np.concatenate(s.bou) if s.bou else np.zeros(0)
    x0 = np.zeros(s.addr.nCells)
    kw = dict(tol=1e-9, maxIter=500, comm=comm, bou=bou)
    xr, nr, fr = pcg_multicolour_reference(P, s.diag, s.upper, s.source, x0, **kw)
    xe, ne, fe, ce = pcg_eisenstat_emulated(P, s.diag, s.upper, s.source, x0, overlap=False, **kw)
    xo, no, fo, co = pcg_eisenstat_emulated(P, s.diag, s.upper, s.source, x0, overlap=True, **kw)
    res = [None] * world
    dist.all_gather_object(res, dict(ref=xr, eis=xe, ovl=xo, colours=int(P.nColours),
                                     ifaceRows=int(P.bRow.size),
                                     multiFaceRows=int((np.diff(P.bStart) > 1).sum())))
    if rank == 0:
        out[name] = dict(iters=[nr, ne, no], final=[fr, fe, fo], checks=[ce, co],
                         colours=[r["colours"] for r in res], ifaceRows=[r["ifaceRows"] for r in res],
                         multiFaceRows=[r["multiFaceRows"] for r in res],
                         ref=[r["ref"].tolist() for r in res], eis=[r["eis"].tolist() for r in res],
                         ovl=[r["ovl"].tolist() for r in res])
if rank == 0:
    print("GLOO_EIS_RESULT " + json.dumps(out))
dist.destroy_process_group()
