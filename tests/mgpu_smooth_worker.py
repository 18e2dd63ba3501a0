"""Worker of tests/test_multigpu.py::test_multigpu_smooth_parity: one process per GPU (torchrun).  smoothSolver on
asymmetric transport matrices with processor patches (explicit, refreshed once per sweep -- upstream's
GaussSeidelSmoother treats them like that) against the N-rank CPU oracle (oracle/smooth_oracle.c, one pthread per
rank) with the same decomposition."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import firefoam_dev_b200 as pkg  # noqa: E402
from firefoam_dev_b200 import cases, meshgen as mg  # noqa: E402
from oracle import oracle as orc  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
buf = torch.zeros(128, dtype=torch.uint8, device=dev)
if rank == 0:
    buf.copy_(torch.frombuffer(bytearray(pkg.Context.unique_id()), dtype=torch.uint8))
dist.broadcast(buf, 0)
ctx = pkg.Context(device=local, rank=rank, nranks=world, nccl_uid=buf.cpu().numpy().tobytes())

PX, PY, PZ = {2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}[world]
NX, NY, NZ = 24, 20, 16
results = {}


def run_case(tag, subs):
    me = subs[rank]
    ctx.set_addressing(me.addr)
    # asymmetric Amul with halo exchange: bit-identical
    x = np.random.default_rng(100 + rank).standard_normal(me.addr.nCells)
    y = ctx.amul_asym(me.matrix, me.bou, x)
    gather = [None] * world
    dist.all_gather_object(gather, (x, y))
    if rank == 0:
        ref = orc.amul_asym(subs, [g[0] for g in gather])
        results[tag + "amul_bit_exact"] = all(np.array_equal(ref[r], gather[r][1]) for r in range(world))
    cases_ = [("exact-fixed3", dict(smoother="symGaussSeidel", nSweeps=-3, B200={"sweepMode": "exact"})),
              ("exact-sym", dict(smoother="symGaussSeidel", tolerance=1e-8, maxIter=500, B200={"sweepMode": "exact"})),
              ("exact-gs-nsweeps2", dict(smoother="GaussSeidel", tolerance=1e-8, maxIter=500, nSweeps=2, B200={"sweepMode": "exact"})),
              ("exact-U-controls", dict(smoother="symGaussSeidel", tolerance=1e-6, maxIter=10, B200={"sweepMode": "exact"})),
              ("mc-sym", dict(smoother="symGaussSeidel", tolerance=1e-11, maxIter=3000)),
              ("mc-gs-nsweeps3", dict(smoother="GaussSeidel", tolerance=1e-11, maxIter=3000, nSweeps=3))]
    for key, ctl in cases_:
        psi = np.zeros(me.addr.nCells)
        perf = pkg.B200smoothSolver("U", me.matrix, me.bou, None, me.interfaces, ctl, context=ctx).solve(psi, me.source)
        allpsi = [None] * world
        dist.all_gather_object(allpsi, psi)
        if rank == 0:
            o = {k: v for k, v in ctl.items() if k != "B200"}
            ref = [np.zeros(s_.addr.nCells) for s_ in subs]
            if not key.startswith("exact"):
                o.update(tolerance=1e-13, maxIter=5000, nSweeps=1)      # the converged solution of the same system
            pr = orc.smooth_solve(subs, ref, **o)
            err = max(np.abs(a - b).max() for a, b in zip(allpsi, ref)) / max(np.abs(b).max() for b in ref)
            results[tag + key] = {"iters": perf.nIterations, "oracle_iters": pr.nIterations,
                                  "bit_identical": all(np.array_equal(a, b) for a, b in zip(allpsi, ref)),
                                  "relerr_vs_oracle": err, "converged": bool(perf.converged),
                                  "init": perf.initialResidual, "oracle_init": pr.initialResidual,
                                  "final": perf.finalResidual, "oracle_final": pr.finalResidual}


g = cases.transport_system(mg.hex_block(NX, NY, NZ), seed=21)
c = np.arange(NX * NY * NZ)
ix, iy, iz = c % NX, (c // NX) % NY, c // (NX * NY)
c2p = (ix * PX // NX) + PX * ((iy * PY // NY) + PY * (iz * PZ // NZ))
run_case("hex-", mg.decompose(g, c2p.astype(np.int32), world))
poly = mg.bcc_poly(10, 10, 12)
pt = cases.transport_system(poly, seed=22, kappa=0.3)
run_case("poly-", mg.decompose(pt, mg.partition_rcb(poly.xyz, world), world))
if rank == 0:
    print("MGPU_SMOOTH_RESULT " + json.dumps(results))
ctx.close()
dist.destroy_process_group()
