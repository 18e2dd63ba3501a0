"""Worker of tests/test_multigpu.py: one process per GPU (torchrun), NCCL halos + all-reduces
through libb200pcg, checked on rank 0 against the N-rank CPU oracle and the 1-rank oracle."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import firefoam_dev_b200 as pkg  # noqa: E402
from firefoam_dev_b200 import meshgen as mg  # noqa: E402
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

PROCS = {2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}[world]
# (B200PCG_TILE / B200PCG_SMALL_N are inherited from the environment: test_multigpu.py runs the worker both ways)
DIMS = (24, 20, 16)
# the Eisenstat form needs the colour-major plan (no B200PCG_TILE); it has run green on 2 GPUs -- more ranks
# are part of the first GPU call of round 2 (B200_TEST_UNVALIDATED=1)
EIS = not os.environ.get("B200PCG_TILE")
results = {"eisenstat_ran": EIS}
s = mg.hex_block(*DIMS, *PROCS, rank)
# collective safety of b200_set_addressing: ONE rank hands over a bad interface (neighbour rank out of range);
# every rank must come back with an error -- none may be left inside a collective -- and the context stays usable
from firefoam_dev_b200.ldu import LduAddressing, ProcessorLduInterface  # noqa: E402
bad = s.addr
if rank == world - 1 and s.addr.interfaces:
    itf = [ProcessorLduInterface(world + 3, i.faceCells) if k == 0 else i for k, i in enumerate(s.addr.interfaces)]
    bad = LduAddressing(s.addr.nCells, s.addr.lowerAddr, s.addr.upperAddr, itf)
try:
    ctx.set_addressing(bad)
    results["bad_rank_rejected_everywhere"] = False
except pkg.B200Error as e:
    results["bad_rank_rejected_everywhere"] = True
    results["bad_rank_message_rank0"] = str(e)
flags = [None] * world
dist.all_gather_object(flags, results["bad_rank_rejected_everywhere"])
results["bad_rank_rejected_everywhere"] = all(flags)
ctx.set_addressing(s.addr)

# Amul with halo exchange
x = np.random.default_rng(100 + rank).standard_normal(s.addr.nCells)
y = ctx.amul(s.matrix, s.bou, x)
ys = [torch.zeros(1)] * world
gather = [None] * world
dist.all_gather_object(gather, (x, y))
if rank == 0:
    subs = [mg.hex_block(*DIMS, *PROCS, r) for r in range(world)]
    ref = orc.amul(subs, [g[0] for g in gather])
    results["amul_bit_exact"] = all(np.array_equal(ref[r], gather[r][1]) for r in range(world))

for pre, exact in (("diagonal", False), ("none", False), ("DIC", True), ("DIC", "multicolour"), ("DIC", "eisenstat")):
    if exact == "eisenstat" and not EIS:
        continue
    ctl = {"preconditioner": pre, "tolerance": 1e-8, "relTol": 0.0, "maxIter": 3000}
    if exact:
        ctl["B200"] = {"dicMode": "exact" if exact is True else exact}
    psi = np.zeros(s.addr.nCells)
    perf = pkg.B200PCG("p_rgh", s.matrix, s.bou, None, s.interfaces, ctl, context=ctx).solve(psi, s.source)
    allpsi = [None] * world
    dist.all_gather_object(allpsi, psi)
    if rank == 0:
        key = pre + ("" if not exact or exact == "multicolour" else "-exact" if exact is True else "-" + exact)
        subs = [mg.hex_block(*DIMS, *PROCS, r) for r in range(world)]
        ref = [np.zeros(x_.addr.nCells) for x_ in subs]
        pr = orc.pcg_solve(subs, ref, "DIC" if pre == "DIC" else pre, 1e-8, 0.0, 3000)
        err = max(np.abs(a - b).max() for a, b in zip(allpsi, ref)) / max(np.abs(b).max() for b in ref)
        xerr = max(np.abs(a - x_.xstar).max() for a, x_ in zip(allpsi, subs))
        results[key] = {"iters": perf.nIterations, "oracle_iters": pr.nIterations, "relerr_vs_oracle": err,
                        "err_vs_xstar": xerr, "converged": bool(perf.converged),
                        "init": perf.initialResidual, "oracle_init": pr.initialResidual}
# polyhedral mesh, RCB ("scotch-class") decomposition: irregular processor patches, cells with
# several processor faces, several neighbours per rank
poly = mg.bcc_poly(10, 10, 12)
c2p = mg.partition_rcb(poly.xyz, world)
subs = mg.decompose(poly, c2p, world)
ps = subs[rank]
ctx.set_addressing(ps.addr)
for pre, exact in (("diagonal", False), ("DIC", True), ("DIC", "multicolour"), ("DIC", "eisenstat")):
    if exact == "eisenstat" and not EIS:
        continue
    ctl = {"preconditioner": pre, "tolerance": 1e-8, "relTol": 0.0, "maxIter": 3000}
    if exact:
        ctl["B200"] = {"dicMode": "exact" if exact is True else exact}
    psi = np.zeros(ps.addr.nCells)
    perf = pkg.B200PCG("p_rgh", ps.matrix, ps.bou, None, ps.interfaces, ctl, context=ctx).solve(psi, ps.source)
    allpsi = [None] * world
    dist.all_gather_object(allpsi, psi)
    if rank == 0:
        ref = [np.zeros(x_.addr.nCells) for x_ in subs]
        pr = orc.pcg_solve(subs, ref, pre, 1e-8, 0.0, 3000)
        err = max(np.abs(a - b).max() for a, b in zip(allpsi, ref)) / max(np.abs(b).max() for b in ref)
        results["poly-" + pre + ("" if not exact or exact == "multicolour" else "-exact" if exact is True else "-" + exact)] = {
            "converged": bool(perf.converged),
            "iters": perf.nIterations, "oracle_iters": pr.nIterations, "relerr_vs_oracle": err,
            "nbrs": [len(x_.bou) for x_ in subs]}
if rank == 0:
    print("MGPU_RESULT " + json.dumps(results))
ctx.close()
dist.destroy_process_group()
