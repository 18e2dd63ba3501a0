#!/usr/bin/env python
"""bench.py -- p_rgh pressure-correction hot path on N B200s of one node (BASELINE.json metric:
p_rgh PCG solve time and GDOF.iter/s, HBM GB/s vs peak).

  python bench.py [--gpus N] [--steps K] [--warmup W]            (N > 1: launched by torchrun)
  python bench.py --impl reference ...                           (CPU arm: the oracle on host cores)

A "step" is one pass of the hot path over one synthetic system: assemble
`- fvm::laplacian(rhorAUf, p_rgh)` into LDU form (face coefficients + negSumDiag) and solve it with
PCG + diagonal preconditioner to tolerance 1e-6 (relTol 0, x0 = 0) -- SURVEY.md 8d config 3.
Default workload (weak scaling): every GPU owns a 256 x 250 x 250 = 16 M-cell block of a uniform hex
box decomposed `hierarchical` (1 1 1)/(2 1 1)/(2 2 1)/(2 2 2); at N = 8 that is BASELINE config 4's
128 M-cell mesh.  Processor-patch halos go over NCCL send/recv, the CG scalars over a peer-memory
one-shot all-reduce (NCCL all-reduce as fall-back).

Other workloads (not the driver's default):
  --scaling strong     BASELINE config 4's 512 x 500 x 500 = 128 M-cell mesh, fixed, decomposed N-way
                       (N = 1 runs all 128 M cells on one B200: the strong-scaling base)
  --workload poly      config 5: BCC polyhedral mesh (`--poly NX NY NZ` lattice, default 250 250 320 =
                       40 M cells, 14 faces/cell, Morton + block-shuffled numbering), RCB N-way
  --workload steckler  config 2: synthetic p_rgh system on the 9 000-cell steckler topology, 1 GPU
                       (launch-latency-bound)

value  = N_global * (PCG iterations executed) / device time, inputs resident in HBM.
e2e    = the same metric through the reference-facing plug-in call B200PCG.solve() with pinned HOST
         buffers (diag/upper/source/psi/interfaceBouCoeffs H2D and psi D2H inside the timed region).
"""
import argparse
import json
import os
import pickle
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault("NCCL_DEBUG", "WARN")      # keep NCCL's version banner off stdout (one JSON line)
# torchrun exports OMP_NUM_THREADS=1; the host-side set-up (plan build, workload generation) is
# OpenMP code that should use the cores this rank is entitled to
if os.environ.get("OMP_NUM_THREADS", "1") == "1":
    _n = max(1, (os.cpu_count() or 1) // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1"))))
    os.environ["OMP_NUM_THREADS"] = str(min(_n, 32))

BLOCK = (256, 250, 250)          # cells per GPU (config 3)
STRONG_DIMS = (512, 500, 500)    # config 4: 128 M cells, fixed
POLY_LATTICE = (250, 250, 320)   # config 5: 2 * 250 * 250 * 320 = 40 M cells
PROCS = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}
TOL, MAXITER = 1e-6, 5000
METRIC = "p_rgh PCG solve throughput (assemble + PCG/diagonal to tol 1e-6), fp64"
PRECOND = "diagonal"   # --precond overrides (DIC = multicolour IC0, BASELINE config 4)
UNIT = "GDOF*iter/s"


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        if sm:
            sm.sort()
            out = {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                   "samples": len(sm)}
        return out


# ---- workloads ---------------------------------------------------------------------------------
def hex_layout(args, n):
    """(global dims, procs) of the hex workload at n GPUs."""
    p = PROCS[n]
    if args.scaling == "strong":
        return tuple(args.block) if args.block else STRONG_DIMS, p
    b = tuple(args.block) if args.block else BLOCK
    return tuple(bb * q for bb, q in zip(b, p)), p


def config_dict(args, n):
    pre = args.precond
    if args.workload == "hex":
        dims, procs = hex_layout(args, n)
        cells = dims[0] * dims[1] * dims[2]
        reduced = bool(args.block)
        if args.scaling == "strong":
            base = "configs[3] (128M hex decomposed N-way), strong scaling; N=1 is the single-GPU base"
        else:
            base = "configs[2] (16M hex, PCG+diagonal, 1xB200) per GPU; N=8 is configs[3]'s 128M mesh"
        return {"workload": f"hex{dims[0]}x{dims[1]}x{dims[2]}_p_rgh_PCG_{pre}" + ("_reduced" if reduced else ""),
                "cells": cells, "cells_per_gpu": cells // n,
                "decomposition": f"hierarchical ({procs[0]} {procs[1]} {procs[2]})",
                "preconditioner": pre, "tolerance": TOL, "relTol": 0.0, "maxIter": MAXITER,
                "baseline_config": base,
                "l2": "working set (>1.5 GB per GPU) exceeds the 126 MB L2; no flush needed"}
    if args.workload == "poly":
        nx, ny, nz = args.poly
        cells = 2 * nx * ny * nz
        return {"workload": f"bcc_poly_2x{nx}x{ny}x{nz}_p_rgh_PCG_{pre}"
                            + ("" if tuple(args.poly) == POLY_LATTICE else "_reduced"),
                "cells": cells, "cells_per_gpu": cells // n,
                "decomposition": f"RCB {n}-way (scotch stand-in)" if n > 1 else "none",
                "preconditioner": pre, "tolerance": TOL, "relTol": 0.0, "maxIter": MAXITER,
                "baseline_config": "configs[4] (40M polyhedral, truncated octahedra, 14 faces/cell, Morton + "
                                   "4096-block shuffled numbering)",
                "l2": "working set exceeds the 126 MB L2 at >= 1 M cells per GPU; no flush needed"}
    return {"workload": f"steckler_9000_p_rgh_PCG_{pre}", "cells": 9000, "cells_per_gpu": 9000,
            "decomposition": "none", "preconditioner": pre, "tolerance": TOL, "relTol": 0.0,
            "maxIter": MAXITER,
            "baseline_config": "configs[1] (steckler topology 30x15x20 with compartment baffles, synthetic "
                               "p_rgh-shaped system; the reference ships no per-time-step matrices)",
            "l2": "the whole system (<2 MB) is L2-resident: launch-latency-bound, not HBM-bound; a "
                  "256 MB buffer is written between timed steps to flush L2"}


def make_system(args, n, rank, pinned=None, dist=None):
    """This rank's System (meshgen.System) + global cell count + host generation seconds."""
    from firefoam_dev_b200 import meshgen as mg
    t0 = time.time()
    if args.workload == "hex":
        dims, procs = hex_layout(args, n)
        s = mg.hex_block(*dims, *procs, rank, pinned=pinned)
        return s, dims[0] * dims[1] * dims[2], time.time() - t0
    if args.workload == "steckler":
        from firefoam_dev_b200 import cases
        s = cases.steckler_p_rgh_system()
        return s, s.addr.nCells, time.time() - t0
    # poly: rank 0 generates and decomposes (decomposePar-style), the others load their part
    nx, ny, nz = args.poly
    if n == 1:
        s = mg.bcc_poly(nx, ny, nz)
        return s, s.addr.nCells, time.time() - t0
    box = [None]
    cache = args.poly_cache and os.path.join(args.poly_cache, f"b200poly_{nx}_{ny}_{nz}_{n}")
    if rank == 0:
        if cache and os.path.exists(os.path.join(cache, "DONE")):
            box[0] = cache
        else:
            # the other ranks wait: give the generator every host core (torchrun pins OMP to 1/N of them)
            try:
                import ctypes
                gomp = ctypes.CDLL("libgomp.so.1")
                nthr_old = gomp.omp_get_max_threads()
                gomp.omp_set_num_threads(min(os.cpu_count() or 1, 64))
            except Exception:
                gomp = None
            full = mg.bcc_poly(nx, ny, nz)
            c2p = mg.partition_rcb(full.xyz, n)
            subs = mg.decompose(full, c2p, n)
            del full
            if gomp is not None:
                gomp.omp_set_num_threads(nthr_old)
            if cache:
                os.makedirs(cache, exist_ok=True)
                d = cache
            else:
                d = tempfile.mkdtemp(prefix="b200poly_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
            for r, sub in enumerate(subs):
                with open(os.path.join(d, f"rank{r}.pkl"), "wb") as f:
                    pickle.dump(sub, f, protocol=4)
            open(os.path.join(d, "DONE"), "w").close()
            box[0] = d
            del subs
    dist.broadcast_object_list(box, src=0)
    with open(os.path.join(box[0], f"rank{rank}.pkl"), "rb") as f:
        s = pickle.load(f)
    dist.barrier()
    if rank == 0 and not cache:
        import shutil
        shutil.rmtree(box[0], ignore_errors=True)
    return s, 2 * nx * ny * nz, time.time() - t0


# ---------------------------------------------------------------------------------------------
def cpu_leg(args, seconds=12.0, steps=1, warmup=0):
    """The reference's CPU path for this workload: oracle/ (plain-C restatement of OpenFOAM-dev's
    PCG + diagonalPreconditioner/DICPreconditioner + Amul + normFactor, `kind: port` -- the
    reference's own implementation is un-vendored and cannot be built here), R emulated ranks = R
    host threads, each owning a decomposePar sub-mesh, like `mpirun -np R fireFoam -parallel`
    (cases/wallFireSpread2D/runParallel.sh:18).  Bounded sample: a fixed number of PCG iterations on
    one GPU's share of the workload, sized for ~`seconds` of CPU work per step."""
    import numpy as np
    from firefoam_dev_b200 import meshgen as mg
    from oracle import oracle as orc
    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    pre = "DIC" if args.precond.startswith("DIC") else args.precond
    R = 1
    while R * 2 <= min(ncpu, 128):
        R *= 2
    t0 = time.time()
    if args.workload == "hex":
        block = tuple(args.block) if args.block else BLOCK
        p = [1, 1, 1]
        d, r = 0, R
        while r > 1:
            p[d % 3] *= 2
            r //= 2
            d += 1
        subs = [mg.hex_block(*block, *p, rank) for rank in range(R)]
        what = (f"the {block[0]}x{block[1]}x{block[2]} hex system on {R} emulated ranks/threads "
                f"({p[0]} {p[1]} {p[2]})")
    elif args.workload == "poly":
        nx, ny, nz = args.poly
        while 2 * nx * ny * nz > 6_000_000:      # bounded sample: <= 6 M cells of the same lattice
            nx, ny, nz = max(8, nx // 2), max(8, ny // 2), max(8, nz // 2)
        full = mg.bcc_poly(nx, ny, nz)
        subs = mg.decompose(full, mg.partition_rcb(full.xyz, R), R) if R > 1 else [full]
        what = f"the 2x{nx}x{ny}x{nz} BCC polyhedral system, RCB on {R} emulated ranks/threads"
    else:
        from firefoam_dev_b200 import cases
        subs, R = [cases.steckler_p_rgh_system()], 1
        what = "the 9000-cell steckler p_rgh system on 1 thread (a 9000-cell case does not scale over ranks)"
    n_global = sum(s.addr.nCells for s in subs)
    gen_s = time.time() - t0

    def run(iters):
        psis = [np.zeros(s.addr.nCells) for s in subs]
        t = time.perf_counter()
        perf = orc.pcg_solve(subs, psis, pre, 1e-30, 0.0, maxIter=iters - 1)
        return time.perf_counter() - t, perf.nIterations
    tc, ic = run(4)                                   # calibration (also warms the page cache)
    per_iter = tc / max(ic, 1)
    cap = 400 if args.workload != "steckler" else 200000
    iters = int(min(cap, max(8, seconds / max(per_iter, 1e-7))))
    times, total_it = [], 0
    for _ in range(warmup):
        run(iters)
    for _ in range(max(1, steps)):
        t, it = run(iters)
        times.append(t)
        total_it += it
    value = n_global * total_it / sum(times) / 1e9
    return {"value": value, "unit": UNIT, "cores": R, "kind": "port",
            "sample": f"{iters} PCG+{pre} iterations of {what} ({n_global} cells); "
                      f"host has {ncpu} usable cores; {sum(times)/len(times):.2f} s per step",
            "ms_per_step": 1e3 * sum(times) / len(times), "gen_s": gen_s}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    leg = cpu_leg(args, seconds=args.cpu_seconds, steps=args.steps, warmup=min(args.warmup, 1))
    line = {"impl": "reference", "metric": METRIC.replace("diagonal", args.precond), "value": leg["value"],
            "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": leg["ms_per_step"], "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(args, args.gpus),
            "cpu_baseline": {k: leg[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": leg["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------
def run_gpu(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import firefoam_dev_b200 as pkg

    n = args.gpus
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != n:
        raise SystemExit(f"--gpus {n} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {n}")
    if not torch.cuda.is_available() or pkg.load_pcg().b200_device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    uid = None
    if n > 1:
        dist.init_process_group("nccl", device_id=dev)
        buf = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            buf.copy_(torch.frombuffer(bytearray(pkg.Context.unique_id()), dtype=torch.uint8))
        dist.broadcast(buf, 0)
        uid = bytes(buf.cpu().numpy().tobytes())

    def barrier():
        if n > 1:
            dist.barrier()
        torch.cuda.synchronize()

    keep = []

    def pinned(count, dtype):
        t = torch.empty(max(1, count), dtype=torch.float64, pin_memory=True)
        keep.append(t)
        return t.numpy()[:count]

    def to_pinned(a):
        out = pinned(a.size, np.float64)
        out[:] = a
        return out

    s, n_global, gen_s = make_system(args, n, rank, pinned=pinned, dist=dist)
    if args.workload != "hex":       # generators without a pinned allocator: page-lock what e2e reads
        s.diag, s.upper, s.source = to_pinned(s.diag), to_pinned(s.upper), to_pinned(s.source)
        s.bou = [to_pinned(b) for b in s.bou]
    a = s.addr
    N, F = a.nCells, a.nFaces
    ctx = pkg.Context(device=local_rank, rank=rank, nranks=n, nccl_uid=uid)
    t0 = time.time()
    ctx.set_addressing(a)
    setaddr_s = time.time() - t0
    sctl = {"preconditioner": "DIC" if args.precond.startswith("DIC") else args.precond, "tolerance": TOL,
            "relTol": 0.0, "maxIter": MAXITER,
            "B200": {"dicMode": {"DIC-exact": "exact", "DIC-eisenstat": "eisenstat"}.get(args.precond, "multicolour")}}
    ctl, _ = pkg.make_controls(sctl)

    up = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    d_gamma, d_magSf, d_delta = up(s.gamma_f), up(s.magSf), up(s.deltaCoeffs)
    d_diag0, d_src = up(s.diag0), up(s.source)
    d_bou = [up(b) for b in s.bou]
    d_diag = torch.empty(N, dtype=torch.float64, device=dev)
    d_upper = torch.empty(F, dtype=torch.float64, device=dev)
    d_psi = torch.zeros(N, dtype=torch.float64, device=dev)
    small = N * 200 < 126e6          # whole system L2-resident: flush between timed steps
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if small else None

    def step_device():
        # assemble: - fvm::laplacian(rhorAUf, p_rgh) on top of the ddt/boundary diagonal
        d_diag.copy_(d_diag0)
        d_psi.zero_()
        if flush is not None:
            flush.zero_()
        torch.cuda.current_stream().synchronize()
        ctx.assemble_laplacian_device(d_gamma, d_magSf, d_delta, s.sign, d_upper, d_diag)
        return ctx.solve_device(d_diag, d_upper, d_bou, d_src, d_psi, ctl)

    for _ in range(args.warmup):
        perf = step_device()
    # ---- timed: device-resident ------------------------------------------------------------------
    clocks = ClockSampler(local_rank)
    ctx.profile(True)
    barrier()
    if rank == 0:
        clocks.start()
    l0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    iters = 0
    solve_ms = setup_ms = 0.0
    for _ in range(args.steps):
        perf = step_device()
        iters += perf.nIterations
        solve_ms += perf.solveMs
        setup_ms += perf.setupMs
    e1.record()
    barrier()
    launches = ctx.launch_count() - l0
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if n > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    clk = clocks.stop() if rank == 0 else None
    prof = ctx.profile_json()
    ctx.profile(False)
    converged = bool(perf.converged)
    err = float(np.abs(d_psi.cpu().numpy() - s.xstar).max()) if s.xstar is not None else None
    value = n_global * iters / (ms * 1e-3) / 1e9

    # ---- e2e: the reference-facing plug-in call with pinned host buffers ------------------------
    psi_h = pinned(N, np.float64)
    src_h = s.source
    solver = pkg.B200PCG("p_rgh", s.matrix, s.bou, None, s.interfaces, sctl, context=ctx)
    psi_h[:] = 0.0
    solver.solve(psi_h, src_h)                          # warm-up (allocates staging)
    barrier()
    t0 = time.perf_counter()
    e_iters = 0
    h2d_ms = d2h_ms = 0.0
    for _ in range(args.steps):
        psi_h[:] = 0.0
        p = solver.solve(psi_h, src_h)
        e_iters += p.nIterations
        h2d_ms += p.h2dMs
        d2h_ms += p.d2hMs
    barrier()
    e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if n > 1:
        dist.all_reduce(e_s, op=dist.ReduceOp.MAX)
    e_s = float(e_s.item())
    e2e_value = n_global * e_iters / e_s / 1e9
    nslots = sum(b.size for b in s.bou)
    h2d_bytes = 8 * (F + 3 * N + nslots)
    d2h_bytes = 8 * N
    e2e_err = float(np.abs(psi_h - s.xstar).max()) if s.xstar is not None else None

    # ---- fixed 200-iteration timing -------------------------------------
    ctx.force_iterations(200)
    step_device()
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    pf = step_device()
    f1.record()
    barrier()
    ctx.force_iterations(0)
    fixed_ms = f0.elapsed_time(f1)
    desc = ctx.describe()

    if rank != 0:
        if n > 1:
            dist.destroy_process_group()
        return 0

    peak, peak_src = measured_peak()
    dic = args.precond.startswith("DIC")
    alg = {"spmv_dot": 24 * N + 16 * F, "precond_dot": 24 * N,
           "asm_face_coeff": 32 * F, "asm_neg_sum_diag": 16 * N + 16 * F,
           # the fused vector kernels move fewer bytes than the unfused loop SURVEY.md 8d counts:
           # k_p: psi, pA read+write, rA (+rD | z) read; k_r: rA read+write, wA (+rD) read
           "p_psi_update": (40 if dic else 48) * N, "r_update_dots": (24 if dic else 32) * N,
           "flux": 16 * F + 8 * N}
    kernels = {}
    for k, v in prof.items():
        ent = {"launches": v["launches"], "avg_us": v["avg_us"]}
        if k in alg:
            ent["alg_bytes"] = alg[k]
            ent["gbs"] = alg[k] / (v["avg_us"] * 1e-6) / 1e9
            ent["frac"] = ent["gbs"] / peak
        kernels[k] = ent
    if dic and "dic_fwd" in kernels and "dic_bwd" in kernels:
        # one preconditioner apply = all forward + backward colour launches: 40N + 32F
        napply = max(1, kernels["dic_bwd"]["launches"] // max(1, perf.nColours - 1))
        t_us = (prof["dic_fwd"]["total_ms"] + prof["dic_bwd"]["total_ms"]) * 1e3 / napply
        kernels["dic_apply"] = {"launches": napply, "avg_us": t_us, "alg_bytes": 40 * N + 32 * F,
                                "gbs": (40 * N + 32 * F) / (t_us * 1e-6) / 1e9,
                                "frac": (40 * N + 32 * F) / (t_us * 1e-6) / 1e9 / peak}
    dom = kernels.get("spmv_dot", {})
    eis = "eis_fwd_dot" in kernels
    if eis:
        # Eisenstat form: per iteration the backward + forward sweeps do the work of Amul AND the
        # preconditioner apply: (24N + 16F) + (40N + 32F) algorithmic bytes, one entry
        nit = max(1, kernels["eis_r_update_rho"]["launches"])
        t_us = (prof["eis_fwd_dot"]["total_ms"] + prof.get("eis_bwd", {"total_ms": 0.0})["total_ms"]) * 1e3 / nit
        ab = 64 * N + 48 * F
        kernels["eis_sweeps"] = {"launches": nit, "avg_us": t_us, "alg_bytes": ab, "gbs": ab / (t_us * 1e-6) / 1e9,
                                 "frac": ab / (t_us * 1e-6) / 1e9 / peak}
        dom = kernels["eis_sweeps"]
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "spmv_traffic.json")) as f:
            tj = json.load(f)
            if tj.get("cells") == N and tj.get("kernel", "").split("<")[0] == desc.get("amul_natural", "").split("<")[0] \
                    and not dic:
                traffic = tj.get("dram_bytes_per_launch")
    except Exception:
        pass
    # algorithmic bytes of one PCG iteration: diagonal 120N+16F, DIC-class 136N+48F
    iter_bytes = (136 * N + 48 * F) if dic else (120 * N + 16 * F)
    amul_kernel = desc.get("amul_permuted" if dic else "amul_natural", "?")
    dom_name = (f"k_eis_bwd + k_eis_fwd per iteration (Eisenstat form: lduMatrix::Amul + DIC-class apply in two sweeps, "
                f"fused with gSumProd(wA,pA))" if eis else f"{amul_kernel} (lduMatrix::Amul fused with gSumProd(wA,pA))")
    line = {
        "metric": METRIC.replace("diagonal", args.precond), "value": value, "unit": UNIT, "n_gpus": n, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": dict(config_dict(args, n), colours=perf.nColours, cells_this_rank=N, faces_this_rank=F,
                       halo_slots_this_rank=nslots),
        "iterations_per_step": iters // args.steps, "converged": converged, "max_err_vs_xstar": err,
        "solve_ms_per_step": solve_ms / args.steps, "setup_ms_per_step": setup_ms / args.steps,
        "pcg_iteration": {"avg_us": 1e3 * solve_ms / max(iters, 1), "alg_bytes": iter_bytes,
                          "gbs": iter_bytes / (1e-3 * solve_ms / max(iters, 1)) / 1e9,
                          "frac": iter_bytes / (1e-3 * solve_ms / max(iters, 1)) / 1e9 / peak},
        "fixed_200_iterations": {"ms": fixed_ms, "iters": pf.nIterations,
                                 "gdof_iter_per_s": n_global * pf.nIterations / (fixed_ms * 1e-3) / 1e9},
        "roofline": {"kernel": dom_name,
                     "bound": "hbm", "achieved": dom.get("gbs"), "peak": peak, "unit": "GB/s",
                     "frac": dom.get("frac"), "traffic": traffic, "peak_source": peak_src,
                     "alg_bytes_per_launch": dom.get("alg_bytes"), "avg_us": dom.get("avg_us")},
        "kernels": kernels, "plan": desc,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
                "d2h_bytes_per_step": d2h_bytes, "ms_per_step": 1e3 * e_s / args.steps,
                "h2d_ms_per_step": h2d_ms / args.steps, "d2h_ms_per_step": d2h_ms / args.steps,
                "iterations_per_step": e_iters // args.steps, "max_err_vs_xstar": e2e_err,
                "timing": "host wall clock around blocking B200PCG.solve() calls, max over ranks"},
        "gpu_launches": launches, "clocks": clk,
        "host": {"gen_s": gen_s, "set_addressing_s": setaddr_s},
    }
    if n == 1 and not args.no_cpu_baseline:
        leg = cpu_leg(args, seconds=args.cpu_seconds)
        line["cpu_baseline"] = {k: leg[k] for k in ("value", "unit", "cores", "kind", "sample")}
    else:
        line["cpu_baseline"] = None
    print(json.dumps(line))
    if n > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="hex", choices=["hex", "poly", "steckler"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="hex only: weak = 16M cells per GPU (default), strong = the 128M mesh split N-way")
    ap.add_argument("--block", type=int, nargs=3, default=None,
                    help="development only: cells per GPU (weak) / global mesh (strong) instead of the BASELINE sizes")
    ap.add_argument("--poly", type=int, nargs=3, default=list(POLY_LATTICE), help="BCC lattice of --workload poly")
    ap.add_argument("--poly-cache", default=None,
                    help="directory in which the decomposed --workload poly sub-meshes are kept between runs")
    ap.add_argument("--precond", default=PRECOND, choices=["none", "diagonal", "DIC", "DIC-exact", "DIC-eisenstat"])
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.gpus not in PROCS:
        raise SystemExit("--gpus must be 1, 2, 4 or 8")
    if args.workload != "hex":
        if args.scaling == "strong" and args.workload == "steckler":
            raise SystemExit("--workload steckler is a single fixed 9000-cell system")
        args.scaling = "strong" if args.workload == "poly" else "weak"
        if args.workload == "steckler" and args.gpus != 1:
            raise SystemExit("--workload steckler runs on 1 GPU")
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
