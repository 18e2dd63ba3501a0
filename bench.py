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
    fvm::laplacian assembly + PCG + diagonalPreconditioner/DICPreconditioner + Amul + normFactor,
    `kind: port` -- the reference's own implementation is un-vendored and cannot be built here), R emulated
    ranks = R host threads, each owning a decomposePar sub-mesh, like `mpirun -np R fireFoam -parallel`
    (cases/wallFireSpread2D/runParallel.sh:18).  Bounded sample: ONE GPU's share of the workload (at N > 1
    that is not the N-GPU mesh: `sample` names the mesh actually solved), assembled (laplacian face
    coefficients + negSumDiag) and iterated a fixed number of PCG iterations, sized for ~`seconds` of CPU work
    per step."""
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
        mesh = f"hex {block[0]}x{block[1]}x{block[2]}"
        what = (f"the {block[0]}x{block[1]}x{block[2]} hex system (one GPU's block of the workload) on {R} emulated "
                f"ranks/threads ({p[0]} {p[1]} {p[2]})")
    elif args.workload == "poly":
        nx, ny, nz = args.poly
        while 2 * nx * ny * nz > 6_000_000:      # bounded sample: <= 6 M cells of the same lattice
            nx, ny, nz = max(8, nx // 2), max(8, ny // 2), max(8, nz // 2)
        full = mg.bcc_poly(nx, ny, nz)
        subs = mg.decompose(full, mg.partition_rcb(full.xyz, R), R) if R > 1 else [full]
        mesh = f"bcc_poly 2x{nx}x{ny}x{nz}"
        what = f"the 2x{nx}x{ny}x{nz} BCC polyhedral system, RCB on {R} emulated ranks/threads"
    else:
        from firefoam_dev_b200 import cases
        subs, R = [cases.steckler_p_rgh_system()], 1
        mesh = "steckler 30x15x20"
        what = "the 9000-cell steckler p_rgh system on 1 thread (a 9000-cell case does not scale over ranks)"
    n_global = sum(s.addr.nCells for s in subs)
    gen_s = time.time() - t0

    def assemble():
        # fvm::laplacian: upper = sign * deltaCoeffs * (gamma * magSf), negSumDiag on top of diag0
        t = time.perf_counter()
        for s in subs:
            if s.gamma_f is None:
                continue
            a = s.addr
            up, dg = orc.laplacian_assemble(a.lowerAddr, a.upperAddr, a.nCells, s.gamma_f, s.magSf, s.deltaCoeffs,
                                            s.sign, s.diag0)
            s.upper[:], s.diag[:] = up, dg
        return time.perf_counter() - t

    def run(iters):
        psis = [np.zeros(s.addr.nCells) for s in subs]
        t = time.perf_counter()
        perf = orc.pcg_solve(subs, psis, pre, 1e-30, 0.0, maxIter=iters - 1)
        return time.perf_counter() - t, perf.nIterations
    tc, ic = run(4)                                   # calibration (also warms the page cache)
    per_iter = tc / max(ic, 1)
    cap = 400 if args.workload != "steckler" else 200000
    iters = int(min(cap, max(8, seconds / max(per_iter, 1e-7))))
    times, asm, total_it = [], [], 0
    for _ in range(warmup):
        assemble()
        run(iters)
    for _ in range(max(1, steps)):
        ta = assemble()
        t, it = run(iters)
        times.append(ta + t)
        asm.append(ta)
        total_it += it
    value = n_global * total_it / sum(times) / 1e9
    return {"value": value, "unit": UNIT, "cores": R, "kind": "port", "mesh": mesh, "cells": n_global,
            "sample": f"laplacian assembly (single thread per sub-mesh, {sum(asm)/len(asm):.2f} s) + {iters} PCG+{pre} "
                      f"iterations of {what} ({n_global} cells); host has {ncpu} usable cores; "
                      f"{sum(times)/len(times):.2f} s per step",
            "ms_per_step": 1e3 * sum(times) / len(times), "gen_s": gen_s}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    leg = cpu_leg(args, seconds=args.cpu_seconds, steps=args.steps, warmup=min(args.warmup, 1))
    cfg = config_dict(args, args.gpus)
    cfg["reference_sample_mesh"] = f"{leg['mesh']} ({leg['cells']} cells): a rate on one GPU's share, not the {cfg['cells']}-cell job"
    line = {"impl": "reference", "metric": METRIC.replace("diagonal", args.precond), "value": leg["value"],
            "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": leg["ms_per_step"], "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": cfg,
            "cpu_baseline": {k: leg[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": leg["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------
class Env:
    """Process-wide state of the GPU arm: rank / device / NCCL rendez-vous, pinned allocator, barrier."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        import firefoam_dev_b200 as pkg
        self.torch, self.dist, self.pkg, self.args = torch, dist, pkg, args
        self.n = args.gpus
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if os.environ.get("B200_BENCH_REVERSE_DEVICES"):      # experiment: does a slow rank follow the GPU or the sub-mesh?
            self.local = int(os.environ.get("WORLD_SIZE", "1")) - 1 - self.local
        world = int(os.environ.get("WORLD_SIZE", "1"))
        if world != self.n:
            raise SystemExit(f"--gpus {self.n} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {self.n}")
        if not torch.cuda.is_available() or pkg.load_pcg().b200_device_count() < 1:
            raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        self.uid = None
        if self.n > 1:
            dist.init_process_group("nccl", device_id=self.dev)
            self.uid = self.new_uid()
        self.keep = []
        self.peak, self.peak_src = measured_peak()

    def new_uid(self):
        torch, dist = self.torch, self.dist
        buf = torch.zeros(128, dtype=torch.uint8, device=self.dev)
        if self.rank == 0:
            buf.copy_(torch.frombuffer(bytearray(self.pkg.Context.unique_id()), dtype=torch.uint8))
        dist.broadcast(buf, 0)
        return bytes(buf.cpu().numpy().tobytes())

    def barrier(self):
        if self.n > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def pinned(self, count, dtype=None):
        t = self.torch.empty(max(1, count), dtype=self.torch.float64, pin_memory=True)
        self.keep.append(t)
        return t.numpy()[:count]

    def to_pinned(self, a):
        out = self.pinned(a.size)
        out[:] = a
        return out

    def up(self, x):
        import numpy as np
        return self.torch.from_numpy(np.ascontiguousarray(x)).to(self.dev)

    def max_over_ranks(self, v):
        t = self.torch.tensor([float(v)], dtype=self.torch.float64, device=self.dev)
        if self.n > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed(self, fn, steps):
        """barrier + sync, `steps` calls of fn() between two CUDA events on torch's current stream (every
        library call ends with a synchronize of its own stream, so the events bracket the whole device work),
        barrier + sync; returns (max over ranks of the elapsed ms, list of fn's results)."""
        torch = self.torch
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = [fn() for _ in range(steps)]
        e1.record()
        self.barrier()
        return self.max_over_ranks(e0.elapsed_time(e1)), out


class Resident:
    """One system resident in HBM + the device-resident step (assemble laplacian, solve)."""

    def __init__(self, env, ctx, s, sctl):
        import numpy as np
        torch = env.torch
        self.env, self.ctx, self.s = env, ctx, s
        a = s.addr
        self.N, self.F = a.nCells, a.nFaces
        self.ctl, _ = env.pkg.make_controls(sctl)
        up = env.up
        self.d_gamma, self.d_magSf, self.d_delta = up(s.gamma_f), up(s.magSf), up(s.deltaCoeffs)
        self.d_diag0, self.d_src = up(s.diag0), up(s.source)
        self.d_bou = [up(b) for b in s.bou]
        self.d_diag = torch.empty(self.N, dtype=torch.float64, device=env.dev)
        self.d_upper = torch.empty(self.F, dtype=torch.float64, device=env.dev)
        self.d_psi = torch.zeros(self.N, dtype=torch.float64, device=env.dev)
        small = self.N * 200 < 126e6          # whole system L2-resident: flush between timed steps
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=env.dev) if small else None
        self.np = np

    def step(self, ctl=None):
        # assemble: - fvm::laplacian(rhorAUf, p_rgh) on top of the ddt/boundary diagonal, then solve
        self.d_diag.copy_(self.d_diag0)
        self.d_psi.zero_()
        if self.flush is not None:
            self.flush.zero_()
        self.env.torch.cuda.current_stream().synchronize()
        self.ctx.assemble_laplacian_device(self.d_gamma, self.d_magSf, self.d_delta, self.s.sign, self.d_upper, self.d_diag)
        return self.ctx.solve_device(self.d_diag, self.d_upper, self.d_bou, self.d_src, self.d_psi, ctl or self.ctl)

    def err(self):
        return float(self.np.abs(self.d_psi.cpu().numpy() - self.s.xstar).max()) if self.s.xstar is not None else None


def sctl_for(precond, relTol=0.0, maxIter=MAXITER):
    mode = {"DIC-exact": "exact", "DIC-eisenstat": "eisenstat", "DIC-multicolour": "multicolour"}.get(precond, "auto")
    return {"preconditioner": "DIC" if precond.startswith("DIC") else precond, "tolerance": TOL, "relTol": relTol,
            "maxIter": maxIter, "B200": {"dicMode": mode}}


def kernel_table(prof, N, F, peak, dic, nColours, col_bytes=4):
    """Per-kernel averages of a PROFILED pass (launches of surplus loop bodies excluded by the library) against
    the bytes each kernel has to move: SURVEY.md 8d's algorithmic figure where it names the kernel, else the
    bytes of the layout the kernel streams (stated in DESIGN.md section 4) -- always a fraction <= 1."""
    alg = {"spmv_dot": 24 * N + 16 * F, "precond_dot": 24 * N,
           "asm_face_coeff": 32 * F, "asm_neg_sum_diag": 16 * N + 16 * F,
           # the fused vector kernels move fewer bytes than the unfused loop SURVEY.md 8d counts:
           # k_p: psi, pA read+write, rA (+rD | z) read; k_r: rA read+write, wA (+rD) read
           "p_psi_update": (40 if dic else 48) * N, "r_update_dots": (24 if dic else 32) * N,
           "flux": 16 * F + 8 * N,
           # Eisenstat form, bytes MOVED by each kernel: p^ update 44 N; r^ update 28 N
           "eis_p_psi_update": 44 * N, "eis_r_update_rho": 28 * N}
    kernels = {}
    for k, v in prof.items():
        ent = {"launches": v["launches"], "avg_us": v["avg_us"]}
        if k in alg and v["avg_us"] > 0:
            ent["alg_bytes"] = alg[k]
            ent["gbs"] = alg[k] / (v["avg_us"] * 1e-6) / 1e9
            ent["frac"] = ent["gbs"] / peak
        kernels[k] = ent
    if dic and "dic_fwd" in kernels and "dic_bwd" in kernels:
        # one preconditioner apply = all forward + backward colour launches: 40N + 32F
        napply = max(1, kernels["dic_bwd"]["launches"] // max(1, nColours - 1))
        t_us = (prof["dic_fwd"]["total_ms"] + prof["dic_bwd"]["total_ms"]) * 1e3 / napply
        ab = 40 * N + 32 * F
        kernels["dic_apply"] = {"launches": napply, "avg_us": t_us, "alg_bytes": ab, "gbs": ab / (t_us * 1e-6) / 1e9,
                                "frac": ab / (t_us * 1e-6) / 1e9 / peak}
    if "eis_fwd_dot" in kernels:
        # Eisenstat form: the backward + forward sweeps of one iteration replace Amul AND the preconditioner
        # apply.  Roofline on the bytes the two sweeps MOVE (ncu: 763 + 771 MB at 16 M hex
        # cells = 96 N): every entry of the full-row ELL once per iteration (value 8 + column `col_bytes`, 2 F
        # entries), 28 B of row streams per swept row (row length, p^ read, two vectors written; a sweep skips
        # the colour that has no neighbours on its side: N (C-1)/C rows each) and one read of every gathered
        # value (8 N).  Against SURVEY.md's UNFUSED figure for what the sweeps replace (64N + 48F) the number would
        # exceed 1: that figure is reported as `replaces_alg_bytes`, not as a roofline fraction.
        nit = max(1, kernels["eis_r_update_rho"]["launches"])
        t_us = (prof["eis_fwd_dot"]["total_ms"] + prof.get("eis_bwd", {"total_ms": 0.0})["total_ms"]) * 1e3 / nit
        C_ = max(2, nColours)
        ab = int(2 * F * (8 + col_bytes) + 28 * 2 * N * (C_ - 1) / C_ + 8 * N)
        kernels["eis_sweeps"] = {"launches": nit, "avg_us": t_us, "alg_bytes": ab, "gbs": ab / (t_us * 1e-6) / 1e9,
                                 "frac": ab / (t_us * 1e-6) / 1e9 / peak, "bytes": "moved by the layout (DESIGN.md 4)",
                                 "replaces_alg_bytes": 64 * N + 48 * F}
    return kernels


def col_bytes_of(ctx):
    """bytes per column index the ELL-bound kernels read on the multicolour plan (16-bit offsets when every
    slice entry fits, plan.hpp)"""
    return 2 if ctx.describe().get("ell_col16_fraction_multicolour", 0.0) == 1.0 else 4


def profiled_pass(res, ctl=None):
    """one step with per-kernel CUDA events (outside every headline timed region)"""
    res.ctx.profile(True)
    perf = res.step(ctl)
    prof = res.ctx.profile_json()
    res.ctx.profile(False)
    return prof, perf


def mgpu_parity(env):
    """N > 1: the N-rank path against the N-rank CPU oracle (emulated ranks, rank-ascending reductions) on a
    reduced mesh with the SAME decomposition, run after the timed regions so that the driver's scaling runs
    carry multi-GPU parity evidence (processor-patch halos over NCCL, fused peer-memory all-reduce):
    Amul bit-exact, identical iteration counts for diagonal and DIC-exact, DIC-class within 1e-6."""
    import numpy as np
    from firefoam_dev_b200 import meshgen as mg
    from oracle import oracle as orc
    dist, pkg, n, rank = env.dist, env.pkg, env.n, env.rank
    procs = PROCS[n]
    dims = (24, 20, 16)
    ctx = pkg.Context(device=env.local, rank=rank, nranks=n, nccl_uid=env.new_uid())
    out = {"mesh": f"hex {dims[0]}x{dims[1]}x{dims[2]} hierarchical ({procs[0]} {procs[1]} {procs[2]}) + bcc_poly 2x10x10x12 RCB {n}-way",
           "oracle": f"{n} emulated ranks (oracle/pcg_oracle.c, oracle/smooth_oracle.c)"}
    try:
        poly = mg.bcc_poly(10, 10, 12)
        polysubs = mg.decompose(poly, mg.partition_rcb(poly.xyz, n), n)
        for tag, subs_of in (("hex", lambda: [mg.hex_block(*dims, *procs, r) for r in range(n)]), ("poly", lambda: polysubs)):
            subs = subs_of()
            s = subs[rank]
            ctx.set_addressing(s.addr)
            x = np.random.default_rng(100 + rank).standard_normal(s.addr.nCells)
            y = ctx.amul(s.matrix, s.bou, x)
            gather = [None] * n
            dist.all_gather_object(gather, (x, y))
            if rank == 0:
                ref = orc.amul(subs, [g[0] for g in gather])
                out[tag + "_amul_bit_exact"] = bool(all(np.array_equal(ref[r], gather[r][1]) for r in range(n)))
            for pre, mode in (("diagonal", None), ("DIC", "exact"), ("DIC", "auto")):
                ctl = {"preconditioner": pre, "tolerance": 1e-8, "relTol": 0.0, "maxIter": 3000}
                if mode:
                    ctl["B200"] = {"dicMode": mode}
                psi = np.zeros(s.addr.nCells)
                perf = pkg.B200PCG("p_rgh", s.matrix, s.bou, None, s.interfaces, ctl, context=ctx).solve(psi, s.source)
                allpsi = [None] * n
                dist.all_gather_object(allpsi, psi)
                if rank == 0:
                    ref = [np.zeros(x_.addr.nCells) for x_ in subs]
                    pr = orc.pcg_solve(subs, ref, pre, 1e-8, 0.0, 3000)
                    err = max(np.abs(a - b).max() for a, b in zip(allpsi, ref)) / max(np.abs(b).max() for b in ref)
                    key = f"{tag}_{pre}" + ("" if not mode else "_exact" if mode == "exact" else "_class")
                    out[key] = {"iters": perf.nIterations, "oracle_iters": pr.nIterations, "relerr": float(err),
                                "converged": bool(perf.converged)}
        # smoothSolver + symGaussSeidel on an asymmetric transport matrix with the same (irregular, RCB)
        # processor patches: level-scheduled sweeps against the N-rank oracle (oracle/smooth_oracle.c) -- identical
        # sweep count, bit-identical psi; reported under its own key with its own verdict.  Every rank agrees on the
        # outcome of each call before the next collective, so a rank-local failure is reported, never waited for.
        from firefoam_dev_b200 import cases
        pt = cases.transport_system(poly, seed=22, kappa=0.3)
        tsubs = mg.decompose(pt, mg.partition_rcb(poly.xyz, n), n)
        s = tsubs[rank]
        sm, failed = {}, None

        def agreed(ok, why):
            flag = env.torch.tensor([1 if ok else 0], device=env.dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            return None if int(flag.item()) == 1 else (why or "another rank failed")
        try:
            ctx.set_addressing(s.addr)
            failed = agreed(True, None)
        except pkg.B200Error as e:
            failed = agreed(False, str(e))
        for key, ctl in (("exact", dict(smoother="symGaussSeidel", tolerance=1e-8, maxIter=500, B200={"sweepMode": "exact"})),
                         ("multicolour", dict(smoother="symGaussSeidel", tolerance=1e-11, maxIter=3000))):
            if failed:
                break
            psi = np.zeros(s.addr.nCells)
            try:
                perf = pkg.B200smoothSolver("U", s.matrix, s.bou, None, s.interfaces, ctl, context=ctx).solve(psi, s.source)
                failed = agreed(True, None)
            except pkg.B200Error as e:
                failed = agreed(False, str(e))
            if failed:
                break
            allpsi = [None] * n
            dist.all_gather_object(allpsi, psi)
            if rank == 0:
                ref = [np.zeros(x_.addr.nCells) for x_ in tsubs]
                o = dict(smoother="symGaussSeidel", tolerance=1e-8, maxIter=500) if key == "exact" else \
                    dict(smoother="symGaussSeidel", tolerance=1e-13, maxIter=5000)
                pr = orc.smooth_solve(tsubs, ref, **o)
                err = max(np.abs(a - b).max() for a, b in zip(allpsi, ref)) / max(np.abs(b).max() for b in ref)
                sm[key] = {"sweeps": perf.nIterations, "oracle_sweeps": pr.nIterations, "relerr": float(err),
                           "bit_identical": bool(all(np.array_equal(a, b) for a, b in zip(allpsi, ref))),
                           "converged": bool(perf.converged)}
        if failed:
            out["smooth_solver"] = {"pass": False, "error": failed}
        elif rank == 0:
            sm["pass"] = bool(sm["exact"]["bit_identical"] and sm["exact"]["sweeps"] == sm["exact"]["oracle_sweeps"]
                              and sm["multicolour"]["converged"] and sm["multicolour"]["relerr"] < 1e-8)
            out["smooth_solver"] = sm
        if rank == 0:
            strict = [k for k in out if k.endswith("_diagonal") or k.endswith("_DIC_exact")]
            out["iters_equal"] = bool(all(out[k]["iters"] == out[k]["oracle_iters"] for k in strict))
            out["relerr"] = max(out[k]["relerr"] for k in strict)
            out["amul_bit_exact"] = bool(out["hex_amul_bit_exact"] and out["poly_amul_bit_exact"])
            out["dic_class_relerr"] = max(out[k]["relerr"] for k in out if k.endswith("_DIC_class"))
            out["pass"] = bool(out["amul_bit_exact"] and out["iters_equal"] and out["relerr"] < 1e-11
                               and out["dic_class_relerr"] < 1e-6)
    finally:
        ctx.close()
    return out


def dic_class_section(env, res, n_global):
    """BASELINE configs[3]: the same mesh, `preconditioner DIC` (DIC-class multicolour IC0, Eisenstat form) to
    the same tolerance; time-to-tolerance beside GDOF*iter/s (the iteration count differs from diagonal's)."""
    ctl, _ = env.pkg.make_controls(sctl_for("DIC"))
    res.step(ctl)
    res.step(ctl)
    ms, perfs = env.timed(lambda: res.step(ctl), 2)
    iters = sum(p.nIterations for p in perfs)
    prof, perf = profiled_pass(res, ctl)
    kern = kernel_table(prof, res.N, res.F, env.peak, True, perf.nColours, col_bytes_of(res.ctx))
    # (eis_iface_rows: the first colour's interface rows swept ahead of / after the exchange, N > 1 only; iface_fix
    # here is k_eis_halo, the halo term B t with its flag wait)
    keep = {k: v for k, v in kern.items() if k.startswith("eis_") or k.startswith("dic_") or k in ("spmv_dot", "iface_fix")}
    solve_ms = sum(p.solveMs for p in perfs) / 2
    return {"preconditioner": "DIC (DIC-class: multicolour IC0, Eisenstat form; log name DIC(mc)B200PCG)",
            "value": n_global * iters / (ms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": ms / 2,
            "iterations_per_step": iters // 2, "colours": perf.nColours, "converged": bool(perf.converged),
            "time_to_tolerance_ms": solve_ms + sum(p.setupMs for p in perfs) / 2,
            "us_per_iteration": 1e3 * solve_ms / max(1, iters // 2),
            "max_err_vs_xstar": res.err(), "kernels": keep}


def corrector_section(env, ctx, s):
    """A real corrector is short: cases/steckler/original/linux64/log.fireFoam:220-223 runs 9-28 DICPCG
    iterations per p_rgh solve (relTol 0.01 / p_rghFinal).  Here: `preconditioner DIC`, relTol 0.01, at most 20
    iterations, on the full-size system, timed three ways through the C ABI: host page-locked arrays, host
    PAGEABLE arrays (what OpenFOAM's fields are; staged through page-locked pieces), and device-resident
    (b200_assemble_laplacian_device + b200_solve_device: no per-solve H2D at all)."""
    import numpy as np
    pkg = env.pkg
    sctl = sctl_for("DIC", relTol=0.01, maxIter=19)
    N, F = s.addr.nCells, s.addr.nFaces
    out = {"controls": "preconditioner DIC; tolerance 1e-6; relTol 0.01; maxIter 19 (<= 20 loop bodies)"}
    res = Resident(env, ctx, s, sctl)
    res.step()
    ms, perfs = env.timed(res.step, 3)
    out["device_resident_ms"] = ms / 3
    out["iterations"] = perfs[-1].nIterations
    psi = env.pinned(N)
    solver = pkg.B200PCG("p_rgh", s.matrix, s.bou, None, s.interfaces, sctl, context=ctx)

    def host_steps(mat_solver, psi_arr, src, reps=3):
        """wall clock of the blocking solve() calls only (the harness's re-zeroing of psi is not part of a solve)"""
        tot, perfs_ = 0.0, []
        for _ in range(reps):
            psi_arr[:] = 0.0
            t0 = time.perf_counter()
            perfs_.append(mat_solver.solve(psi_arr, src))
            tot += time.perf_counter() - t0
        return 1e3 * tot / reps, perfs_
    host_steps(solver, psi, s.source, 1)
    out["host_pinned_ms"], pp = host_steps(solver, psi, s.source)
    out["host_pinned_h2d_ms"], out["host_pinned_d2h_ms"] = pp[-1].h2dMs, pp[-1].d2hMs
    # pageable copies of the same arrays
    from firefoam_dev_b200.ldu import LduMatrix
    pg = LduMatrix(s.addr, np.array(s.diag), np.array(s.upper))
    src_pg, psi_pg = np.array(s.source), np.zeros(N)
    solver_pg = pkg.B200PCG("p_rgh", pg, [np.array(b) for b in s.bou], None, s.interfaces, sctl, context=ctx)
    host_steps(solver_pg, psi_pg, src_pg, 1)
    out["host_pageable_ms"], pq = host_steps(solver_pg, psi_pg, src_pg)
    out["host_pageable_h2d_ms"], out["host_pageable_d2h_ms"] = pq[-1].h2dMs, pq[-1].d2hMs
    h2d_bytes = 8 * (F + 3 * N)
    out["h2d_bytes"] = h2d_bytes
    out["host_pageable_h2d_gbs"] = h2d_bytes / (pq[-1].h2dMs * 1e-3) / 1e9 if pq[-1].h2dMs > 0 else None
    out["pageable_over_device"] = out["host_pageable_ms"] / out["device_resident_ms"]
    out["note"] = ("the host routes are bound by the PCIe copy of the matrix (upper: 8 F bytes) every solve; the "
                   "device-assembly route (b200_assemble_laplacian_device / b200_assemble_p_rgh_device) keeps it in HBM")
    return out


def transport_section(env, ctx, base):
    """the OTHER linear solves of a time step.  The reference solves U, Yi, h and k with
    `smoothSolver` + `symGaussSeidel` (cases/steckler/system/fvSolution:48-61; 261 of the golden log's solves) on
    ASYMMETRIC transport matrices.  Here: a U-shaped system (ddt + convection + diffusion, cases.transport_system) on
    the same mesh through b200_smooth_solve_device, `smoother symGaussSeidel`, to tolerance 1e-6 (the reference's U
    controls without its maxIter 10 cap), default multicolour sweeps.  Reported: time to tolerance, us per
    sweep-iteration (sweeps + residual evaluation), per-kernel fractions on BYTES MOVED (per row 4 B row length + b +
    diag + psi write, per entry value + column, every gathered psi once), and the CPU restatement (one thread: the
    sweep is sequential upstream) on a bounded sample beside it."""
    import numpy as np
    from firefoam_dev_b200 import cases
    from firefoam_dev_b200.ldu import make_smooth_controls
    from oracle import oracle as orc
    torch = env.torch
    t = cases.transport_system(base, seed=31)
    N, F = t.addr.nCells, t.addr.nFaces
    d, up, lo, b = env.up(t.diag), env.up(t.upper), env.up(t.lower), env.up(t.source)
    x = torch.zeros(N, dtype=torch.float64, device=env.dev)
    ctl, _, _ = make_smooth_controls(dict(smoother="symGaussSeidel", tolerance=1e-6, relTol=0.0, maxIter=1000))

    def step():
        x.zero_()
        torch.cuda.current_stream().synchronize()
        return ctx.smooth_solve_device(d, up, lo, [], b, x, ctl)
    step()
    ms, perfs = env.timed(step, 3)
    p = perfs[-1]
    err = float(np.abs(x.cpu().numpy() - t.xstar).max())
    ctx.force_iterations(20)
    step()
    fms, fp = env.timed(step, 1)
    ctx.profile(True)
    step()
    prof = ctx.profile_json()
    ctx.profile(False)
    ctx.force_iterations(0)
    C = p.nColours
    cb = col_bytes_of(ctx)
    rows = N / max(1, C)                      # rows per colour pass (2 equal colours on the hex box)
    pass_bytes = rows * 28 + (2.0 * F / max(1, C)) * (8 + cb) + 8 * (N - rows)
    res_rows = N - rows                       # explicit residual: every row but the last-updated colour
    res_bytes = res_rows * 28 + (2.0 * F * res_rows / N) * (8 + cb) + 8 * rows
    kern = {}
    for name, nbytes in (("gs_sweep_rows", pass_bytes), ("gs_residual", res_bytes)):
        if name in prof:
            us = prof[name]["avg_us"]
            kern[name] = {"avg_us": us, "launches": prof[name]["launches"], "bytes_moved": nbytes,
                          "achieved_gbs": nbytes / (us * 1e-6) / 1e9, "frac": nbytes / (us * 1e-6) / 1e9 / env.peak}
    # the same system through PBiCG + DILU (what the reference's other cases select for these equations;
    # first, unfused version of that path): time to the same tolerance and per iteration
    pbicg = None
    try:
        from firefoam_dev_b200.ldu import make_bicg_controls
        bctl, _ = make_bicg_controls(dict(preconditioner="DILU", tolerance=1e-6, relTol=0.0, maxIter=1000))

        def bstep():
            x.zero_()
            torch.cuda.current_stream().synchronize()
            return ctx.bicg_solve_device(d, up, lo, b, x, bctl)
        bstep()
        bms, bperfs = env.timed(bstep, 2)
        bp = bperfs[-1]
        berr = float(np.abs(x.cpu().numpy() - t.xstar).max())
        ctx.force_iterations(20)
        try:
            bstep()
            _, bf = env.timed(bstep, 1)
        finally:
            ctx.force_iterations(0)
        pbicg = {"solver": "PBiCG + DILU-class (log name DILU(mc)B200PBiCG), tolerance 1e-6", "iterations": bp.nIterations,
                 "converged": bool(bp.converged), "time_to_tolerance_ms": bms / 2, "final_residual": bp.finalResidual,
                 "us_per_iteration": 1e3 * bf[0].solveMs / max(1, bf[0].nIterations), "max_err_vs_xstar": berr}
    except Exception as e:      # (never lose the bench line to the secondary solver)
        pbicg = {"error": str(e)[:300]}
    # CPU restatement on a bounded sample: set-up (Amul, normFactor) + 3 sweep-iterations, one thread
    t0 = time.perf_counter()
    psi = np.zeros(N)
    cp = orc.smooth_solve(t, psi, tolerance=1e-30, maxIter=3)
    cpu_s = time.perf_counter() - t0
    return {"system": "U-shaped asymmetric lduMatrix (ddt + div + laplacian) on the same mesh; smoothSolver + symGaussSeidel, "
                      "tolerance 1e-6, relTol 0; multicolour sweeps (log name B200smoothSolver(mc))",
            "N": N, "F": F, "colours": C, "sweeps_to_tolerance": p.nIterations, "converged": bool(p.converged),
            "time_to_tolerance_ms": ms / 3, "solve_ms": p.solveMs, "setup_ms": p.setupMs,
            "us_per_sweep_iteration": 1e3 * fp[0].solveMs / max(1, fp[0].nIterations),
            "gdof_sweeps_per_s": N * p.nIterations / (ms / 3 * 1e-3) / 1e9,
            "final_residual": p.finalResidual, "max_err_vs_xstar": err, "kernels": kern, "pbicg": pbicg,
            "cpu_port": {"kind": "port", "cores": 1, "sample": "set-up + 3 sweep-iterations of the same system",
                         "seconds": cpu_s, "us_per_sweep_iteration_incl_setup": 1e6 * cpu_s / max(1, cp.nIterations)}}


def strong_base_section(env):
    """The N-GPU weak run solves a mesh of N blocks (N = 8: BASELINE configs[3]'s 128 M mesh): rank 0 alone runs
    the SAME global mesh on one GPU for a fixed 200 iterations, so that the strong-scaling speed-up 1 -> N is
    verifiable from this line."""
    from firefoam_dev_b200 import meshgen as mg
    pkg = env.pkg
    out = None
    dims, _ = hex_layout(env.args, env.n)
    if env.rank == 0:
        t0 = time.time()
        s = mg.hex_block(*dims)
        ctx = pkg.Context(device=env.local)
        try:
            ctx.set_addressing(s.addr)
            res = Resident(env, ctx, s, sctl_for("diagonal"))
            ctx.force_iterations(200)
            res.step()
            e0, e1 = env.torch.cuda.Event(enable_timing=True), env.torch.cuda.Event(enable_timing=True)
            env.torch.cuda.synchronize()
            e0.record()
            p = res.step()
            e1.record()
            env.torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            out = {"mesh": f"hex {dims[0]}x{dims[1]}x{dims[2]} on ONE GPU (rank 0 alone)", "iters": p.nIterations, "ms": ms,
                   "solve_ms": p.solveMs, "us_per_iteration": 1e3 * p.solveMs / max(1, p.nIterations),
                   "gdof_iter_per_s": dims[0] * dims[1] * dims[2] * p.nIterations / (ms * 1e-3) / 1e9,
                   "host_s": time.time() - t0}
        finally:
            ctx.close()
        del s
    env.barrier()
    return out


def run_gpu(args):
    import numpy as np
    env = Env(args)
    torch, dist, pkg, n, rank = env.torch, env.dist, env.pkg, env.n, env.rank
    extras = set(args.extras.split(",")) if args.extras not in ("auto", "none") else set()
    default_run = (args.workload == "hex" and args.scaling == "weak" and args.precond == "diagonal")
    if args.extras == "auto":
        extras = {"mgpu_parity"}
        if default_run:
            extras |= {"dic_class"}
            if n == 1:
                extras |= {"corrector", "transport"}
            if n == 8 and not args.block:
                extras |= {"strong_base", "poly"}

    s, n_global, gen_s = make_system(args, n, rank, pinned=env.pinned, dist=dist)
    if args.workload != "hex":       # generators without a pinned allocator: page-lock what e2e reads
        s.diag, s.upper, s.source = env.to_pinned(s.diag), env.to_pinned(s.upper), env.to_pinned(s.source)
        s.bou = [env.to_pinned(b) for b in s.bou]
    a = s.addr
    N, F = a.nCells, a.nFaces
    ctx = pkg.Context(device=env.local, rank=rank, nranks=n, nccl_uid=env.uid)
    t0 = time.time()
    ctx.set_addressing(a)
    setaddr_s = time.time() - t0
    sctl = sctl_for(args.precond)
    res = Resident(env, ctx, s, sctl)

    for _ in range(args.warmup):
        perf = res.step()
    # ---- timed: device-resident; per-kernel profiling is OFF here (its event records sit between the kernels)
    clocks = ClockSampler(env.local)
    if rank == 0:
        clocks.start()
    l0 = ctx.launch_count()
    ms, perfs = env.timed(res.step, args.steps)
    launches = ctx.launch_count() - l0
    clk = clocks.stop() if rank == 0 else None
    perf = perfs[-1]
    iters = sum(p.nIterations for p in perfs)
    solve_ms = sum(p.solveMs for p in perfs)
    setup_ms = sum(p.setupMs for p in perfs)
    converged = bool(perf.converged)
    err = res.err()
    value = n_global * iters / (ms * 1e-3) / 1e9

    # ---- e2e: the reference-facing plug-in call with pinned host buffers ------------------------
    psi_h = env.pinned(N)
    src_h = s.source
    solver = pkg.B200PCG("p_rgh", s.matrix, s.bou, None, s.interfaces, sctl, context=ctx)
    psi_h[:] = 0.0
    solver.solve(psi_h, src_h)                          # warm-up (allocates staging)
    env.barrier()
    t0 = time.perf_counter()
    e_iters = 0
    h2d_ms = d2h_ms = 0.0
    for _ in range(args.steps):
        psi_h[:] = 0.0
        p = solver.solve(psi_h, src_h)
        e_iters += p.nIterations
        h2d_ms += p.h2dMs
        d2h_ms += p.d2hMs
    env.barrier()
    e_s = env.max_over_ranks(time.perf_counter() - t0)
    e2e_value = n_global * e_iters / e_s / 1e9
    nslots = sum(b.size for b in s.bou)
    h2d_bytes = 8 * (F + 3 * N + nslots)
    d2h_bytes = 8 * N
    e2e_err = float(np.abs(psi_h - s.xstar).max()) if s.xstar is not None else None

    # ---- profiled pass (separate from every timed region): per-kernel table + roofline of the dominant kernel
    prof, pperf = profiled_pass(res)
    dic = args.precond.startswith("DIC")
    kernels = kernel_table(prof, N, F, env.peak, dic, perf.nColours, col_bytes_of(ctx))

    # ---- fixed 200-iteration timing -------------------------------------
    ctx.force_iterations(200)
    res.step()
    fixed_ms, pfs = env.timed(res.step, 1)
    pf = pfs[0]
    ctx.force_iterations(0)
    desc = ctx.describe()
    slots_all = [nslots]
    per_rank = None
    if n > 1:
        slots_all = [None] * n
        dist.all_gather_object(slots_all, nslots)
        # per-rank view of the profiled pass: kernel averages and the in-kernel waits (halo flags, peer reductions)
        mine = {"rank": rank}
        for k_, nm in (("spmv_dot", "amul_us"), ("p_psi_update", "k_p_us"), ("r_update_dots", "k_r_us"), ("iface_fix", "iface_fix_us"),
                       ("_wait_halo_flags", "wait_halo_us"), ("_wait_peer_reduction", "wait_reduction_us")):
            if k_ in prof:
                mine[nm] = round(prof[k_]["avg_us"], 2)
        per_rank = [None] * n
        dist.all_gather_object(per_rank, mine)

    sections = {}
    if "dic_class" in extras:
        sections["dic_class"] = dic_class_section(env, res, n_global)
    if "corrector" in extras and n == 1:
        sections["corrector"] = corrector_section(env, ctx, s)
    if "transport" in extras and n == 1:
        sections["transport"] = transport_section(env, ctx, s)
    if "mgpu_parity" in extras and n > 1:
        sections["mgpu_parity"] = mgpu_parity(env)
    if "strong_base" in extras and n > 1:
        sections["strong_base_1gpu"] = strong_base_section(env)
    if "poly" in extras and n > 1:
        sections["poly"] = poly_section(env, args)

    if rank != 0:
        if n > 1:
            dist.destroy_process_group()
        return 0

    peak, peak_src = env.peak, env.peak_src
    dom = kernels.get("spmv_dot", {})
    eis = "eis_sweeps" in kernels
    if eis:
        dom = kernels["eis_sweeps"]
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "spmv_traffic.json")) as f:
            tj = json.load(f)
            if tj.get("cells") == N and tj.get("kernel", "").split("<")[0] == desc.get("amul_natural", "").split("<")[0] \
                    and not dic:
                traffic = tj.get("dram_bytes_per_launch")
                traffic_src = "static: " + tj.get("source", "ncu --set full capture of this kernel at this size (profiles/)")
    except Exception:
        pass
    iters_per_step = max(1, iters // args.steps)
    iter_us = 1e3 * solve_ms / max(iters, 1)
    # bytes one PCG iteration has to move.  SURVEY.md 8d counts the UNFUSED loop (diagonal 120N+16F, DIC-class
    # 136N+48F); the fused loops move fewer, so both are given: `frac` is on bytes moved (<= 1 by construction),
    # `frac_of_unfused_alg` relates the time to the SURVEY figure and may exceed 1
    unfused = (136 * N + 48 * F) if dic else (120 * N + 16 * F)
    if eis:
        moved = kernels["eis_sweeps"]["alg_bytes"] + 44 * N + 28 * N
    elif dic:
        moved = (24 * N + 24 * F + 4 * N) + (40 * N + 32 * F) + 40 * N + 24 * N
    else:
        moved = (24 * N + 16 * F + 4 * N) + 48 * N + 32 * N
    # compute kernels of one iteration, from the profiled pass (its event pool may cover fewer iterations than the
    # solve has: normalise by the iterations it recorded)
    it_rec = max(1, max(kernels.get("p_psi_update", {}).get("launches", 0), kernels.get("eis_p_psi_update", {}).get("launches", 0)))
    compute_us = sum(kernels[k]["avg_us"] * kernels[k]["launches"] for k in kernels
                     if k in ("spmv_dot", "p_psi_update", "r_update_dots", "dic_fwd", "dic_bwd", "eis_p_psi_update",
                              "eis_bwd", "eis_fwd_dot", "eis_r_update_rho", "eis_true_residual")) / it_rec
    amul_kernel = desc.get("amul_permuted" if dic else "amul_natural", "?")
    dom_name = ("k_eis_bwd + k_eis_fwd per iteration (Eisenstat form: lduMatrix::Amul + DIC-class apply in two sweeps, "
                "fused with gSumProd(wA,pA))" if eis else f"{amul_kernel} (lduMatrix::Amul fused with gSumProd(wA,pA))")
    line = {
        "metric": METRIC.replace("diagonal", args.precond), "value": value, "unit": UNIT, "n_gpus": n, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": dict(config_dict(args, n), colours=perf.nColours, cells_this_rank=N, faces_this_rank=F,
                       halo_slots_this_rank=nslots, halo_faces_per_rank=slots_all,
                       edge_cut_faces=sum(slots_all) // 2),
        "iterations_per_step": iters // args.steps, "converged": converged, "max_err_vs_xstar": err,
        "solve_ms_per_step": solve_ms / args.steps, "setup_ms_per_step": setup_ms / args.steps,
        "time_to_tolerance_ms": (solve_ms + setup_ms) / args.steps,
        "pcg_iteration": {"avg_us": iter_us, "bytes_moved": moved, "gbs": moved / (1e-6 * iter_us) / 1e9,
                          "frac": moved / (1e-6 * iter_us) / 1e9 / peak, "unfused_alg_bytes": unfused,
                          "frac_of_unfused_alg": unfused / (1e-6 * iter_us) / 1e9 / peak,
                          "kernel_sum_us": compute_us,
                          "non_kernel_us": iter_us - compute_us,
                          "note": "avg_us from the UNPROFILED timed region; kernel_sum_us from the profiled pass "
                                  "(compute kernels only; at N > 1 non_kernel_us = exposed halo / reduction / "
                                  "launch time per iteration)"},
        "fixed_200_iterations": {"ms": fixed_ms, "iters": pf.nIterations,
                                 "gdof_iter_per_s": n_global * pf.nIterations / (fixed_ms * 1e-3) / 1e9,
                                 "us_per_iteration": 1e3 * pf.solveMs / max(1, pf.nIterations)},
        "roofline": {"kernel": dom_name,
                     "bound": "hbm", "achieved": dom.get("gbs"), "peak": peak, "unit": "GB/s",
                     "frac": dom.get("frac"), "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                     "alg_bytes_per_launch": dom.get("alg_bytes"), "avg_us": dom.get("avg_us"),
                     "timing": "CUDA events on the library's stream around each launch, in a profiled pass outside the "
                               "timed region; launches of surplus loop bodies (returned on S->done) excluded"},
        "kernels": kernels, "plan": desc,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
                "d2h_bytes_per_step": d2h_bytes, "ms_per_step": 1e3 * e_s / args.steps,
                "h2d_ms_per_step": h2d_ms / args.steps, "d2h_ms_per_step": d2h_ms / args.steps,
                "iterations_per_step": e_iters // args.steps, "max_err_vs_xstar": e2e_err,
                "timing": "host wall clock around blocking B200PCG.solve() calls, max over ranks"},
        "gpu_launches": launches, "clocks": clk,
        "host": {"gen_s": gen_s, "set_addressing_s": setaddr_s},
    }
    if n > 1:
        line["exposed_comm_us"] = iter_us - compute_us
        line["per_rank_profile"] = per_rank
    line.update(sections)
    if sections.get("strong_base_1gpu"):
        sb = sections["strong_base_1gpu"]
        line["strong_scaling_1_to_%d" % n] = {
            "speedup": sb["us_per_iteration"] / (1e3 * pf.solveMs / max(1, pf.nIterations)),
            "basis": "fixed 200 PCG+diagonal iterations of the same global mesh: us per iteration on 1 GPU / on %d GPUs" % n}
    if n == 1 and not args.no_cpu_baseline:
        leg = cpu_leg(args, seconds=args.cpu_seconds)
        line["cpu_baseline"] = {k: leg[k] for k in ("value", "unit", "cores", "kind", "sample")}
    else:
        line["cpu_baseline"] = None
    print(json.dumps(line))
    if n > 1:
        dist.destroy_process_group()
    return 0


def poly_section(env, args):
    """BASELINE configs[4] inside the N = 8 line: the 40 M-cell polyhedral mesh, RCB N-way, PCG + DIC-class and
    PCG + diagonal to tolerance, with the per-kernel roofline fractions."""
    import copy
    pa = copy.copy(args)
    pa.workload, pa.scaling, pa.precond = "poly", "strong", "DIC"
    t0 = time.time()
    s, n_global, gen_s = make_system(pa, env.n, env.rank, pinned=env.pinned, dist=env.dist)
    ctx = env.pkg.Context(device=env.local, rank=env.rank, nranks=env.n, nccl_uid=env.new_uid())
    out = {"config": config_dict(pa, env.n)}
    try:
        ctx.set_addressing(s.addr)
        for pre in ("DIC", "diagonal"):
            res = Resident(env, ctx, s, sctl_for(pre))
            res.step()
            ms, perfs = env.timed(res.step, 2)
            iters = sum(p.nIterations for p in perfs)
            prof, perf = profiled_pass(res)
            kern = kernel_table(prof, res.N, res.F, env.peak, pre == "DIC", perf.nColours, col_bytes_of(ctx))
            out[pre] = {"value": n_global * iters / (ms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": ms / 2,
                        "iterations_per_step": iters // 2, "converged": bool(perf.converged), "colours": perf.nColours,
                        "us_per_iteration": 1e3 * sum(p.solveMs for p in perfs) / max(1, iters),
                        "max_err_vs_xstar": res.err(),
                        "kernels": {k: v for k, v in kern.items() if "frac" in v or k in ("iface_fix", "eis_bwd", "eis_fwd_dot")}}
            del res
        out["plan"] = ctx.describe()
        out["host_s"] = time.time() - t0
    finally:
        ctx.close()
    return out


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
    ap.add_argument("--precond", default=PRECOND,
                    choices=["none", "diagonal", "DIC", "DIC-exact", "DIC-eisenstat", "DIC-multicolour"])
    ap.add_argument("--extras", default="auto",
                    help="extra sections of the JSON line: auto | none | comma list of dic_class,corrector,transport,"
                         "mgpu_parity,strong_base,poly.  auto: mgpu_parity at N > 1; on the default workload also dic_class "
                         "(configs[3]'s preconditioner), corrector + transport (smoothSolver, SURVEY 8f-4) at N = 1, "
                         "strong_base + poly (configs[4]) at N = 8")
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
