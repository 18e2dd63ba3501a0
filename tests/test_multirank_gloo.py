"""N > 1 host logic on the CPU: 2 processes over gloo run PCG+diagonal with numpy kernels driven by
the plan structures the CUDA kernels use, and must agree with the 2-rank oracle."""
import json
import os
import subprocess
import sys

import numpy as np

from conftest import ROOT
from firefoam_dev_b200 import meshgen as mg
from oracle import oracle as orc


def test_two_rank_gloo_matches_oracle():
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29631",
           os.path.join(ROOT, "tests", "gloo_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    res = json.loads([l for l in r.stdout.splitlines() if l.startswith("GLOO_RESULT ")][-1][12:])
    subs = [mg.hex_block(8, 6, 4, 2, 1, 1, k) for k in range(2)]
    ref = [np.zeros(s.addr.nCells) for s in subs]
    p = orc.pcg_solve(subs, ref, "diagonal", 1e-8, 0.0, 1000)
    assert res["iters"] == p.nIterations
    assert abs(res["init"] - p.initialResidual) <= 1e-12 * p.initialResidual
    assert res["nbr"] == [[1], [0]]
    for k in range(2):
        np.testing.assert_allclose(np.array(res["psi"][k]), ref[k], rtol=1e-11, atol=1e-14)


import pytest  # noqa: E402


@pytest.mark.parametrize("world", [2, 4])
def test_multi_rank_eisenstat_gloo(world):
    """Eisenstat form of the DIC-class loop on 2 and 4 ranks (numpy kernels, gloo): plain and overlapped halo
    sequence against the three-kernel multicolour loop on the same ranks and the N-rank oracle.  4 ranks:
    several neighbours per rank, cells with processor faces towards different ranks."""
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29633 + world),
           os.path.join(ROOT, "tests", "gloo_eis_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    res = json.loads([l for l in r.stdout.splitlines() if l.startswith("GLOO_EIS_RESULT ")][-1][16:])
    poly = mg.bcc_poly(4, 3, 3, shuffle_block=64)
    procs = {2: (2, 1, 1), 4: (2, 2, 1)}[world]
    cases = {"hex": [mg.hex_block(8, 6, 4, *procs, k) for k in range(world)],
             "poly": mg.decompose(poly, mg.partition_rcb(poly.xyz, world), world)}
    for name, subs in cases.items():
        d = res[name]
        nr, ne, no = d["iters"]
        assert nr <= ne <= nr + 2 and nr <= no <= nr + 2, d["iters"]
        assert max(d["final"]) < 1e-9
        assert all(c >= 2 for c in d["colours"]) and all(n > 0 for n in d["ifaceRows"])
        ref = [np.zeros(s.addr.nCells) for s in subs]
        orc.pcg_solve(subs, ref, "DIC", 1e-11, 0.0, 1000)
        for k in range(world):
            xr, xe, xo = (np.array(d[key][k]) for key in ("ref", "eis", "ovl"))
            assert np.linalg.norm(xe - xr) / np.linalg.norm(xr) < 1e-8
            # the overlapped sequence performs the same row operations; only the order in which the
            # shares of (p^, w^) are summed differs
            assert np.linalg.norm(xo - xe) / np.linalg.norm(xe) < 1e-10
            assert np.linalg.norm(xe - ref[k]) / np.linalg.norm(ref[k]) < 1e-6
    assert sum(res["poly"]["multiFaceRows"]) > 0      # cells with several processor faces are exercised


def test_reference_arm_under_torchrun_only_rank0_prints():
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29632", os.path.join(ROOT, "bench.py"),
           "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0",
           "--block", "16", "12", "10", "--cpu-seconds", "0.2"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0


@pytest.mark.parametrize("world", [2, 4])
def test_multi_rank_smooth_solver_gloo(world):
    """smoothSolver with processor patches on 2 and 4 ranks (numpy kernels, gloo) against the N-rank oracle
    (oracle/smooth_oracle.c): the level-scheduled sequence is bit-identical (psi, sweep counts), also with nSweeps 2 and
    with fixed sweeps; the multicolour sequence reaches the same solution.  4 ranks: several neighbours per rank."""
    from firefoam_dev_b200 import cases as cs
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29643 + world),
           os.path.join(ROOT, "tests", "gloo_smooth_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    res = json.loads([l for l in r.stdout.splitlines() if l.startswith("GLOO_SMOOTH_RESULT ")][-1][19:])
    NX, NY, NZ = 8, 6, 4
    PX, PY, PZ = {2: (2, 1, 1), 4: (2, 2, 1)}[world]
    g = cs.transport_system(mg.hex_block(NX, NY, NZ), seed=21)
    c = np.arange(NX * NY * NZ)
    c2p = ((c % NX) * PX // NX) + PX * ((((c // NX) % NY) * PY // NY) + PY * ((c // (NX * NY)) * PZ // NZ))
    poly = mg.bcc_poly(4, 3, 3, shuffle_block=64)
    pt = cs.transport_system(poly, seed=22, kappa=0.3)
    systems = {"hex": mg.decompose(g, c2p.astype(np.int32), world),
               "poly": mg.decompose(pt, mg.partition_rcb(poly.xyz, world), world)}
    for name, subs in systems.items():
        d = res[name]
        for key, o in (("exact", dict(tolerance=1e-8, maxIter=500)),
                       ("exact_gs2", dict(tolerance=1e-8, maxIter=500, nSweeps=2, smoother="GaussSeidel")),
                       ("exact_fixed3", dict(nSweeps=-3))):
            ref = [np.zeros(s.addr.nCells) for s in subs]
            p = orc.smooth_solve(subs, ref, **o)
            assert d[key]["n"] == p.nIterations, (name, key)
            for k in range(world):
                assert np.array_equal(np.array(d[key]["psi"][k]), ref[k]), (name, key, k)
            if key != "exact_fixed3":
                assert d[key]["init"] == pytest.approx(p.initialResidual, rel=1e-12)
                assert d[key]["final"] == pytest.approx(p.finalResidual, rel=1e-9)
        ref = [np.zeros(s.addr.nCells) for s in subs]
        p = orc.smooth_solve(subs, ref, tolerance=1e-13, maxIter=5000)
        assert d["mc"]["final"] < 1e-11
        for k in range(world):
            assert np.abs(np.array(d["mc"]["psi"][k]) - ref[k]).max() <= 1e-8 * max(np.abs(r_).max() for r_ in ref)
    assert sum(res["poly"]["exact"]["multiFaceRows"]) > 0      # cells with several processor faces are exercised
