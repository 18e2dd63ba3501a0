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
