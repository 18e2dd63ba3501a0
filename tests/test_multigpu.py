"""Multi-GPU parity (needs >= 2 GPUs: run with `gpurun --gpus 2`): NCCL send/recv halos overlapped
with the interior Amul, NCCL all-reduce of the CG scalars, against the N-rank CPU oracle (pthreads,
rank-ascending sums).  Rank-local DIC (block-Jacobi, like upstream) makes DIC iteration counts
depend on the decomposition -- the N-rank oracle has the same decomposition."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def ngpus():
    try:
        from firefoam_dev_b200 import _lib
        return _lib.load_pcg().b200_device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("world,tile", [(2, None), (2, "64"), (2, "no-overlap"), (2, "nccl-halo"), (2, "fuse"), (2, "split"), (4, None),
                                        (4, "no-overlap"), (8, None), (8, "128"), (8, "no-overlap"), (8, "nccl-halo"),
                                        (8, "fuse")])
def test_multigpu_parity(world, tile):
    """tile: B200PCG_TILE for the ranks (tiled multicolour order + symmetric Amul in the DIC-class mode;
    the default tile of 8192 rows does not engage on these small sub-meshes); "no-overlap": the Eisenstat
    form with the halo exchange exposed between the sweeps (B200PCG_EIS_OVERLAP=0; the default overlaps it with
    the first colour's backward sweep)."""
    if ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29500 + world),
           os.path.join(ROOT, "tests", "mgpu_worker.py")]
    env = dict(os.environ)
    if tile == "nccl-halo":      # processor-patch halos over ncclSend/ncclRecv instead of peer-memory stores
        env["B200PCG_HALO"] = "nccl"
        tile = None
    elif tile == "split":        # opt-in: k_iface_pre (comm stream, concurrent with the Amul) + k_iface_apply
        env["B200PCG_SPLIT_IFACE"] = "1"
        tile = None
    elif tile == "fuse":         # opt-in: pack fused into k_p's tail, interface fix-up into the Amul's tail
        env["B200PCG_FUSE_IFACE"] = "1"
        tile = None
    elif tile == "no-overlap":
        env["B200PCG_EIS_OVERLAP"] = "0"
        tile = None
    elif tile:
        env["B200PCG_TILE"] = tile
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("MGPU_RESULT ")][-1]
    res = json.loads(line[len("MGPU_RESULT "):])
    assert res["amul_bit_exact"]
    # a rank-local argument error in the collective b200_set_addressing fails on EVERY rank instead of hanging
    assert res["bad_rank_rejected_everywhere"] and "another rank" in res["bad_rank_message_rank0"]
    for key in ("diagonal", "DIC-exact"):
        assert res[key]["iters"] == res[key]["oracle_iters"], (key, res[key])
        assert res[key]["relerr_vs_oracle"] < 1e-11, (key, res[key])
        assert res[key]["init"] == pytest.approx(res[key]["oracle_init"], rel=1e-12)
    assert res["none"]["converged"] and res["none"]["relerr_vs_oracle"] < 1e-5
    assert res["DIC"]["converged"] and res["DIC"]["relerr_vs_oracle"] < 1e-6
    assert res["DIC"]["iters"] < res["diagonal"]["iters"]
    for key in ("poly-diagonal", "poly-DIC-exact"):
        assert res[key]["iters"] == res[key]["oracle_iters"], (key, res[key])
        assert res[key]["relerr_vs_oracle"] < 1e-11, (key, res[key])
    assert res["poly-DIC"]["converged"] and res["poly-DIC"]["relerr_vs_oracle"] < 1e-6
    if res["eisenstat_ran"]:
        # Eisenstat form (halo term inside the forward sweep): same iterates as the three-kernel DIC-class loop
        for key, base in (("DIC-eisenstat", "DIC"), ("poly-DIC-eisenstat", "poly-DIC")):
            assert res[key]["converged"] and res[key]["relerr_vs_oracle"] < 1e-6, (key, res[key])
            assert res[base]["iters"] <= res[key]["iters"] <= res[base]["iters"] + 2, (key, res[key], res[base])


@pytest.mark.parametrize("world,halo", [(2, None), (2, "nccl-halo"), (4, None), (8, None)])
def test_multigpu_smooth_parity(world, halo):
    """smoothSolver with processor patches against the N-rank CPU oracle: asymmetric Amul
    bit-identical; level-scheduled sweeps bit-identical psi and identical sweep counts (peer-memory halos when one
    sweep + the residual lie between two reductions, ncclSend/ncclRecv for nSweeps > 1 and fixed sweeps); the
    multicolour sweeps reach the same solution."""
    if ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29600 + world),
           os.path.join(ROOT, "tests", "mgpu_smooth_worker.py")]
    env = dict(os.environ)
    if halo == "nccl-halo":
        env["B200PCG_HALO"] = "nccl"
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("MGPU_SMOOTH_RESULT ")][-1]
    res = json.loads(line[len("MGPU_SMOOTH_RESULT "):])
    for tag in ("hex-", "poly-"):
        assert res[tag + "amul_bit_exact"]
        for key in ("exact-fixed3", "exact-sym", "exact-gs-nsweeps2", "exact-U-controls"):
            e = res[tag + key]
            assert e["iters"] == e["oracle_iters"] and e["bit_identical"], (tag + key, e)
            if key != "exact-fixed3":
                assert e["init"] == pytest.approx(e["oracle_init"], rel=1e-12)
                assert e["final"] == pytest.approx(e["oracle_final"], rel=1e-9)
        for key in ("mc-sym", "mc-gs-nsweeps3"):
            e = res[tag + key]
            assert e["converged"] and e["relerr_vs_oracle"] < 1e-8, (tag + key, e)
