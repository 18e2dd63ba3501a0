#!/bin/bash
# round 2, call U (4 GPUs): smoothSolver with processor patches at 4 ranks (several neighbours per rank), and the bench's
# mgpu_parity section with its new smooth_solver entry on a small block
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_multigpu.py -q --tb=short -k "smooth and 4-" > gpurun_out/r2u_pytest_mgpu4_smooth.log 2>&1; echo "smooth exit $?"; tail -20 gpurun_out/r2u_pytest_mgpu4_smooth.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 4 --steps 1 --warmup 3 --block 96 96 96 --extras mgpu_parity --no-cpu-baseline > gpurun_out/r2u_bench_4gpu_small_block.json 2> gpurun_out/r2u_bench_4gpu_small_block.err; echo "bench exit $?"; tail -3 gpurun_out/r2u_bench_4gpu_small_block.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2u_bench_4gpu_small_block.json").read().strip().splitlines()[-1])
m = d["mgpu_parity"]
print("pcg pass", m.get("pass"), "smooth", json.dumps(m.get("smooth_solver")))
PY
echo done
