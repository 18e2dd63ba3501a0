#!/bin/bash
# round 2, call Y (8 GPUs): the bench's mgpu_parity section (PCG paths + smooth_solver) at 8 ranks on a small block --
# the code path the driver's scaling run executes at N = 8
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29731 bench.py --gpus 8 --steps 1 --warmup 3 --block 64 64 64 --extras mgpu_parity --no-cpu-baseline > gpurun_out/r2y_bench_8gpu_small_block.json 2> gpurun_out/r2y_bench_8gpu_small_block.err; echo "bench exit $?"; tail -3 gpurun_out/r2y_bench_8gpu_small_block.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2y_bench_8gpu_small_block.json").read().strip().splitlines()[-1])
m = d["mgpu_parity"]
print("pcg parity pass", m.get("pass"), "smooth", json.dumps(m.get("smooth_solver")))
PY
echo done
