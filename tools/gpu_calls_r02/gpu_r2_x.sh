#!/bin/bash
# round 2, call X (2 GPUs): final code: multi-GPU parity (PCG paths with peer-memory / NCCL halos, smoothSolver) and a
# short driver-shaped bench line with the mgpu_parity section (incl. its smooth_solver entry)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multigpu.py -q --tb=short -k "2-None or 2-nccl-halo or 2-no-overlap" > gpurun_out/r2x_pytest_mgpu2.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/r2x_pytest_mgpu2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29721 bench.py --gpus 2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2x_bench_2gpu.json 2> gpurun_out/r2x_bench_2gpu.err; echo "bench exit $?"; tail -3 gpurun_out/r2x_bench_2gpu.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2x_bench_2gpu.json").read().strip().splitlines()[-1])
m = d["mgpu_parity"]
print("value", round(d["value"], 2), "iter_us", round(d["pcg_iteration"]["avg_us"], 1), "dic", round(d["dic_class"]["value"], 2))
print("pcg parity pass", m.get("pass"), "smooth", json.dumps(m.get("smooth_solver")))
PY
echo done
