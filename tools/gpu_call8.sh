#!/bin/bash
# 8-GPU measurement call: multi-GPU parity tests, hex weak-scaling (diag + DIC-class), 40 M polyhedral (DIC-class + diag)
cd "$(dirname "$0")/.."
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_multigpu.py -x -q > gpurun_out/pytest_mgpu.log 2>&1; echo "pytest mgpu exit $?"; tail -4 gpurun_out/pytest_mgpu.log
timeout 300 $TR --master-port 29701 bench.py --gpus 8 --workload poly --poly 40 40 40 --precond DIC --steps 1 --warmup 1 > gpurun_out/bench_poly_small_8gpu.json 2> gpurun_out/bench_poly_small_8gpu.err; echo "poly small exit $?"; tail -2 gpurun_out/bench_poly_small_8gpu.err
timeout 600 $TR --master-port 29702 bench.py --gpus 8 --steps 2 --warmup 2 > gpurun_out/bench_8gpu_diag.json 2> gpurun_out/bench_8gpu_diag.err; echo "hex diag exit $?"
timeout 600 $TR --master-port 29703 bench.py --gpus 8 --steps 2 --warmup 2 --precond DIC > gpurun_out/bench_8gpu_dic.json 2> gpurun_out/bench_8gpu_dic.err; echo "hex dic exit $?"
timeout 900 $TR --master-port 29704 bench.py --gpus 8 --workload poly --precond DIC --steps 2 --warmup 2 --poly-cache /dev/shm > gpurun_out/bench_poly40m_8gpu_dic.json 2> gpurun_out/bench_poly40m_8gpu_dic.err; echo "poly dic exit $?"; tail -2 gpurun_out/bench_poly40m_8gpu_dic.err
timeout 600 $TR --master-port 29705 bench.py --gpus 8 --workload poly --precond diagonal --steps 2 --warmup 2 --poly-cache /dev/shm > gpurun_out/bench_poly40m_8gpu_diag.json 2> gpurun_out/bench_poly40m_8gpu_diag.err; echo "poly diag exit $?"
echo done
