#!/bin/bash
cd "$(dirname "$0")/.."
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29703 bench.py --gpus 8 --steps 2 --warmup 2 --precond DIC > gpurun_out/bench_8gpu_dic_v7.json 2> gpurun_out/bench_8gpu_dic_v7.err; echo "hex dic exit $?"
timeout 400 $TR --master-port 29702 bench.py --gpus 8 --steps 2 --warmup 2 > gpurun_out/bench_8gpu_diag_v7.json 2> gpurun_out/bench_8gpu_diag_v7.err; echo "hex diag exit $?"
