#!/bin/bash
# round 2, call B (1 GPU): full GPU suite after the defaults changed, new bench.py (small block first), graph A/B
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/r2b_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/r2b_pytest_gpu.log
timeout 300 python bench.py --block 64 50 50 --steps 2 --warmup 1 --cpu-seconds 1 > gpurun_out/r2b_bench_small.json 2> gpurun_out/r2b_bench_small.err; echo "bench small exit $?"; tail -3 gpurun_out/r2b_bench_small.err
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2b_bench_1gpu.json 2> gpurun_out/r2b_bench_1gpu.err; echo "bench exit $?"; tail -3 gpurun_out/r2b_bench_1gpu.err
cut -c1-400 gpurun_out/r2b_bench_1gpu.json
B200PCG_GRAPH=0 timeout 300 python bench.py --steps 5 --warmup 3 --extras none --no-cpu-baseline > gpurun_out/r2b_bench_1gpu_nograph.json 2>> gpurun_out/r2b_bench_1gpu.err; echo "nograph exit $?"
cut -c1-400 gpurun_out/r2b_bench_1gpu_nograph.json
echo done
