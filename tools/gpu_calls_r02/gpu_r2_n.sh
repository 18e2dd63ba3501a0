#!/bin/bash
# round 2, call N (1 GPU): last check of the final library: full GPU suite + smoke + short default bench
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -q > gpurun_out/r2n_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/r2n_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2n_smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/r2n_smoke.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2n_bench_1gpu.json 2> gpurun_out/r2n_bench_1gpu.err; echo "bench exit $?"; tail -3 gpurun_out/r2n_bench_1gpu.err
cut -c1-300 gpurun_out/r2n_bench_1gpu.json
echo done
