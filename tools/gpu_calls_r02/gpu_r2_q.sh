#!/bin/bash
# round 2, call Q (1 GPU): first GPU run of the smoothSolver path: parity tests, a memcheck of one small
# solve, per-sweep timing on the 16 M hex box and 5 M polyhedra
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_smooth.py -q --tb=short > gpurun_out/r2q_pytest_smooth.log 2>&1; echo "pytest exit $?"; tail -40 gpurun_out/r2q_pytest_smooth.log
timeout 400 python tools/smooth_perf.py 256 250 250 iters=20 > gpurun_out/r2q_perf_hex16m.log 2>&1; echo "perf hex exit $?"; tail -2 gpurun_out/r2q_perf_hex16m.log
timeout 300 python tools/smooth_perf.py 125 125 160 poly iters=20 > gpurun_out/r2q_perf_poly5m.log 2>&1; echo "perf poly exit $?"; tail -2 gpurun_out/r2q_perf_poly5m.log
echo done
