#!/bin/bash
# round 2, call Z2 (1 GPU): final code with the PBiCG path: full GPU suite, smoke(), PBiCG timings on the 16 M hex box
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r2z2_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/r2z2_pytest_gpu.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/r2z2_smoke.log 2>&1; echo "smoke exit $?"; tail -4 gpurun_out/r2z2_smoke.log
timeout 200 python tools/smooth_perf.py 256 250 250 bicg iters=20 > gpurun_out/r2z2_perf_bicg_hex16m.log 2>&1; echo "perf exit $?"; tail -1 gpurun_out/r2z2_perf_bicg_hex16m.log
echo done
