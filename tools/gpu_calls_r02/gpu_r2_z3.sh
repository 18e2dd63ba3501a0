#!/bin/bash
# round 2, call Z3 (1 GPU): final library (dump format with PBiCG solves): full GPU suite
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q > gpurun_out/r2z3_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/r2z3_pytest_gpu.log
echo done
