#!/bin/bash
# round 2, call Z (1 GPU): first GPU run of the PBiCG + DILU path
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_bicg.py -q --tb=short > gpurun_out/r2z_pytest_bicg.log 2>&1; echo "pytest exit $?"; tail -40 gpurun_out/r2z_pytest_bicg.log
echo done
