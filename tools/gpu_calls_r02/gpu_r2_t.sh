#!/bin/bash
# round 2, call T (2 GPUs): smoothSolver with processor patches against the 2-rank oracle (peer-memory and NCCL halos),
# and the existing 2-GPU PCG parity once (nothing else may have moved)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_multigpu.py -q --tb=short -k "smooth and 2-" > gpurun_out/r2t_pytest_mgpu_smooth.log 2>&1; echo "smooth exit $?"; tail -30 gpurun_out/r2t_pytest_mgpu_smooth.log
timeout 400 python -m pytest tests/test_multigpu.py -q --tb=short -k "test_multigpu_parity and 2-None" > gpurun_out/r2t_pytest_mgpu_pcg.log 2>&1; echo "pcg exit $?"; tail -5 gpurun_out/r2t_pytest_mgpu_pcg.log
echo done
