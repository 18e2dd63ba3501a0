#!/bin/bash
# round 2, call R (1 GPU): smoothSolver with the redundant passes skipped and the last group's residual fused:
# parity tests, smoke(), per-sweep timing, bench transport section alone, ncu --set full of the sweep kernels
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_smooth.py -q --tb=short > gpurun_out/r2r_pytest_smooth.log 2>&1; echo "pytest exit $?"; tail -30 gpurun_out/r2r_pytest_smooth.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2r_smoke.log 2>&1; echo "smoke exit $?"; tail -4 gpurun_out/r2r_smoke.log
timeout 400 python tools/smooth_perf.py 256 250 250 iters=20 > gpurun_out/r2r_perf_hex16m.log 2>&1; echo "perf hex exit $?"; tail -1 gpurun_out/r2r_perf_hex16m.log
timeout 300 python tools/smooth_perf.py 125 125 160 poly iters=20 > gpurun_out/r2r_perf_poly5m.log 2>&1; echo "perf poly exit $?"; tail -1 gpurun_out/r2r_perf_poly5m.log
timeout 300 python tools/smooth_perf.py 30 15 20 iters=20 > gpurun_out/r2r_perf_steckler_size.log 2>&1; echo "perf small exit $?"; tail -1 gpurun_out/r2r_perf_steckler_size.log
timeout 300 python tools/smooth_perf.py 30 15 20 exact iters=20 > gpurun_out/r2r_perf_steckler_size_exact.log 2>&1; echo "perf small exact exit $?"; tail -1 gpurun_out/r2r_perf_steckler_size_exact.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_gs_rows|k_gs_resid" -s 4 -c 6 -o gpurun_out/r2r_prof_gs_hex \
    python tools/smooth_perf.py 256 250 250 iters=6 > gpurun_out/r2r_ncu_gs_hex.log 2>&1; echo "ncu exit $?"
ls -la gpurun_out/*.ncu-rep | tail -2
echo done
