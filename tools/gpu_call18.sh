#!/bin/bash
# Eisenstat form of the DIC-class loop: GPU parity tests, then A/B against the three-kernel loop (16 M hex)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_eisenstat.py -x -q > gpurun_out/pytest_eis.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/pytest_eis.log
for pre in DIC DIC-eisenstat; do
  echo "=== $pre"
  timeout 200 python tools/quick_perf.py 256 250 250 $pre 100 > gpurun_out/perf_$pre.log 2>&1; echo "exit $?"
  grep -E "rep2|us|tolerance" gpurun_out/perf_$pre.log
done
echo done
