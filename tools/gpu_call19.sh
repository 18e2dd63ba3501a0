#!/bin/bash
# Eisenstat form after the byte savers: parity tests, timing, sweep grid sizes, ncu --set full of one iteration
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_eisenstat.py -x -q > gpurun_out/pytest_eis.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_eis.log
run() { label="$1"; shift; echo "=== $label"; env "$@" timeout 200 python tools/quick_perf.py 256 250 250 DIC-eisenstat 100 2>&1 | grep -E "eis_|rep2|tolerance"; }
{
run "default (8 CTAs/SM sweeps)"
run "4 CTAs/SM" B200PCG_SWEEP_CTAS=4
run "16 CTAs/SM" B200PCG_SWEEP_CTAS=16
run "col32" B200PCG_COL16=0
} > gpurun_out/eis_sweep.log 2>&1
cat gpurun_out/eis_sweep.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_eis" -s 73 -c 5 -o gpurun_out/prof_eis_hex python tools/quick_perf.py 256 250 250 DIC-eisenstat 12 noconv > gpurun_out/ncu_eis.log 2>&1; echo "ncu exit $?"
ls -la gpurun_out/*.ncu-rep; tail -3 gpurun_out/ncu_eis.log
echo done
