#!/bin/bash
# scaled Eisenstat form: parity tests, batched vs plain sweeps, ncu of one iteration
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_eisenstat.py -x -q > gpurun_out/pytest_eis.log 2>&1; echo "pytest exit $?"; tail -12 gpurun_out/pytest_eis.log
run() { label="$1"; shift; echo "=== $label"; env "$@" timeout 200 python tools/quick_perf.py 256 250 250 DIC-eisenstat 100 2>&1 | grep -E "eis_|rep2|tolerance|dic_calc"; }
{
run "default (batched sweeps, one wave)"
run "plain loops" B200PCG_EIS_BATCH=0
run "batched, 8 CTAs/SM grid" B200PCG_SWEEP_CTAS=8
} > gpurun_out/eis_sweep2.log 2>&1
cat gpurun_out/eis_sweep2.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"k_eis_(p|bwd|fwd|r)" -s 56 -c 4 -o gpurun_out/prof_eis2_hex python tools/quick_perf.py 256 250 250 DIC-eisenstat 12 noconv > gpurun_out/ncu_eis2.log 2>&1; echo "ncu exit $?"
echo done
