#!/bin/bash
# round-1 final: 1-GPU regression of the whole GPU suite (full-size / multi-GPU files excluded: run separately),
# A/B of the Eisenstat sweep builds, bench line of PCG + DIC-class in the Eisenstat form
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -x -q --deselect tests/test_full_size.py --deselect tests/test_multigpu.py > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_gpu_final.log
run() { label="$1"; shift; echo "=== $label"; env "$@" timeout 100 python tools/quick_perf.py 256 250 250 DIC-eisenstat 100 2>&1 | grep -E "eis_|rep2|tolerance"; }
{
run "default (6-entry batches, 4 CTAs/SM build)"
run "3 CTAs/SM build (80 registers)" B200PCG_EIS_CTAS=3
run "32-bit columns" B200PCG_COL16=0
run "32-bit columns, 3 CTAs/SM build" B200PCG_COL16=0 B200PCG_EIS_CTAS=3
} > gpurun_out/eis_sweep3.log 2>&1
cat gpurun_out/eis_sweep3.log
timeout 150 python bench.py --precond DIC-eisenstat --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_1gpu_dic_eisenstat.json 2> gpurun_out/bench_eis.err; echo "bench exit $?"; cut -c1-600 gpurun_out/bench_1gpu_dic_eisenstat.json; tail -3 gpurun_out/bench_eis.err
echo done
