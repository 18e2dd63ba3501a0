#!/bin/bash
cd "$(dirname "$0")/.."
timeout 1200 python -m pytest tests -m gpu -x -q --deselect tests/test_full_size.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/pytest_gpu.log
run() { label="$1"; shift; echo "=== $label"; env "$@" timeout 200 python tools/quick_perf.py 256 250 250 DIC 100 noconv 2>&1 | grep -E "spmv_dot|dic_|p_psi|r_update|rep2"; }
{
run "DIC col16 (default)"
run "DIC col32" B200PCG_COL16=0
} > gpurun_out/col16.log 2>&1
cat gpurun_out/col16.log
timeout 300 python bench.py --workload poly --poly 125 125 160 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/bench_poly5m_diag_c16.json 2>gpurun_out/bench_poly.err; echo "poly diag exit $?"
timeout 300 python bench.py --workload poly --poly 125 125 160 --precond DIC --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/bench_poly5m_dic_c16.json 2>>gpurun_out/bench_poly.err; echo "poly dic exit $?"
echo done
