#!/bin/bash
cd "$(dirname "$0")/.."
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke.log
run() { label="$1"; shift; echo "=== $label"; env "$@" timeout 200 python tools/quick_perf.py 256 250 250 diagonal 100 noconv 2>&1 | grep -E "spmv_dot|rep2"; }
{
run "tma u8 lens, default (4 CTAs/SM)"
run "tma ctas=3" B200PCG_CTAS=3
run "tma ctas=5" B200PCG_CTAS=5
run "tma ctas=6" B200PCG_CTAS=6
} > gpurun_out/sweep3.log 2>&1
cat gpurun_out/sweep3.log
timeout 600 python bench.py > gpurun_out/bench_default_1gpu.json 2> gpurun_out/bench_default_1gpu.err; echo "bench exit $?"
timeout 300 python bench.py --precond DIC --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/bench_1gpu_dic.json 2>/dev/null; echo "bench dic exit $?"
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_reference_arm.json 2>/dev/null; echo "ref exit $?"
echo done
