#!/bin/bash
cd "$(dirname "$0")/.."
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/pytest_gpu.log
run() { label="$1"; shift; echo "=== $label"; env "$@" timeout 200 python tools/quick_perf.py 256 250 250 DIC 100 noconv 2>&1 | grep -E "spmv_dot|dic_|p_psi|r_update|rep2"; }
{
run "DIC default (tile 8192, sym, fused first colour)"
run "DIC tile 2048" B200PCG_TILE=2048
run "DIC tile 32768" B200PCG_TILE=32768
run "DIC untiled ELL (previous)" B200PCG_TILE=0 B200PCG_FUSE_FIRST=0
run "DIC untiled, fused first" B200PCG_TILE=0
} > gpurun_out/dic_tiles.log 2>&1
cat gpurun_out/dic_tiles.log
timeout 300 python bench.py --workload poly --poly 125 125 160 --precond DIC --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/bench_poly5m_dic_tiled.json 2>gpurun_out/bench_poly.err; echo "poly exit $?"
echo done
