#!/bin/bash
cd "$(dirname "$0")/.."
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu.log
run() { label="$1"; shift; echo "=== $label"; env "$@" python tools/quick_perf.py 256 250 250 diagonal 100 noconv 2>&1 | grep -E "spmv_dot|rep2"; }
{
run "tma exact (3,3)"  B200PCG_SPMV=tma
run "tma generic (4,4)" B200PCG_SPMV=tma B200PCG_EXACT=0
run "tma exact ctas=5" B200PCG_SPMV=tma B200PCG_CTAS=5
run "tma exact ctas=6" B200PCG_SPMV=tma B200PCG_CTAS=6
} > gpurun_out/sweep2.log 2>&1
cat gpurun_out/sweep2.log
python bench.py --workload poly --poly 125 125 160 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/bench_poly5m_diag_rcm.json 2>gpurun_out/bench_poly.err
python bench.py --workload poly --poly 125 125 160 --precond DIC --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/bench_poly5m_dic_rcm.json 2>>gpurun_out/bench_poly.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 2 --warmup 2 > gpurun_out/bench_2gpu.json 2>gpurun_out/bench_2gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 2 --warmup 2 --precond DIC > gpurun_out/bench_2gpu_dic.json 2>>gpurun_out/bench_2gpu.err
tail -3 gpurun_out/bench_2gpu.err gpurun_out/bench_poly.err
echo done
