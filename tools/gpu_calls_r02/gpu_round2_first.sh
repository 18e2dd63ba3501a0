#!/bin/bash
# First GPU calls of round 2: validate what round 1 left CPU-checked only, then measure it.
#   usage (1 GPU):  gpurun --timeout 600 -- 'bash tools/gpu_round2_first.sh 1'
#         (N GPUs): gpurun --gpus N --timeout 600 -- 'bash tools/gpu_round2_first.sh N'
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
N=${1:-1}
if [ "$N" = "1" ]; then
  # whole 1-GPU suite (incl. full-size), then the DIC-class A/B on hex and polyhedra
  B200_TEST_UNVALIDATED=1 timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/r2_pytest_gpu.log
  for pre in DIC DIC-eisenstat; do
    timeout 100 python tools/quick_perf.py 256 250 250 $pre 100 2>&1 | grep -E "eis_|dic_|spmv_dot|rep2|tolerance" > gpurun_out/r2_perf_hex_$pre.log
    timeout 300 python bench.py --workload poly --poly 125 125 160 --precond $pre --steps 2 --warmup 3 --no-cpu-baseline \
        > gpurun_out/r2_bench_poly5m_$pre.json 2>> gpurun_out/r2_bench.err; echo "poly $pre exit $?"
  done
  B200PCG_SORT_COLS=1 timeout 300 python bench.py --workload poly --poly 125 125 160 --precond DIC-eisenstat --steps 2 --warmup 3 \
      --no-cpu-baseline > gpurun_out/r2_bench_poly5m_DIC-eisenstat_sortcols.json 2>> gpurun_out/r2_bench.err; echo "poly sortcols exit $?"
  cat gpurun_out/r2_perf_hex_*.log
  # pageable caller memory: plain vs staged copies (h2d= / d2h= columns of quick_perf)
  for st in 0 1; do B200PCG_STAGED_COPY=$st timeout 100 python tools/quick_perf.py 256 250 250 diagonal 50 noconv 2>&1 | grep rep2; done
else
  # multi-GPU parity incl. the overlapped Eisenstat halo sequence, then config 4 (128 M hex at 8) both ways
  B200_TEST_UNVALIDATED=1 timeout 500 python -m pytest tests/test_multigpu.py -x -q > gpurun_out/r2_pytest_mgpu_$N.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/r2_pytest_mgpu_$N.log
  run() { tag="$1"; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 \
          --master-port 29577 bench.py --gpus $N --steps 2 --warmup 3 --precond DIC-eisenstat > gpurun_out/r2_bench_${N}gpu_$tag.json 2>> gpurun_out/r2_bench.err; \
          echo "$tag exit $?"; cut -c1-220 gpurun_out/r2_bench_${N}gpu_$tag.json; }
  run eis_plain
  run eis_overlap B200PCG_EIS_OVERLAP=1
fi
echo done
