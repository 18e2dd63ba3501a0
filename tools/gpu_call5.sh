#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/pytest_gpu.log
{
for dims in "40 40 40" "64 50 40" "64 64 64" "100 80 64"; do
  for small in 0 1000000; do
    for ctas in 8 16; do
      [ "$small" = "0" ] && [ "$ctas" = "16" ] && continue
      echo "=== dims $dims SMALL_N=$small CTAS=$ctas"
      B200PCG_SMALL_N=$small B200PCG_SMALL_CTAS=$ctas timeout 120 python tools/quick_perf.py $dims diagonal 200 noconv 2>&1 | grep -E "rep2"
    done
  done
done
} > gpurun_out/small_crossover.log 2>&1
cat gpurun_out/small_crossover.log
timeout 900 python bench.py --scaling strong --gpus 1 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_strong_1gpu.json 2> gpurun_out/bench_strong_1gpu.err; echo "strong exit $?"; tail -3 gpurun_out/bench_strong_1gpu.err
echo done
