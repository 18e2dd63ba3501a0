#!/bin/bash
# round 2, call C (2 GPUs): new N > 1 bench sections on reduced sizes (logic check), then the real 2-GPU line
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
N=${1:-2}
run() { tag="$1"; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 \
        --master-port 29577 bench.py --gpus $N "$@" > gpurun_out/r2c_bench_${N}gpu_$tag.json 2> gpurun_out/r2c_bench_${N}gpu_$tag.err; \
        echo "$tag exit $?"; tail -2 gpurun_out/r2c_bench_${N}gpu_$tag.err; cut -c1-300 gpurun_out/r2c_bench_${N}gpu_$tag.json; }
run small --block 64 50 50 --poly 40 40 50 --steps 2 --warmup 1 --extras mgpu_parity,dic_class,strong_base,poly
run full --steps 3 --warmup 3
timeout 300 python -m pytest tests/test_multigpu.py -x -q > gpurun_out/r2c_pytest_mgpu_$N.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2c_pytest_mgpu_$N.log
echo done
