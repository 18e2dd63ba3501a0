#!/bin/bash
# round 2, call M (4 GPUs): does the slow rank follow the GPU or the sub-mesh?  same run with the rank -> device map reversed
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
N=${1:-4}
run() { tag="$1"; shift; env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 \
        --master-port 29577 bench.py --gpus $N --steps 3 --warmup 3 --extras none > gpurun_out/r2m_bench_${N}gpu_$tag.json 2> gpurun_out/r2m_bench_${N}gpu_$tag.err; \
        echo "$tag exit $?"; tail -2 gpurun_out/r2m_bench_${N}gpu_$tag.err | cut -c1-300; \
        python - gpurun_out/r2m_bench_${N}gpu_$tag.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("  value", round(d["value"],2), "iter_us", round(d["pcg_iteration"]["avg_us"],1))
    for r in d.get("per_rank_profile") or []: print("   ", r)
except Exception as e: print("  parse error", e)
PY
}
run straight B200PCG_X=0
run reversed B200_BENCH_REVERSE_DEVICES=1
timeout 300 python -m pytest tests/test_multigpu.py -q -k "2-" > gpurun_out/r2m_pytest_mgpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2m_pytest_mgpu.log
echo done
