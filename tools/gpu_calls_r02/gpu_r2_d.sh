#!/bin/bash
# round 2, call D (2 GPUs): peer-memory halo exchange: parity first, then A/B against the NCCL halo and against no graphs
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
N=${1:-2}
timeout 400 python -m pytest tests/test_multigpu.py -x -q > gpurun_out/r2d_pytest_mgpu_$N.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2d_pytest_mgpu_$N.log
run() { tag="$1"; shift; env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 \
        --master-port 29577 bench.py --gpus $N --steps 3 --warmup 3 $EXTRA > gpurun_out/r2d_bench_${N}gpu_$tag.json 2> gpurun_out/r2d_bench_${N}gpu_$tag.err; \
        echo "$tag exit $?"; tail -2 gpurun_out/r2d_bench_${N}gpu_$tag.err | cut -c1-300; \
        python - gpurun_out/r2d_bench_${N}gpu_$tag.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("  value", round(d["value"],2), "iter_us", round(d["pcg_iteration"]["avg_us"],1), "kernel_sum", round(d["pcg_iteration"]["kernel_sum_us"],1), "halo", d["plan"].get("halo_exchange"), "dic", d.get("dic_class",{}).get("value"), d.get("dic_class",{}).get("us_per_iteration"))
except Exception as e: print("  parse error", e)
PY
}
EXTRA="--extras dic_class,mgpu_parity"
run p2p B200PCG_X=0
run nccl B200PCG_HALO=nccl
EXTRA="--extras none"
run p2p_nograph B200PCG_GRAPH=0
run nccl_nograph B200PCG_HALO=nccl B200PCG_GRAPH=0
echo done
