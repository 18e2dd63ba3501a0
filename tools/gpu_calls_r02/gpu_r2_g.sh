#!/bin/bash
# round 2, call G (4 GPUs): where does the multi-GPU overhead sit?  per-rank kernel times + in-kernel wait accounting
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
N=${1:-4}
run() { tag="$1"; shift; env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 \
        --master-port 29577 bench.py --gpus $N --steps 3 --warmup 3 --extras none > gpurun_out/r2g_bench_${N}gpu_$tag.json 2> gpurun_out/r2g_bench_${N}gpu_$tag.err; \
        echo "$tag exit $?"; tail -2 gpurun_out/r2g_bench_${N}gpu_$tag.err | cut -c1-300; \
        python - gpurun_out/r2g_bench_${N}gpu_$tag.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("  value", round(d["value"],2), "iter_us", round(d["pcg_iteration"]["avg_us"],1), "kernel_sum", round(d["pcg_iteration"]["kernel_sum_us"],1), d["plan"].get("halo_exchange"), d["clocks"])
    for r in d.get("per_rank_profile") or []: print("   ", r)
except Exception as e: print("  parse error", e)
PY
}
run p2p B200PCG_X=0
run nccl B200PCG_HALO=nccl
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,power.limit,temperature.gpu --format=csv
echo done
