#!/bin/bash
# round 2, call H (2 GPUs): fused pack (k_p tail) + fused interface fix-up (Amul tail) + release/acquire flags: parity, then A/B
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
N=${1:-2}
timeout 500 python -m pytest tests/test_multigpu.py -x -q > gpurun_out/r2i_pytest_mgpu_$N.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2i_pytest_mgpu_$N.log
run() { tag="$1"; shift; env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 \
        --master-port 29577 bench.py --gpus $N --steps 3 --warmup 3 $EXTRA > gpurun_out/r2i_bench_${N}gpu_$tag.json 2> gpurun_out/r2i_bench_${N}gpu_$tag.err; \
        echo "$tag exit $?"; tail -2 gpurun_out/r2i_bench_${N}gpu_$tag.err | cut -c1-300; \
        python - gpurun_out/r2i_bench_${N}gpu_$tag.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("  value", round(d["value"],2), "iter_us", round(d["pcg_iteration"]["avg_us"],1), "kernel_sum", round(d["pcg_iteration"]["kernel_sum_us"],1), d["plan"].get("halo_exchange"), "dic", d.get("dic_class",{}).get("value"), d.get("dic_class",{}).get("us_per_iteration"), "parity", d.get("mgpu_parity",{}).get("pass"))
    for r in d.get("per_rank_profile") or []: print("   ", r)
except Exception as e: print("  parse error", e)
PY
}
EXTRA="--extras dic_class,mgpu_parity"
run fused B200PCG_X=0
EXTRA="--extras none"
run nofuse B200PCG_FUSE_IFACE=0
run fused_graph B200PCG_GRAPH_MULTI=1
echo done
