#!/bin/bash
# round 2, call L (8 GPUs): split vs single interface update at 8 ranks, then the N = 8 line (all sections but poly) with the winner
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
N=${1:-8}
run() { tag="$1"; shift; env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 \
        --master-port 29577 bench.py --gpus $N --steps $STEPS --warmup 3 $EXTRA > gpurun_out/r2l_bench_${N}gpu_$tag.json 2> gpurun_out/r2l_bench_${N}gpu_$tag.err; \
        echo "$tag exit $?"; tail -2 gpurun_out/r2l_bench_${N}gpu_$tag.err | cut -c1-300; \
        python - gpurun_out/r2l_bench_${N}gpu_$tag.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("  value", round(d["value"],2), "iter_us", round(d["pcg_iteration"]["avg_us"],1), "kernel_sum", round(d["pcg_iteration"]["kernel_sum_us"],1), d["plan"].get("halo_exchange"), "dic", d.get("dic_class",{}).get("value"), d.get("dic_class",{}).get("us_per_iteration"), "parity", d.get("mgpu_parity",{}).get("pass"), "strong", d.get("strong_scaling_1_to_8"))
    for r in d.get("per_rank_profile") or []: print("   ", r)
except Exception as e: print("  parse error", e)
PY
}
STEPS=3; EXTRA="--extras none"
run nosplit B200PCG_SPLIT_IFACE=0
run split B200PCG_SPLIT_IFACE=1
BEST=$(python - <<'PY'
import json
v={}
for t in ("nosplit","split"):
    try: v[t]=json.loads(open(f"gpurun_out/r2l_bench_8gpu_{t}.json").read().strip().splitlines()[-1])["pcg_iteration"]["avg_us"]
    except Exception: v[t]=1e9
print(1 if v["split"] < 0.99*v["nosplit"] else 0)
PY
)
echo "winner: B200PCG_SPLIT_IFACE=$BEST"
STEPS=5; EXTRA="--extras dic_class,mgpu_parity,strong_base"
run final B200PCG_SPLIT_IFACE=$BEST
echo done
