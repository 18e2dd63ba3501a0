#!/bin/bash
# round 2, call F (8 GPUs): multi-GPU parity at 8 and 4 ranks, then the driver-shaped N = 8 line with every section
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
N=${1:-8}
timeout 600 python -m pytest tests/test_multigpu.py -q -k "8- or 4-None" > gpurun_out/r2f_pytest_mgpu_$N.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2f_pytest_mgpu_$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29577 \
    bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2f_bench_${N}gpu.json 2> gpurun_out/r2f_bench_${N}gpu.err; echo "bench exit $?"
tail -3 gpurun_out/r2f_bench_${N}gpu.err | cut -c1-300
python - gpurun_out/r2f_bench_${N}gpu.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value", round(d["value"],2), "iter_us", round(d["pcg_iteration"]["avg_us"],1), "kernel_sum", round(d["pcg_iteration"]["kernel_sum_us"],1), d["plan"].get("halo_exchange"))
print("dic", d.get("dic_class",{}).get("value"), d.get("dic_class",{}).get("us_per_iteration"), "parity", d.get("mgpu_parity",{}).get("pass"))
print("strong", d.get("strong_scaling_1_to_8"), (d.get("strong_base_1gpu") or {}).get("us_per_iteration"))
p=d.get("poly",{}); print("poly", {k:(round(p[k]["value"],2), p[k]["iterations_per_step"], round(p[k]["us_per_iteration"],1)) for k in ("DIC","diagonal") if k in p}, p.get("host_s"))
PY
echo done
