#!/bin/bash
# round 2, call O (2 GPUs): k_eis_fwd_rows merged into k_eis_halo; separate profile class for the interface-row kernels
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m pytest tests/test_multigpu.py -x -q > gpurun_out/r2o_pytest_mgpu_$N.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2o_pytest_mgpu_$N.log
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29577 \
    bench.py --gpus $N --steps 3 --warmup 3 --extras dic_class,mgpu_parity > gpurun_out/r2o_bench_${N}gpu.json 2> gpurun_out/r2o_bench_${N}gpu.err; echo "bench exit $?"
python - gpurun_out/r2o_bench_${N}gpu.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value", round(d["value"],2), "iter_us", round(d["pcg_iteration"]["avg_us"],1), "parity", d["mgpu_parity"]["pass"])
dc=d["dic_class"]; print("dic", round(dc["value"],2), round(dc["us_per_iteration"],1), {k:(v["launches"], round(v["avg_us"],1), round(v.get("frac",0),3)) for k,v in dc["kernels"].items()})
PY
echo done
