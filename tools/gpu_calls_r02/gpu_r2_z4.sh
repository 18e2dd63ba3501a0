#!/bin/bash
# round 2, call Z4 (1 GPU): the bench's transport section with its PBiCG entry
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 200 python bench.py --steps 1 --warmup 3 --extras transport --no-cpu-baseline > gpurun_out/r2z4_bench_transport.json 2> gpurun_out/r2z4_bench_transport.err; echo "bench exit $?"; tail -2 gpurun_out/r2z4_bench_transport.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2z4_bench_transport.json").read().strip().splitlines()[-1])
print("value", round(d["value"], 2), "transport.pbicg:", json.dumps(d["transport"].get("pbicg")), "sweeps", d["transport"]["sweeps_to_tolerance"], d["transport"]["time_to_tolerance_ms"])
PY
echo done
