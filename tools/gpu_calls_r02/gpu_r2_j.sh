#!/bin/bash
# round 2, call J (1 GPU): final code: full GPU suite, smoke(), default bench, reference arm (short), ncu of the Eisenstat sweeps
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -q > gpurun_out/r2j_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/r2j_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2j_smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/r2j_smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2j_bench_1gpu.json 2> gpurun_out/r2j_bench_1gpu.err; echo "bench exit $?"; tail -3 gpurun_out/r2j_bench_1gpu.err
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r2j_bench_reference.json 2> gpurun_out/r2j_bench_reference.err; echo "reference exit $?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2j_bench_1gpu.json").read().strip().splitlines()[-1])
print("value", round(d["value"],2), "iter_us", round(d["pcg_iteration"]["avg_us"],1), "kernel_sum", round(d["pcg_iteration"]["kernel_sum_us"],1), "amul", d["roofline"]["frac"], d["clocks"])
print("dic", d["dic_class"]["value"], d["dic_class"]["time_to_tolerance_ms"]); print("corrector", {k:v for k,v in d["corrector"].items() if k.endswith("_ms") or k=="iterations"})
r=json.loads(open("gpurun_out/r2j_bench_reference.json").read().strip().splitlines()[-1]); print("reference", r["value"], r["cpu_baseline"]["sample"][:160])
PY
ncu --set full --clock-control none --import-source on -k regex:"k_eis_bwd|k_eis_fwd|k_eis_p|k_eis_r" -s 80 -c 8 -o gpurun_out/r2j_prof_hex_eis \
    python tools/quick_perf.py 256 250 250 DIC 30 noconv > gpurun_out/r2j_ncu_hex_eis.log 2>&1; echo "ncu hex eis exit $?"
ls -la gpurun_out/*.ncu-rep | tail -3
echo done
