#!/bin/bash
# round 2, call W (1 GPU): final code of the session: full GPU suite, smoke(), smoothSolver timings (two red-black sweeps
# per counted symGaussSeidel sweep on two-colour plans), default bench line with every N = 1 section, reference arm,
# ncu --set full of the final sweep kernels
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2w_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/r2w_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2w_smoke.log 2>&1; echo "smoke exit $?"; tail -4 gpurun_out/r2w_smoke.log
timeout 300 python tools/smooth_perf.py 256 250 250 iters=20 > gpurun_out/r2w_perf_hex16m.log 2>&1; echo "perf hex exit $?"
timeout 300 python tools/smooth_perf.py 125 125 160 poly iters=20 > gpurun_out/r2w_perf_poly5m.log 2>&1; echo "perf poly exit $?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2w_perf_*.log")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        for sm in ("symGaussSeidel", "GaussSeidel"):
            k = d[sm]["profiled"]["kernels"]
            print(f.split("r2w_perf_")[1], sm, "us/iter", round(d[sm]["timed"]["us_per_iter"], 1), "to tol", d[sm]["to_tolerance"]["iters"], round(d[sm]["to_tolerance"]["solve_ms"], 2), "ms",
                  {n: (round(v["avg_us"], 1), v["launches"]) for n, v in k.items()})
    except Exception as e:
        print(f, "unreadable", e)
PY
( time timeout 900 python bench.py ) > gpurun_out/r2w_bench_1gpu.json 2> gpurun_out/r2w_bench_1gpu.err; echo "bench exit $?"; tail -4 gpurun_out/r2w_bench_1gpu.err
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r2w_bench_reference.json 2> gpurun_out/r2w_bench_reference.err; echo "reference exit $?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2w_bench_1gpu.json").read().strip().splitlines()[-1])
print("value", round(d["value"], 2), "e2e", round(d["e2e"]["value"], 2), "iter_us", round(d["pcg_iteration"]["avg_us"], 1), "amul frac", d["roofline"]["frac"], "launches", d.get("gpu_launches"), d["clocks"])
print("dic", d["dic_class"]["value"], d["dic_class"]["time_to_tolerance_ms"]); print("corrector", {k: v for k, v in d["corrector"].items() if k.endswith("_ms") or k == "iterations"})
print("transport", json.dumps(d["transport"])[:1400])
r = json.loads(open("gpurun_out/r2w_bench_reference.json").read().strip().splitlines()[-1]); print("reference", r["value"], r["cpu_baseline"]["sample"][:120])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_gs_rows" -s 6 -c 6 -o gpurun_out/r2w_prof_gs_hex \
    python tools/smooth_perf.py 256 250 250 iters=8 > gpurun_out/r2w_ncu_gs_hex.log 2>&1; echo "ncu exit $?"
ls -la gpurun_out/r2w*.ncu-rep | tail -2
echo done
