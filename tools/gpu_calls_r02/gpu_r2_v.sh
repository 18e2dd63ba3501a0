#!/bin/bash
# round 2, call V (1 GPU): smoothSolver, two-colour plans with the residual completed by the next iteration's first
# pass (no residual kernel): parity tests, A/B against the explicit residual kernel, bench transport section
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_smooth.py -q --tb=short > gpurun_out/r2v_pytest_smooth.log 2>&1; echo "pytest exit $?"; tail -25 gpurun_out/r2v_pytest_smooth.log
timeout 300 python tools/smooth_perf.py 256 250 250 iters=20 > gpurun_out/r2v_perf_hex16m_lagged.log 2>&1; echo "perf lagged exit $?"
B200PCG_GS_LAGGED=0 timeout 300 python tools/smooth_perf.py 256 250 250 iters=20 > gpurun_out/r2v_perf_hex16m_explicit.log 2>&1; echo "perf explicit exit $?"
timeout 300 python tools/smooth_perf.py 30 15 20 iters=20 > gpurun_out/r2v_perf_steckler_size.log 2>&1; echo "perf small exit $?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2v_perf_*.log")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        for sm in ("symGaussSeidel", "GaussSeidel"):
            k = d[sm]["profiled"]["kernels"]
            print(f.split("r2v_perf_")[1], sm, "us/iter", round(d[sm]["timed"]["us_per_iter"], 1), "to tol", d[sm]["to_tolerance"]["iters"], round(d[sm]["to_tolerance"]["solve_ms"], 2), "ms",
                  {n: (round(v["avg_us"], 1), v["launches"]) for n, v in k.items()})
    except Exception as e:
        print(f, "unreadable", e)
PY
timeout 900 python bench.py --steps 2 --warmup 3 --extras transport --no-cpu-baseline > gpurun_out/r2v_bench_transport.json 2> gpurun_out/r2v_bench_transport.err; echo "bench exit $?"; tail -3 gpurun_out/r2v_bench_transport.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2v_bench_transport.json").read().strip().splitlines()[-1])
print("value", round(d["value"], 2), "transport:", json.dumps(d.get("transport"))[:1600])
PY
echo done
