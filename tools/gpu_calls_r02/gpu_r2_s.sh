#!/bin/bash
# round 2, call S (1 GPU): smoothSolver after the HALO / one-wave refactor: parity tests, register-build A/B
# (B200PCG_GS_CTAS=3|4) and grid A/B on the 16 M hex box and 5 M polyhedra, bench transport section
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_smooth.py -q --tb=short > gpurun_out/r2s_pytest_smooth.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/r2s_pytest_smooth.log
for ct in 3 4; do
  B200PCG_GS_CTAS=$ct timeout 300 python tools/smooth_perf.py 256 250 250 iters=20 > gpurun_out/r2s_perf_hex16m_ct$ct.log 2>&1; echo "perf hex ct=$ct exit $?"
  B200PCG_GS_CTAS=$ct B200PCG_SWEEP_CTAS=8 timeout 300 python tools/smooth_perf.py 256 250 250 iters=20 > gpurun_out/r2s_perf_hex16m_ct${ct}_grid8.log 2>&1; echo "perf hex ct=$ct grid8 exit $?"
done
timeout 300 python tools/smooth_perf.py 125 125 160 poly iters=20 > gpurun_out/r2s_perf_poly5m.log 2>&1; echo "perf poly exit $?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2s_perf_*.log")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        for sm in ("symGaussSeidel", "GaussSeidel"):
            k = d[sm]["profiled"]["kernels"]
            print(f.split("r2s_perf_")[1], sm, "us/iter", round(d[sm]["timed"]["us_per_iter"], 1), "to tol", d[sm]["to_tolerance"]["iters"], round(d[sm]["to_tolerance"]["solve_ms"], 2), "ms",
                  {n: round(v["avg_us"], 1) for n, v in k.items()})
    except Exception as e:
        print(f, "unreadable", e)
PY
timeout 900 python bench.py --steps 2 --warmup 3 --extras transport > gpurun_out/r2s_bench_transport.json 2> gpurun_out/r2s_bench_transport.err; echo "bench exit $?"; tail -3 gpurun_out/r2s_bench_transport.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2s_bench_transport.json").read().strip().splitlines()[-1])
print("value", round(d["value"], 2), "transport:", json.dumps(d.get("transport"))[:1500])
PY
echo done
