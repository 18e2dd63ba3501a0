#!/bin/bash
# round 2, call E (1 GPU): full suite with the SR layout / natural negSumDiag, polyhedral A/B, ncu evidence
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/r2e_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/r2e_pytest_gpu.log
for v in sr ell; do
  env=""; [ $v = ell ] && env="B200PCG_SPMV=ell"
  env $env timeout 300 python bench.py --workload poly --poly 125 125 160 --precond diagonal --steps 2 --warmup 2 --extras none --no-cpu-baseline \
      > gpurun_out/r2e_bench_poly5m_diag_$v.json 2> gpurun_out/r2e_bench_poly5m_diag_$v.err; echo "poly diag $v exit $?"
done
timeout 300 python bench.py --workload poly --poly 125 125 160 --precond DIC --steps 2 --warmup 2 --extras none --no-cpu-baseline \
      > gpurun_out/r2e_bench_poly5m_dic.json 2> gpurun_out/r2e_bench_poly5m_dic.err; echo "poly dic exit $?"
python - <<'PY'
import json
for f in ("diag_sr","diag_ell","dic"):
    try:
        d=json.loads(open(f"gpurun_out/r2e_bench_poly5m_{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"],2), "iter_us", round(d["pcg_iteration"]["avg_us"],1), {k:(v["launches"],round(v["avg_us"],1),round(v.get("frac",0),3)) for k,v in d["kernels"].items() if v["launches"]>1 or k.startswith("asm")}, d["plan"]["amul_natural"], "setup_ms", d["setup_ms_per_step"])
    except Exception as e: print(f, "ERR", e)
PY
# ncu: launch list of the default bench (shares), then --set full of the dominant kernels
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/r2e_launches.csv \
    python bench.py --steps 1 --warmup 1 --extras none --no-cpu-baseline > gpurun_out/r2e_ncu_launches.log 2>&1; echo "ncu launches exit $?"
ncu --set full --clock-control none --import-source on -k regex:"k_spmv_sym_tma|k_p|k_r" -s 90 -c 6 -o gpurun_out/r2e_prof_hex_diag \
    python tools/quick_perf.py 256 250 250 diagonal 40 noconv > gpurun_out/r2e_ncu_hex_diag.log 2>&1; echo "ncu hex diag exit $?"
ncu --set full --clock-control none --import-source on -k regex:"k_eis_bwd|k_eis_fwd|k_eis_p|k_eis_r" -s 80 -c 8 -o gpurun_out/r2e_prof_hex_eis \
    python tools/quick_perf.py 256 250 250 DIC 30 noconv > gpurun_out/r2e_ncu_hex_eis.log 2>&1; echo "ncu hex eis exit $?"
ncu --set full --clock-control none --import-source on -k regex:"k_spmv_sr|k_neg_sum|k_fill" -s 2 -c 8 -o gpurun_out/r2e_prof_poly_sr \
    python tools/poly_perf.py diagonal 12 > gpurun_out/r2e_ncu_poly_sr.log 2>&1; echo "ncu poly exit $?"
ls -la gpurun_out/*.ncu-rep
echo done
