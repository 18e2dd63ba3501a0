#!/bin/bash
# round 2, call P (1 GPU): Eisenstat sweeps on 5 M polyhedra: batched vs plain entry loops, sweep grid sizes
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
run() { tag="$1"; shift; env "$@" timeout 300 python bench.py --workload poly --poly 125 125 160 --precond DIC --steps 2 --warmup 2 --extras none --no-cpu-baseline \
      > gpurun_out/r2p_poly5m_dic_$tag.json 2> gpurun_out/r2p_poly5m_dic_$tag.err; echo "$tag exit $?"; \
      python - gpurun_out/r2p_poly5m_dic_$tag.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("  value", round(d["value"],2), "iter_us", round(d["pcg_iteration"]["avg_us"],1), "iters", d["iterations_per_step"], {k:(v["launches"],round(v["avg_us"],1)) for k,v in d["kernels"].items() if k.startswith("eis_")})
PY
}
run default B200PCG_X=0
run plainloop B200PCG_EIS_BATCH=0
run plainloop_16 B200PCG_EIS_BATCH=0 B200PCG_SWEEP_CTAS=16
run batch_ctas2 B200PCG_SWEEP_CTAS=2
run sortcols B200PCG_SORT_COLS=1 B200PCG_EIS_BATCH=0
echo done
