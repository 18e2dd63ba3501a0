#!/bin/bash
# Amul variant sweep on the 16 M hex box (1 GPU): CUDA-event averages per kernel class.
# usage (on the GPU box): bash tools/gpu_sweep_spmv.sh > gpurun_out/sweep.log
cd "$(dirname "$0")/.."
run() {  # label, env...
  label="$1"; shift
  echo "=== $label"
  env "$@" python tools/quick_perf.py 256 250 250 diagonal 100 noconv 2>&1 | grep -E "spmv_dot|p_psi|r_update|rep2"
}
run "tma (previous kernel)"        B200PCG_SPMV=tma
run "win run=8 next=1 ctas=4"      B200PCG_RUN=8 B200PCG_NEXT=1 B200PCG_CTAS=4
run "win run=8 next=0 ctas=4"      B200PCG_RUN=8 B200PCG_NEXT=0 B200PCG_CTAS=4
run "win run=8 next=1 ctas=3"      B200PCG_RUN=8 B200PCG_NEXT=1 B200PCG_CTAS=3
run "win run=8 next=1 ctas=5"      B200PCG_RUN=8 B200PCG_NEXT=1 B200PCG_CTAS=5
run "win run=4 next=1 ctas=4"      B200PCG_RUN=4 B200PCG_NEXT=1 B200PCG_CTAS=4
run "win run=16 next=1 ctas=4"     B200PCG_RUN=16 B200PCG_NEXT=1 B200PCG_CTAS=4
run "win run=32 next=1 ctas=4"     B200PCG_RUN=32 B200PCG_NEXT=1 B200PCG_CTAS=4
run "win run=2 next=1 ctas=4"      B200PCG_RUN=2 B200PCG_NEXT=1 B200PCG_CTAS=4
run "win run=16 next=1 ctas=5"     B200PCG_RUN=16 B200PCG_NEXT=1 B200PCG_CTAS=5
