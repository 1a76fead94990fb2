#!/bin/bash
# ncu --set full of the hand-written kernels driven by tools/kbench.py (args: kernel regex, extra kbench flags)
mkdir -p gpurun_out
REGEX=${1:-k_corr_gemm|k_corr_lookup|k_pillar_nhwc}
shift
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$REGEX" --launch-skip-before-match 0 -c 40 -f -o gpurun_out/profk python tools/kbench.py --reps 1 "$@" > gpurun_out/ncu_k.log 2>&1; echo "ncu exit $?"
tail -3 gpurun_out/ncu_k.log
