#!/bin/bash
# ncu --set full of the hand-written kernels driven by tools/kbench.py (args: kernel regex, extra kbench flags)
mkdir -p gpurun_out
REGEX=${1:-k_corr_gemm|k_corr_lookup|k_tile_encode}
shift
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$REGEX" -s 6 -c 4 -f -o gpurun_out/profk python tools/kbench.py --reps 2 "$@" > gpurun_out/ncu_k.log 2>&1; echo "ncu exit $?"
tail -5 gpurun_out/ncu_k.log
