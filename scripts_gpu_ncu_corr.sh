#!/bin/bash
# ncu --set full of the correlation-stage kernels and the update-block glue inside one eager bench step
mkdir -p gpurun_out
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'k_corr_gemm|k_corr_lookup|k_feat_pack|k_gru_gate|k_nhwc_pack|k_iter_update' -c 30 -f -o gpurun_out/prof_corr python bench.py --profile-one-step --no-cuda-graph --warmup 3 > gpurun_out/ncu_corr.log 2>&1; echo "ncu corr exit $?"
tail -2 gpurun_out/ncu_corr.log
