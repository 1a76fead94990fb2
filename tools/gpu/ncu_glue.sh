#!/bin/bash
# ncu --set full of the InstanceNorm and update-block glue kernels inside one eager bench step
mkdir -p gpurun_out
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'k_in_stats|k_in_apply|k_gru_gate|k_nhwc_pack|k_iter_update|k_add_relu' -c 60 -f -o gpurun_out/prof_glue python bench.py --profile-one-step --no-cuda-graph --warmup 3 > gpurun_out/ncu_glue.log 2>&1; echo "ncu glue exit $?"
tail -2 gpurun_out/ncu_glue.log
