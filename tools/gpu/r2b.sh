#!/bin/bash
# lookup kernels: parity, timings, one ncu --set full capture
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_corr.py -q -m gpu --timeout=300 -x > gpurun_out/test_gpu_corr.log 2>&1; echo "corr tests exit $?" > gpurun_out/summary.txt
tail -n 5 gpurun_out/test_gpu_corr.log
timeout 600 python tools/kbench.py --skip-pillar > gpurun_out/kbench.txt 2>&1; echo "kbench exit $?" >> gpurun_out/summary.txt
grep "lookup\|fused" gpurun_out/kbench.txt | grep -v "gen0\|gen1 nhwc"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_lookup_conv_tmem|k_lookup_conv_tf32' -s 1 -c 2 -f -o gpurun_out/prof_lookup python tools/kprof_lookup.py > gpurun_out/ncu_lookup.log 2>&1; echo "ncu exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
