#!/bin/bash
# first GPU shake-out: every test file under its own timeout so a hung kernel cannot eat the lease
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for t in test_gpu_pillar test_gpu_corr test_gpu_slim_e2e; do
  timeout 600 python -m pytest tests/$t.py -q -m gpu -x --timeout=300 -s > gpurun_out/$t.log 2>&1
  echo "$t exit $?" >> gpurun_out/summary.txt
done
tail -5 gpurun_out/test_gpu_pillar.log gpurun_out/test_gpu_corr.log gpurun_out/test_gpu_slim_e2e.log
cat gpurun_out/summary.txt
