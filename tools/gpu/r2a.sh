#!/bin/bash
# round 2, first GPU pass: the new lookup kernels (parity, then timings), then the whole GPU suite
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/smi.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_corr.py -q -m gpu --timeout=300 -x > gpurun_out/test_gpu_corr.log 2>&1; echo "corr tests exit $?" > gpurun_out/summary.txt
tail -n 25 gpurun_out/test_gpu_corr.log
timeout 600 python tools/kbench.py --skip-pillar > gpurun_out/kbench.txt 2>&1; echo "kbench exit $?" >> gpurun_out/summary.txt
grep -v "^$" gpurun_out/kbench.txt | tail -40
timeout 1200 python -m pytest tests -q -m gpu --timeout=600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/summary.txt
tail -n 8 gpurun_out/pytest_gpu.log; cat gpurun_out/summary.txt
