#!/bin/bash
# the whole GPU suite
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -q -m gpu --timeout=900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" > gpurun_out/summary.txt
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu.log | tail -25; cat gpurun_out/summary.txt
