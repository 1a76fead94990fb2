#!/bin/bash
# whole GPU suite, then the round-2 measurement pass (bench, reference arm, ncu launch list, ncu --set full of the hot kernels)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout=900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" > gpurun_out/summary_tests.txt
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu.log | tail -25
bash tools/gpu/round2.sh
cat gpurun_out/summary_tests.txt
