#!/bin/bash
# whole GPU suite, then the round-2 measurement pass (bench, reference arm, ncu launch list, ncu --set full of the hot kernels)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout=900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" > gpurun_out/summary_tests.txt
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu.log | tail -25
EXPORT_PAIRS=${EXPORT_PAIRS:-10000} bash tools/gpu/round2.sh
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
timeout 300 python tools/kbench.py --skip-pillar --skip-corr > gpurun_out/kbench_deflate.txt 2>&1; tail -8 gpurun_out/kbench_deflate.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_deflate -c 4 -f -o gpurun_out/prof_deflate python tools/kbench.py --skip-pillar --skip-corr --reps 1 > gpurun_out/ncu_deflate.log 2>&1; echo "ncu deflate exit $?"
cat gpurun_out/summary_tests.txt
