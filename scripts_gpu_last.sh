#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_glue.py tests/test_gpu_slim_e2e.py -q -m gpu --timeout=200 -x > gpurun_out/pytest_new.log 2>&1; echo "pytest exit $?" > gpurun_out/summary.txt
bash scripts_gpu_ncu_corr.sh >> gpurun_out/summary.txt 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.txt
grep -E "^(FAILED|ERROR)|passed|failed|Error|assert" gpurun_out/pytest_new.log | head -20
cat gpurun_out/summary.txt; tail -n 3 gpurun_out/bench.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench.json'))
    print('value',d['value'],'e2e',d['e2e']['value'],'ms/step',d['ms_per_step'],'other',d.get('other_mode'))
    for k in d['kernels'][:8]: print(k['kernel'], round(k['avg_ms'],4), k['launches_per_step'], round(k.get('frac',0),3))
except Exception as e: print('bench parse failed', e)
PY
