#!/bin/bash
# correlation stage check: parity tests of the pyramid / lookup + end-to-end, kernel timings, bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_corr.py tests/test_gpu_slim_e2e.py -q -m gpu --timeout=300 > gpurun_out/pytest_new.log 2>&1; echo "pytest exit $?" > gpurun_out/summary.txt
timeout 300 python tools/kbench.py --skip-pillar > gpurun_out/kbench.txt 2>&1; echo "kbench exit $?" >> gpurun_out/summary.txt
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.txt
grep -E "^(FAILED|ERROR)|passed|failed|Error|assert" gpurun_out/pytest_new.log | head -30
cat gpurun_out/kbench.txt; cat gpurun_out/summary.txt; tail -n 5 gpurun_out/bench.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench.json'))
    print('value',d['value'],'e2e',d['e2e']['value'],'ms/step',d['ms_per_step'],'other',d.get('other_mode'))
    for k in d['kernels'][:12]: print(k['kernel'], round(k['avg_ms'],4), k['launches_per_step'], round(k.get('frac',0),3), k.get('tensor',{}).get('frac'))
except Exception as e: print('bench parse failed', e)
PY
