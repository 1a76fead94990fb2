#!/bin/bash
# short check: the given test files (default: decoder + glue), then a bench line without the CPU arm
mkdir -p gpurun_out
TESTS=${@:-tests/test_gpu_decoder.py tests/test_gpu_glue.py}
timeout 600 python -m pytest $TESTS -q -m gpu --timeout=300 > gpurun_out/pytest_new.log 2>&1; echo "pytest exit $?" > gpurun_out/summary.txt
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.txt
grep -E "^(FAILED|ERROR)|passed|failed|Error|assert" gpurun_out/pytest_new.log | head -40
cat gpurun_out/summary.txt; tail -n 5 gpurun_out/bench.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench.json'))
    print('value',d['value'],'e2e',d['e2e']['value'],'ms/step',d['ms_per_step'],'other',d.get('other_mode'),'parity',d.get('parity'))
    for k in d['kernels']: print(k['kernel'], round(k['avg_ms'],4), k['launches_per_step'], round(k.get('frac',0),3))
except Exception as e: print('bench parse failed', e)
PY
