#!/bin/bash
# quick GPU check: parity tests + bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout=600 -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" > gpurun_out/summary.txt
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.txt
tail -n 15 gpurun_out/pytest_gpu.log; cat gpurun_out/summary.txt; tail -n 5 gpurun_out/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'ms/step',d['ms_per_step'])
for k in d['kernels']: print(k['kernel'], round(k['avg_ms'],4), k['launches_per_step'], round(k.get('frac',0),3), k.get('tensor',{}).get('frac'))
PY
