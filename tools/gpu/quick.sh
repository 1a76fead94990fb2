#!/bin/bash
# quick GPU check: parity tests + per-kernel timings (+ optional bench line with QUICK_BENCH=1)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout=600 -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" > gpurun_out/summary.txt
timeout 600 python tools/kbench.py > gpurun_out/kbench.txt 2>&1; echo "kbench exit $?" >> gpurun_out/summary.txt
tail -n 15 gpurun_out/pytest_gpu.log; cat gpurun_out/summary.txt; cat gpurun_out/kbench.txt | tail -20
if [ -n "$QUICK_BENCH" ]; then
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
tail -n 5 gpurun_out/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'ms/step',d['ms_per_step'])
for k in d['kernels']: print(k['kernel'], round(k['avg_ms'],4), k['launches_per_step'], round(k.get('frac',0),3), k.get('tensor',{}).get('frac'))
PY
fi
