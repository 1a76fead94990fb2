#!/bin/bash
# round-2 export pass: new GPU tests (npz stream, compressed / raw-scan export), then a bench with a short export
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_npz_stream.py tests/test_gpu_slim_e2e.py -q -m gpu -x --timeout=600 > gpurun_out/pytest_new.log 2>&1; echo "pytest exit $?" > gpurun_out/summary.txt
tail -5 gpurun_out/pytest_new.log
timeout 1200 python bench.py --steps 20 --warmup 5 --export-pairs ${EXPORT_PAIRS:-1600} --no-other-workloads > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -5 gpurun_out/bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench.json'))
print('value', round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), 'd2h', d['e2e']['d2h_bytes_per_step'], 'raw e2e', round(d['e2e']['uncompressed_d2h']['value'], 1), 'ms/step', round(d['ms_per_step'], 3))
print(json.dumps(d.get('export'), indent=1))
PY
