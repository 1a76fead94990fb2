#!/bin/bash
# decoder parity + bench lines of the other workloads (N: nuScenes-sized sparse pillars, A: AV2-sized 920x920 grid)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_decoder.py -q -m gpu --timeout=300 > gpurun_out/pytest_new.log 2>&1; echo "pytest exit $?" > gpurun_out/summary.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench K exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --workload N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_N.json 2> gpurun_out/bench_N.err; echo "bench N exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --workload A --batch 4 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_A.json 2> gpurun_out/bench_A.err; echo "bench A exit $?" >> gpurun_out/summary.txt
grep -E "^(FAILED|ERROR)|passed|failed|Error|assert" gpurun_out/pytest_new.log | head -30
cat gpurun_out/summary.txt; tail -n 3 gpurun_out/bench.err gpurun_out/bench_N.err gpurun_out/bench_A.err
python - <<'PY'
import json
for f in ('bench','bench_N','bench_A'):
    try:
        d=json.load(open('gpurun_out/%s.json'%f))
        print(f,'value',d['value'],'e2e',d['e2e']['value'],'ms/step',d['ms_per_step'],'roofline',d['roofline']['kernel'],round(d['roofline']['frac'],3))
        for k in d['kernels'][:10]: print('   ',k['kernel'], round(k['avg_ms'],4), k['launches_per_step'], round(k.get('frac',0),3), k.get('tensor',{}).get('frac'))
    except Exception as e: print(f,'parse failed', e)
PY
