#!/bin/bash
# e2e tests + bench with and without the fused lookup
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_slim_e2e.py tests/test_gpu_corr.py -q -m gpu --timeout=600 -x > gpurun_out/pytest_e2e.log 2>&1; echo "e2e tests exit $?" > gpurun_out/summary.txt
tail -n 5 gpurun_out/pytest_e2e.log
timeout 900 python bench.py --steps 20 --warmup 6 --no-cpu-baseline --fused-lookup > gpurun_out/bench_fused.json 2> gpurun_out/bench_fused.err; echo "bench fused exit $?" >> gpurun_out/summary.txt
timeout 900 python bench.py --steps 20 --warmup 6 --no-cpu-baseline  > gpurun_out/bench_unfused.json 2> gpurun_out/bench_unfused.err; echo "bench unfused exit $?" >> gpurun_out/summary.txt
python - <<'PY'
import json
for n in ("fused", "unfused"):
    try:
        d = json.load(open('gpurun_out/bench_%s.json' % n))
    except Exception as e:
        print(n, "no json", e); continue
    print(n, 'value', round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), 'ms/step', round(d['ms_per_step'], 3), 'other', round(d['other_mode']['value'], 1))
    for k in d['kernels']:
        if 'lookup' in k['kernel'] or 'gemm' in k['kernel'] or 'pillar' in k['kernel']:
            print('   ', k['kernel'], round(k['avg_ms'], 4), k['launches_per_step'], round(k.get('frac', 0), 3))
PY
cat gpurun_out/summary.txt; tail -3 gpurun_out/bench_fused.err
