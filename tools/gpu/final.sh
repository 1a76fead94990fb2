#!/bin/bash
# last check of a round: whole GPU suite, smoke, one default bench line (+ the reference arm)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout=900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu.log | tail -10
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/smoke.log
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref exit $?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench.json'))
print('value', round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), 'ms/step', round(d['ms_per_step'], 3), 'launches', d['gpu_launches'], 'other', d['other_mode'])
r = d['roofline']
print('roofline', r['stage'], r['kernel'], round(r['frac'], 3), r['traffic'])
for k, v in r['stages'].items():
    print('  ', k, v['kernel'], round(v['avg_launch_ms'], 4), round(v['frac'], 3), v.get('tensor', {}).get('frac'), v.get('stage_incl_prep', {}).get('frac'))
e = d['export']
print('export pairs', e['pairs']['pairs_per_s'], 'triples', e['triples']['samples_per_s'], e['triples']['pairs_per_s'], 'zlib', e.get('pairs_host_zlib_writer', {}).get('pairs_per_s'))
print({k: (v.get('value'), v.get('e2e')) for k, v in d['other_workloads'].items()})
print('cpu', d.get('cpu_baseline', {}).get('value'), d.get('parity'), d['clocks'])
PY
