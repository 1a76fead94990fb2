#!/bin/bash
# round-2 GPU pass: bench (all lines), reference arm, ncu launch list of one step, ncu --set full of the hot kernels
mkdir -p gpurun_out
timeout 1200 python bench.py --steps 20 --warmup 5 --export-pairs ${EXPORT_PAIRS:-2400} > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" > gpurun_out/summary.txt
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref exit $?" >> gpurun_out/summary.txt
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --profile-one-step --no-cuda-graph --warmup 3 > gpurun_out/ncu_list.log 2>&1; echo "ncu list exit $?" >> gpurun_out/summary.txt
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'k_corr_gemm|k_pillar_nhwc|k_corr_lookup|k_lookup_conv_tf32|k_feat_pack|k_point_keys|k_scan_local|k_scan_global|k_rank_scatter|k_decode_bev|k_raft_output|k_ctx_split|k_add_bias_relu' -c 40 -f -o gpurun_out/prof python bench.py --profile-one-step --no-cuda-graph --warmup 3 > gpurun_out/ncu_full.log 2>&1; echo "ncu full exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench.json'))
print('value', round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), 'ms/step', round(d['ms_per_step'], 3), 'warmup', d['warmup'], 'launches', d['gpu_launches'])
r = d['roofline']
print('roofline', r['stage'], r['kernel'], round(r['frac'], 3), r['traffic'])
for k, v in r['stages'].items():
    print('  ', k, v['kernel'], round(v['avg_launch_ms'], 4), round(v['frac'], 3), v.get('tensor', {}).get('frac'), v.get('stage_incl_prep', {}).get('frac'))
print('other', json.dumps(d.get('other_workloads'))[:1500])
print('cpu', d.get('cpu_baseline', {}).get('value'), d.get('parity'))
PY
