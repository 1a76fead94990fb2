#!/bin/bash
# A/B of the fused lookup generations inside the bench step
mkdir -p gpurun_out
for g in 3 4; do
  SLIMB200_LOOKUP_CONV_GEN=$g timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-other-workloads > gpurun_out/bench_g$g.json 2> gpurun_out/bench_g$g.err
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-other-workloads  > gpurun_out/bench_unfused.json 2> gpurun_out/bench_unfused.err
python - <<'PY'
import json
for n in ("g3", "g4", "unfused"):
    try:
        d = json.load(open('gpurun_out/bench_%s.json' % n))
    except Exception as e:
        print(n, "no json", e); continue
    st = d['roofline']['stages']['lookup']
    print(n, 'value', round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), 'ms/step', round(d['ms_per_step'], 3), 'last-mode', round(d['other_mode']['value'], 1),
          '| lookup', st['kernel'], round(st['avg_launch_ms'], 4), round(st['frac'], 3))
PY
