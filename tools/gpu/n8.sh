#!/bin/bash
# N-GPU bench line (default 8): one rank per GPU under torchrun, NCCL; export over 10k distinct pairs sharded over the ranks
N=${N:-8}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench n$N exit $?"
tail -3 gpurun_out/bench_n$N.err
python - <<PY
import json
d = json.load(open('gpurun_out/bench_n$N.json'))
print('value', round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), 'ms/step', round(d['ms_per_step'], 3), 'n_gpus', d['n_gpus'])
print(json.dumps(d.get('per_rank'), indent=None)[:2500])
print(json.dumps({k: v for k, v in d.get('export', {}).items() if k in ('pairs', 'triples', 'triple_vs_three_pair_calls')}, indent=1))
print(d['clocks'])
PY
