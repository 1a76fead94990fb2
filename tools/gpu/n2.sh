#!/bin/bash
# 2-GPU frame-sharded bench line (torchrun, NCCL) + the newest e2e tests on GPU 0
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_slim_e2e.py -q -m gpu --timeout=300 -k "graphed or graph" > gpurun_out/pytest_new.log 2>&1; echo "pytest exit $?" > gpurun_out/summary.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 exit $?" >> gpurun_out/summary.txt
grep -E "^(FAILED|ERROR)|passed|failed|Error|assert" gpurun_out/pytest_new.log | head -20
cat gpurun_out/summary.txt; tail -n 5 gpurun_out/bench_n2.err; head -c 900 gpurun_out/bench_n2.json
