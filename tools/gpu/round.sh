#!/bin/bash
# GPU round: parity tests, bench, ncu launch list of one step, full ncu capture of the hot kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout=600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" > gpurun_out/summary.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref exit $?" >> gpurun_out/summary.txt
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --profile-one-step --no-cuda-graph --warmup 3 > gpurun_out/ncu_list.log 2>&1; echo "ncu list exit $?" >> gpurun_out/summary.txt
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'k_corr_gemm|k_pillar_nhwc|k_corr_lookup|k_feat_pack|k_decode_bev|k_decode_points|k_raft_output|k_decode_aggr' -c 20 -f -o gpurun_out/prof python bench.py --profile-one-step --no-cuda-graph --warmup 3 > gpurun_out/ncu_full.log 2>&1; echo "ncu full exit $?" >> gpurun_out/summary.txt
tail -n 3 gpurun_out/pytest_gpu.log; cat gpurun_out/summary.txt; cat gpurun_out/bench.json | head -c 3000
