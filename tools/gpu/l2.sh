#!/bin/bash
# realistic-cache view of the glue / gather kernels inside one bench step: ncu WITHOUT cache flushes between kernels
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --cache-control none --clock-control none \
  --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_bytes.sum \
  -k regex:'k_in_stats|k_in_apply|k_corr_lookup|k_nhwc_pack|k_gru_gate|k_add_relu|k_raft_output|k_decode_bev|k_decode_aggr|k_kabsch_moments|k_decode_points' \
  --csv --log-file gpurun_out/l2_view.csv python bench.py --profile-one-step --no-cuda-graph --warmup 3 > gpurun_out/ncu_l2.log 2>&1; echo "ncu l2 exit $?"
tail -2 gpurun_out/ncu_l2.log
