#!/bin/bash
# ncu launch list (durations) of ONE resident bench step; extra args go to bench.py
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --profile-one-step --warmup 3 "$@" > gpurun_out/ncu_list.log 2>&1; echo "ncu list exit $?"
tail -3 gpurun_out/ncu_list.log
