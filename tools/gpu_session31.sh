#!/bin/bash
# Session 31: ncu launch list of the final commit (C3, one GPU).
mkdir -p gpurun_out
timeout 45 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/s31_launches_c3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s31_ncu_list.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/s31_ncu_list.log | cut -c1-300
