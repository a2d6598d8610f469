#!/bin/bash
# Round 2, session 16: ncu --set full of fft_zy_kernel at 1024^3
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fft_zy_kernel -c 1 -o gpurun_out/r2s16_zy --force-overwrite python bench.py --steps 1 --warmup 3 --no-e2e --no-self-check > gpurun_out/r2s16_ncu.log 2>&1
tail -3 gpurun_out/r2s16_ncu.log
ls -la gpurun_out/r2s16_zy.ncu-rep
