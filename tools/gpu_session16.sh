#!/bin/bash
# Session 16: fused x-pass v4 (cp.async staging, half-tile complex exchange, forward-only walk, two histograms).
mkdir -p gpurun_out
echo "== pytest fftx + march"; timeout 900 python -m pytest tests/test_gpu_fftx.py tests/test_gpu_march.py -q > gpurun_out/s16_pytest_fftx.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/s16_pytest_fftx.log
echo "== A/B"; timeout 600 python tools/ab_fft.py > gpurun_out/s16_ab_fft.txt 2> gpurun_out/s16_ab_fft.err; echo "rc=$?"; cat gpurun_out/s16_ab_fft.txt; tail -3 gpurun_out/s16_ab_fft.err
echo "== ncu full: fftx_power_kernel default tile (c3)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fftx_power" -s 2 -c 1 -o gpurun_out/s16_prof_fftx -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s16_ncu_full.log 2>&1; echo "rc=$?"; ls -la gpurun_out/s15*.ncu-rep
