#!/bin/bash
# Session 17: own in-place y pass (fft_cols_kernel) + fused x pass v5 (preloaded bin walk).
mkdir -p gpurun_out
echo "== pytest fftx + march + slab"; timeout 900 python -m pytest tests/test_gpu_fftx.py tests/test_gpu_march.py tests/test_gpu_slab.py -q -x > gpurun_out/s17_pytest_fftx.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/s17_pytest_fftx.log
echo "== A/B"; timeout 600 python tools/ab_fft.py > gpurun_out/s17_ab_fft.txt 2> gpurun_out/s17_ab_fft.err; echo "rc=$?"; cat gpurun_out/s17_ab_fft.txt; tail -3 gpurun_out/s17_ab_fft.err
echo "== ncu full: fftx_power_kernel + fft_cols_kernel (c3)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fftx_power|fft_cols" -s 4 -c 2 -o gpurun_out/s17_prof_fft -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s17_ncu_full.log 2>&1; echo "rc=$?"; ls -la gpurun_out/s17*.ncu-rep
