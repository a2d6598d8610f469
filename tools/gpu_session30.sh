#!/bin/bash
# Session 30: early cp.async issue kept in the x pass only (the y pass queues behind its own store burst).
mkdir -p gpurun_out
echo "== pytest fftx + slab"; timeout 300 python -m pytest tests/test_gpu_fftx.py tests/test_gpu_slab.py -q > gpurun_out/s30_pytest.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/s30_pytest.log
echo "== A/B"; timeout 200 python tools/ab_fft.py > gpurun_out/s30_ab_fft.txt 2> gpurun_out/s30_ab_fft.err; echo "rc=$?"; cat gpurun_out/s30_ab_fft.txt; tail -3 gpurun_out/s30_ab_fft.err
