#!/bin/bash
# Round 2, session 36: compute-sanitizer on the scatter variant of the fused (y,z) kernel (slabs emulated on one GPU, 256^3)
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_fftx.py -m gpu -q -x --timeout 800 -k "scatter_y_pass_emulated_on_one_gpu and 256-2" > gpurun_out/r2s36_sanitizer_racecheck_scatter.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/r2s36_sanitizer_racecheck_scatter.log
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_fftx.py -m gpu -q -x --timeout 800 -k "scatter_y_pass_emulated_on_one_gpu and 256-2" > gpurun_out/r2s36_sanitizer_memcheck_scatter.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r2s36_sanitizer_memcheck_scatter.log
