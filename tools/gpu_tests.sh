#!/bin/bash
# All GPU tests with per-test timeouts (a hung kernel must not eat the box budget).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 180 "$@" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
