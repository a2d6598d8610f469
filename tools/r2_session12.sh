#!/bin/bash
# Round 2, session 12: everything that changed since session 9 on one GPU (multi handle emulated with shared devices)
mkdir -p gpurun_out
echo "== pytest gpu (all)"; timeout 1800 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2s12_pytest_gpu.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/r2s12_pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
