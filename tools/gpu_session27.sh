#!/bin/bash
# Session 27: f8 positions (genpk_deposit_f64, C++ host bigfile path) + full GPU suite on the committed state.
mkdir -p gpurun_out
echo "== pytest gpu (all)"; timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/s27_pytest_gpu.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/s27_pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
