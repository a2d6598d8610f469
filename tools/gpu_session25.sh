#!/bin/bash
# Session 25: closing run of the committed state on one GPU: full GPU suite, smoke, A/B of the FFT variants, default bench line.
mkdir -p gpurun_out
echo "== pytest gpu (all)"; timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/s25_pytest_gpu.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/s25_pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== A/B"; timeout 600 python tools/ab_fft.py > gpurun_out/s25_ab_fft.txt 2> gpurun_out/s25_ab_fft.err; echo "rc=$?"; cat gpurun_out/s25_ab_fft.txt; tail -3 gpurun_out/s25_ab_fft.err
echo "== bench default"
timeout 600 python bench.py > gpurun_out/s25_bench_default.json 2> gpurun_out/s25_bench_default.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/s25_bench_default.json").read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("value","ms_per_step","gpu_launches")}, {k:round(v,3) for k,v in d["stage_ms"].items()}, d["roofline"], "e2e", d["e2e"], d["clocks"])
PY
tail -3 gpurun_out/s25_bench_default.err
