#!/bin/bash
# Session 3: binning v2 (mirror rows) parity + timing.
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
for w in "c2 auto cached" "c3 auto cached" "c3 auto fused"; do
  set -- $w
  echo "== bench $1 $2 $3"
  timeout 900 python bench.py --workload $1 --deposit $2 --power $3 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_$1_$2_$3.json 2> gpurun_out/bench_$1_$2_$3.err
  echo "rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$1_$2_$3.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","stage_ms","gpu_launches")}, {k:round(v["frac"],3) for k,v in d["roofline_all"].items()})
except Exception as e: print("ERR", e)
PY
  tail -3 gpurun_out/bench_$1_$2_$3.err
done
