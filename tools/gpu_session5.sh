#!/bin/bash
mkdir -p gpurun_out
echo "== pytest march"; timeout 300 python -m pytest tests/test_gpu_march.py -x -q --timeout 60 > gpurun_out/pytest_march.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_march.log
run() {
  name=$1; shift
  echo "== bench $name: $*"
  timeout 200 python bench.py "$@" --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  echo "rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$name.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","stage_ms","gpu_launches")}, {k:round(v["frac"],3) for k,v in d["roofline_all"].items()}, d["config"].get("order_probe"))
except Exception as e: print("ERR", e)
PY
  tail -3 gpurun_out/bench_$name.err
}
run c3_auto --workload c3
run c3_march_y16 --workload c3 --deposit march --lattice-hint --march-ry 16 --march-rx 16
run c2_auto --workload c2
run c3_fixed --workload c3 --fixed-point
echo "== pytest all gpu"; timeout 600 python -m pytest tests -m gpu -x -q --timeout 120 > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_gpu.log
exit 0
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"deposit_march" -s 3 -c 1 -o gpurun_out/prof_c3_march -f python bench.py --workload c3 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_c3_march.log 2>&1
echo "ncu full rc=$?"; ls -la gpurun_out/*.ncu-rep
