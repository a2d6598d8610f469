#!/bin/bash
# Session 4: lattice-march deposit parity + timing.
mkdir -p gpurun_out
echo "== pytest march"; timeout 900 python -m pytest tests/test_gpu_march.py -x -q > gpurun_out/pytest_march.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_march.log
echo "== pytest gpu (rest)"; timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_march.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
run() {
  name=$1; shift
  echo "== bench $name: $*"
  timeout 900 python bench.py "$@" --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
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
run c3_march_hint --workload c3 --deposit march --lattice-hint
run c3_march_y16 --workload c3 --deposit march --lattice-hint --march-ry 16 --march-rx 8
run c3_march_y4x16 --workload c3 --deposit march --lattice-hint --march-ry 4 --march-rx 16
run c3_march_y8x1 --workload c3 --deposit march --lattice-hint --march-ry 8 --march-rx 1
run c3_march_y1x1 --workload c3 --deposit march --lattice-hint --march-ry 1 --march-rx 1
run c3_fixed --workload c3 --fixed-point
run c2_auto --workload c2
echo "== ncu full: march (c3)"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"deposit_march" -s 3 -c 1 -o gpurun_out/prof_c3_march -f python bench.py --workload c3 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_c3_march.log 2>&1
echo "ncu full rc=$?"; ls -la gpurun_out/*.ncu-rep
