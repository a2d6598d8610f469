#!/bin/bash
# occupancy / block-shape sweep of the march kernel on C3
mkdir -p gpurun_out
run() {
  name=$1; shift
  timeout 200 python bench.py "$@" --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$name.json").read().strip().splitlines()[-1])
    print("$name", {k:round(v,3) for k,v in d["stage_ms"].items()}, {k:round(v["frac"],3) for k,v in d["roofline_all"].items()})
except Exception as e: print("$name ERR", e)
PY
  tail -2 gpurun_out/bench_$name.err
}
for b in 3 4 2; do
  echo "=== MINB=$b"
  (cd genpk_b200/csrc && touch deposit_march.cu && make -j8 EXTRA_NVFLAGS="-DGENPK_MARCH_MINB=$b" > /dev/null 2>&1)
  run c3_b${b}_y8x8 --workload c3 --deposit march --lattice-hint
  run c3_b${b}_y16x16 --workload c3 --deposit march --lattice-hint --march-ry 16 --march-rx 16
  run c3_b${b}_y4x32 --workload c3 --deposit march --lattice-hint --march-ry 4 --march-rx 32
done
(cd genpk_b200/csrc && touch deposit_march.cu && make -j8 > /dev/null 2>&1)
