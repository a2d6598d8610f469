#!/bin/bash
# Round 2, session 34: march kernel built for 4 (64 registers) and 2 (92 registers) resident CTAs per SM instead of 3 (80)
mkdir -p gpurun_out
run() { name=$1; shift; timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline "$@" 2>/dev/null | tail -1 > gpurun_out/r2s34_$name.json; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2s34_$name.json").read()); print("$name", round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["stage_ms"].items()}, d["clocks"]["sm_mhz"], list((d.get("self_check") or {}).keys()))
except Exception as e: print("$name failed", e)
PY
}
run base
GENPK_LIB=$PWD/genpk_b200/libgenpk_cuda_minb4.so run minb4
GENPK_LIB=$PWD/genpk_b200/libgenpk_cuda_minb2.so run minb2
GENPK_LIB=$PWD/genpk_b200/libgenpk_cuda_minb4.so run minb4b
run base2
