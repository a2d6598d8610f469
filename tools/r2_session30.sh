#!/bin/bash
# Round 2, session 30: head start of the leading half in fftx_power_halves_kernel
mkdir -p gpurun_out
run() { name=$1; shift; timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline "$@" 2>/dev/null | tail -1 > gpurun_out/r2s30_$name.json; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2s30_$name.json").read()); print("$name", round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["stage_ms"].items()}, d["clocks"]["sm_mhz"])
except Exception as e: print("$name failed", e)
PY
}
for d in 0 1000 2000 5000; do GENPK_XHALVES_DELAY_NS=$d run halves_$d --xpass-halves; done
run wide
