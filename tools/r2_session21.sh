#!/bin/bash
# Round 2, session 21: fft_zy_kernel with rfft twiddles from constants only (hoisting reverted), lag 2 / 3, twice each
mkdir -p gpurun_out
run() { name=$1; shift; timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e "$@" 2>/dev/null | tail -1 > gpurun_out/r2s21_$name.json; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2s21_$name.json").read()); print("$name", round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["stage_ms"].items()}, d["clocks"]["sm_mhz"])
except Exception as e: print("$name failed", e)
PY
}
run lag3a
run lag2a --zy-lag 2
run lag3b
run lag2b --zy-lag 2
run zy0 --fused-zy 0
