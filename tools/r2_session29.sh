#!/bin/bash
# Round 2, session 29: fused x pass as two out-of-step halves of 256 threads on one 8192-mode tile (GENPK_OPT_FUSED_XPASS = 3)
mkdir -p gpurun_out
echo "== pytest fused_vs_unfused"; timeout 600 python -m pytest tests/test_gpu_fftx.py -m gpu -q -x --timeout 300 -k "fused_vs_unfused" 2>&1 | tail -4
run() { name=$1; shift; timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline "$@" 2>/dev/null | tail -1 > gpurun_out/r2s29_$name.json; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2s29_$name.json").read()); print("$name", round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["stage_ms"].items()}, d["clocks"]["sm_mhz"], (d.get("self_check") or {}).get("pk",{}).get("max_rel_power"))
except Exception as e: print("$name failed", e)
PY
}
run wide
run halves --xpass-halves
run halves2 --xpass-halves
