#!/bin/bash
# Round 2, session 15: fft_zy_kernel (z rows + y columns in one persistent kernel) -- tests, then A/B bench
mkdir -p gpurun_out
echo "== pytest fftx"; timeout 900 python -m pytest tests/test_gpu_fftx.py -m gpu -q -x --timeout 300 2>&1 | tail -15
run() { name=$1; shift; timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e "$@" 2>/dev/null | tail -1 > gpurun_out/r2s15_$name.json; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2s15_$name.json").read()); print("$name", round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["stage_ms"].items()}, d.get("cufft_execs_per_step"), (d.get("self_check") or {}).get("pk",{}).get("max_rel_power"))
except Exception as e: print("$name failed", e)
PY
}
run zy1 --fused-zy 1
run zy0 --fused-zy 0
run zy2 --fused-zy 2
run zy1_lag1 --fused-zy 1 --zy-lag 1
run zy1_lag4 --fused-zy 1 --zy-lag 4
run zy1_lag8 --fused-zy 1 --zy-lag 8
