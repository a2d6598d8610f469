#!/bin/bash
# Round 2, session 19: twiddle powers computed (one lookup per butterfly unit) in all three FFT kernels
mkdir -p gpurun_out
echo "== pytest fftx+fullsize+parity"; timeout 1200 python -m pytest tests/test_gpu_fftx.py tests/test_gpu_fullsize.py tests/test_gpu_parity.py -m gpu -q -x --timeout 600 2>&1 | tail -6
run() { name=$1; shift; timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e "$@" 2>/dev/null | tail -1 > gpurun_out/r2s19_$name.json; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2s19_$name.json").read()); print("$name", round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["stage_ms"].items()}, d.get("cufft_execs_per_step"), (d.get("self_check") or {}).get("pk",{}).get("max_rel_power"))
except Exception as e: print("$name failed", e)
PY
}
run zy1 --fused-zy 1
run zy0 --fused-zy 0
run zy1_lag3 --fused-zy 1 --zy-lag 3
run zy2 --fused-zy 2
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fft_zy_kernel -c 1 -o gpurun_out/r2s19_zy --force-overwrite python bench.py --steps 1 --warmup 3 --no-e2e --no-self-check > gpurun_out/r2s19_ncu.log 2>&1
ls -la gpurun_out/r2s19_zy.ncu-rep
