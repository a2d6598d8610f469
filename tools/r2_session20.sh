#!/bin/bash
# Round 2, session 20: fft_zy_kernel -- rfft twiddles from constants, pass-1 twiddles fetched ahead of the tile wait, lag sweep
mkdir -p gpurun_out
echo "== pytest fftx"; timeout 1200 python -m pytest tests/test_gpu_fftx.py -m gpu -q -x --timeout 600 2>&1 | tail -4
run() { name=$1; shift; timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e "$@" 2>/dev/null | tail -1 > gpurun_out/r2s20_$name.json; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2s20_$name.json").read()); print("$name", round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["stage_ms"].items()}, d.get("cufft_execs_per_step"), (d.get("self_check") or {}).get("pk",{}).get("max_rel_power"))
except Exception as e: print("$name failed", e)
PY
}
run lag3
run lag2 --zy-lag 2
run lag4 --zy-lag 4
run lag5 --zy-lag 5
run lag6 --zy-lag 6
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fft_zy_kernel -c 1 -o gpurun_out/r2s20_zy --force-overwrite python bench.py --steps 1 --warmup 3 --no-e2e --no-self-check > gpurun_out/r2s20_ncu.log 2>&1
ls -la gpurun_out/r2s20_zy.ncu-rep
