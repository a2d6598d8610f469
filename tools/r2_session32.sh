#!/bin/bash
# Round 2, session 32: fused x pass with the two-pass plan (GENPK_OPT_FUSED_XPASS = 4)
mkdir -p gpurun_out
echo "== pytest fused_vs_unfused"; timeout 600 python -m pytest tests/test_gpu_fftx.py -m gpu -q -x --timeout 300 -k "fused_vs_unfused" 2>&1 | tail -4
run() { name=$1; shift; timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline "$@" 2>/dev/null | tail -1 > gpurun_out/r2s32_$name.json; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2s32_$name.json").read()); print("$name", round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["stage_ms"].items()}, d["clocks"]["sm_mhz"], (d.get("self_check") or {}).get("pk",{}).get("max_rel_power"))
except Exception as e: print("$name failed", e)
PY
}
run wide
run twopass --xpass-two-pass
run twopass2 --xpass-two-pass
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fftx_power2_kernel -c 1 -o gpurun_out/r2s32_x2 --force-overwrite python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-self-check --xpass-two-pass > gpurun_out/r2s32_ncu.log 2>&1; ls -la gpurun_out/r2s32_x2.ncu-rep
