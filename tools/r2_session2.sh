#!/bin/bash
# Round 2, session 2: sweep kernel v4 (128-thread CTAs, cp.async ring, dedicated zero CTAs, coupling).
mkdir -p gpurun_out
echo "== sweep tests"; timeout 900 python -m pytest tests/test_gpu_sweep.py -m gpu -x -q --timeout 120 > gpurun_out/r2s2_pytest_sweep.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/r2s2_pytest_sweep.log
show() {
python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("ms_per_step",)}, {k:round(v,3) for k,v in d.get("stage_ms",{}).items() if k in ("zero","deposit","fft","binning")}, {k:round(v["frac"],3) for k,v in d.get("roofline_all",{}).items()}, d["config"].get("sweep"))
except Exception as e: print("ERR", e)
PY
}
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-self-check"
run() { name=$1; shift; echo "== $name: $*"; timeout 200 $B "$@" > gpurun_out/r2s2_$name.json 2> gpurun_out/r2s2_$name.err; echo "rc=$?"; show gpurun_out/r2s2_$name.json; }
run c3_memset_c6 --no-zero-ahead
GENPK_X=1 run c3_memset_c0 --no-zero-ahead --sweep-couple 0
run c3_memset_c2 --no-zero-ahead --sweep-couple 2
run c3_memset_c16 --no-zero-ahead --sweep-couple 16
run c3_za_auto
run c3_za_w6 --za-window 6
run c3_za_w4 --za-window 4
run c3_za_w8_s4 --za-window 8 --za-slack 4
run c3_za_w6_z24 --za-window 6 --za-zero-ctas 24
run c3_za_w6_z96 --za-window 6 --za-zero-ctas 96
run c3_za_w6_c3 --za-window 6 --sweep-couple 3
run c3_za_w6_c12 --za-window 6 --sweep-couple 12
run c3_za_fixed --fixed-point
echo "== ncu full: sweep memset / za (c3)"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"deposit_sweep" -s 3 -c 1 -o gpurun_out/r2s2_prof_sweep_memset -f $B --steps 1 --no-zero-ahead > gpurun_out/r2s2_ncu1.log 2>&1; echo "rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"deposit_sweep" -s 3 -c 1 -o gpurun_out/r2s2_prof_sweep_za -f $B --steps 1 > gpurun_out/r2s2_ncu2.log 2>&1; echo "rc=$?"
ls -la gpurun_out/*.ncu-rep
