#!/bin/bash
# Round 2, session 4: sweep coupling experiments (weak polls, mark steps), uncoupled profile, AHEAD=3
mkdir -p gpurun_out
echo "== sweep tests"; timeout 900 python -m pytest tests/test_gpu_sweep.py -m gpu -x -q --timeout 120 > gpurun_out/r2s4_pytest_sweep.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r2s4_pytest_sweep.log
show() {
python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:round(v,3) for k,v in d.get("stage_ms",{}).items() if k in ("zero","deposit")}, {k:round(v["frac"],3) for k,v in d.get("roofline_all",{}).items() if k=="deposit"}, d["config"].get("sweep"))
except Exception as e: print("ERR", e)
PY
}
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-self-check"
run() { name=$1; shift; echo "== $name: $*"; timeout 200 $B "$@" > gpurun_out/r2s4_$name.json 2> gpurun_out/r2s4_$name.err; echo "rc=$?"; show gpurun_out/r2s4_$name.json; }
run memset_default --no-zero-ahead
run memset_strong --no-zero-ahead --sweep-poll-strong
run memset_s1_c6 --no-zero-ahead --sweep-couple-step 1 --sweep-couple 6
run memset_s8_c1 --no-zero-ahead --sweep-couple-step 8 --sweep-couple 1
run memset_s8_c2 --no-zero-ahead --sweep-couple-step 8 --sweep-couple 2
run memset_s16_c1 --no-zero-ahead --sweep-couple-step 16 --sweep-couple 1
run memset_s4_c4 --no-zero-ahead --sweep-couple-step 4 --sweep-couple 4
run memset_c0 --no-zero-ahead --sweep-couple 0
run memset_ry8 --no-zero-ahead --sweep-ry 16
run za_default
run za_w6 --za-window 6
export GENPK_LIB=$PWD/genpk_b200/libgenpk_cuda_a3.so
run a3_memset_default --no-zero-ahead
run a3_memset_c0 --no-zero-ahead --sweep-couple 0
unset GENPK_LIB
echo "== ncu: uncoupled and default coupled (memset)"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"deposit_sweep" -s 3 -c 1 -o gpurun_out/r2s4_prof_c0 -f $B --steps 1 --no-zero-ahead --sweep-couple 0 > gpurun_out/r2s4_ncu1.log 2>&1; echo "rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"deposit_sweep" -s 3 -c 1 -o gpurun_out/r2s4_prof_def -f $B --steps 1 --no-zero-ahead > gpurun_out/r2s4_ncu2.log 2>&1; echo "rc=$?"
