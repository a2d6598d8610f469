#!/bin/bash
# Round 2, session 5: sweep kernel as independent tasks (rx planes x ry rows x segment), L2 prefetch
mkdir -p gpurun_out
echo "== sweep tests"; timeout 900 python -m pytest tests/test_gpu_sweep.py -m gpu -x -q --timeout 120 > gpurun_out/r2s5_pytest_sweep.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r2s5_pytest_sweep.log
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
run() { name=$1; shift; echo "== $name: $*"; timeout 200 $B "$@" > gpurun_out/r2s5_$name.json 2> gpurun_out/r2s5_$name.err; echo "rc=$?"; show gpurun_out/r2s5_$name.json; }
run task_rx8_ry8
run task_rx8_ry12 --sweep-ry 12
run task_rx8_ry16 --sweep-ry 16
run task_rx16_ry8 --sweep-rx 16
run task_rx4_ry8 --sweep-rx 4
run task_rx32_ry8 --sweep-rx 32
run persistent_c0 --sweep-rx 0 --no-zero-ahead --sweep-couple 0
run persistent_def --sweep-rx 0 --no-zero-ahead
run persistent_za --sweep-rx 0
run task_fixed --fixed-point
export GENPK_LIB=$PWD/genpk_b200/libgenpk_cuda_t256.so
run t256_task_rx8_ry8
run t256_task_rx8_ry12 --sweep-ry 12
unset GENPK_LIB
echo "== ncu: task mode"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"deposit_sweep" -s 3 -c 1 -o gpurun_out/r2s5_prof_task -f $B --steps 1 > gpurun_out/r2s5_ncu1.log 2>&1; echo "rc=$?"
