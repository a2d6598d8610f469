#!/bin/bash
# Round 2, session 6: sweep (task mode) after the instruction cuts vs march
mkdir -p gpurun_out
echo "== sweep tests"; timeout 900 python -m pytest tests/test_gpu_sweep.py -m gpu -x -q --timeout 120 > gpurun_out/r2s6_pytest_sweep.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r2s6_pytest_sweep.log
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
run() { name=$1; shift; echo "== $name: $*"; timeout 200 $B "$@" > gpurun_out/r2s6_$name.json 2> gpurun_out/r2s6_$name.err; echo "rc=$?"; show gpurun_out/r2s6_$name.json; }
run task_rx8_ry8
run task_rx8_ry12 --sweep-ry 12
run task_rx16_ry12 --sweep-ry 12 --sweep-rx 16
run march --no-sweep
run task_fixed --fixed-point
run march_fixed --no-sweep --fixed-point
echo "== ncu launch list (metrics only)"
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"deposit_sweep|deposit_march" -s 2 -c 2 --csv --log-file gpurun_out/r2s6_sweep_metrics.csv $B --steps 1 > /dev/null 2>&1; tail -3 gpurun_out/r2s6_sweep_metrics.csv | cut -c1-400
