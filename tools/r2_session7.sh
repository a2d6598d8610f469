#!/bin/bash
# Round 2, session 7: does pulling the grid lines into L2 ahead of the reductions help?
mkdir -p gpurun_out
show() {
python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:round(v,3) for k,v in d.get("stage_ms",{}).items() if k in ("zero","deposit")}, {k:round(v["frac"],3) for k,v in d.get("roofline_all",{}).items() if k=="deposit"})
except Exception as e: print("ERR", e)
PY
}
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-self-check"
run() { name=$1; shift; echo "== $name: $*"; timeout 200 $B "$@" > gpurun_out/r2s7_$name.json 2> gpurun_out/r2s7_$name.err; echo "rc=$?"; show gpurun_out/r2s7_$name.json; }
run gp0
run gp1 --sweep-grid-prefetch 1
run gp3 --sweep-grid-prefetch 3
run gp9 --sweep-grid-prefetch 9
run gp1_rx4 --sweep-grid-prefetch 1 --sweep-rx 4
run gp1_rx2 --sweep-grid-prefetch 1 --sweep-rx 2
run gp5_rx4 --sweep-grid-prefetch 5 --sweep-rx 4
echo "== sweep tests with prefetch on"; timeout 600 python -m pytest tests/test_gpu_sweep.py -m gpu -x -q --timeout 120 -k "column_heights or slab" > gpurun_out/r2s7_pytest.log 2>&1; tail -2 gpurun_out/r2s7_pytest.log
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_red_lookup_miss.sum,lts__t_sectors_srcunit_tex_op_red_lookup_hit.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"deposit_sweep" -s 2 -c 1 --csv --log-file gpurun_out/r2s7_gp1_metrics.csv $B --steps 1 --sweep-grid-prefetch 1 > /dev/null 2>&1; grep -E "deposit_sweep" gpurun_out/r2s7_gp1_metrics.csv | awk -F'","' '{print $(NF-2), $(NF)}'
