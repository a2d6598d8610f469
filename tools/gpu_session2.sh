#!/bin/bash
# Session 2: parity after deposit v1 / binning v1, bulk-reduce microbench, bench, ncu captures.
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
echo "== microbench"; timeout 600 tools/microbench > gpurun_out/microbench2.txt 2>&1; grep -E "TMA|copy|CIC" gpurun_out/microbench2.txt
for w in "c2 auto cached" "c2 direct cached" "c2 sorted fused" "c3 auto cached" "c3 sorted fused"; do
  set -- $w
  echo "== bench $1 $2 $3"
  timeout 900 python bench.py --workload $1 --deposit $2 --power $3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$1_$2_$3.json 2> gpurun_out/bench_$1_$2_$3.err
  echo "rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$1_$2_$3.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","stage_ms","gpu_launches")}, "e2e", d["e2e"] and d["e2e"]["ms_per_step"], {k:round(v["frac"],3) for k,v in d["roofline_all"].items()})
except Exception as e: print("ERR", e)
PY
  tail -3 gpurun_out/bench_$1_$2_$3.err
done
echo "== ncu launch list (c3)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_c3.csv python bench.py --workload c3 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_c3.log 2>&1
echo "ncu rc=$?"; tail -8 gpurun_out/launches_c3.csv | cut -c1-220
echo "== ncu full: deposit + binning (c3)"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"deposit_direct|bin_power" -s 6 -c 4 -o gpurun_out/prof_c3 -f python bench.py --workload c3 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_c3.log 2>&1
echo "ncu full rc=$?"; ls -la gpurun_out/*.ncu-rep
