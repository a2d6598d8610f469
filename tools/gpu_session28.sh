#!/bin/bash
# Session 28: default bench line on the final commit of the round (after the deposit chunk-loop refactor for f8 positions).
mkdir -p gpurun_out
timeout 400 python bench.py > gpurun_out/s28_bench_default.json 2> gpurun_out/s28_bench_default.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/s28_bench_default.json").read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("value","ms_per_step","gpu_launches")}, {k:round(v,3) for k,v in d["stage_ms"].items()}, "e2e", d["e2e"]["ms_per_step"], d["clocks"], d["cpu_baseline"]["value"])
PY
tail -3 gpurun_out/s28_bench_default.err
