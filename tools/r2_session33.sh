#!/bin/bash
# Round 2, session 33 (1 GPU): validation of the head -- full GPU suite, smoke, default bench, reference arm
mkdir -p gpurun_out
echo "== pytest gpu (all)"; timeout 1800 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2s33_pytest_gpu.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/r2s33_pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
show() {
python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step","gpu_launches","cufft_execs_per_step")}, {k:round(v,3) for k,v in d.get("stage_ms",{}).items()}, {k:round(v["frac"],3) for k,v in d.get("roofline_all",{}).items()}, "e2e", d.get("e2e") and round(d["e2e"].get("ms_per_step",0),2), d.get("roofline",{}).get("traffic"), d.get("roofline",{}).get("traffic_source"), d.get("cpu_baseline",{}).get("value"), d.get("self_check") and list(d["self_check"]))
except Exception as e: print("ERR", e)
PY
}
echo "== bench reference arm"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2s33_bench_reference.json 2> gpurun_out/r2s33_bench_reference.err; echo "rc=$?"; show gpurun_out/r2s33_bench_reference.json
echo "== bench default"
timeout 600 python bench.py > gpurun_out/r2s33_bench_default.json 2> gpurun_out/r2s33_bench_default.err; echo "rc=$?"; show gpurun_out/r2s33_bench_default.json; tail -3 gpurun_out/r2s33_bench_default.err
