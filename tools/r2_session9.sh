#!/bin/bash
# Round 2, session 9: full GPU suite, smoke, default bench (+reference arm), C2 line, launch list, sanitizer
mkdir -p gpurun_out
echo "== pytest gpu (all)"; timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2s9_pytest_gpu.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/r2s9_pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
show() {
python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","ms_per_step","gpu_launches")}, {k:round(v,3) for k,v in d.get("stage_ms",{}).items()}, {k:round(v["frac"],3) for k,v in d.get("roofline_all",{}).items()}, "e2e", d.get("e2e") and round(d["e2e"].get("ms_per_step",0),2), d.get("clocks"), d.get("cpu_baseline",{}).get("value"), d.get("self_check") and list(d["self_check"]))
except Exception as e: print("ERR", e)
PY
}
echo "== bench reference arm"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2s9_bench_reference.json 2> gpurun_out/r2s9_bench_reference.err; echo "rc=$?"; show gpurun_out/r2s9_bench_reference.json
echo "== bench default"
timeout 600 python bench.py > gpurun_out/r2s9_bench_default.json 2> gpurun_out/r2s9_bench_default.err; echo "rc=$?"; show gpurun_out/r2s9_bench_default.json; tail -3 gpurun_out/r2s9_bench_default.err
echo "== bench c2"
timeout 300 python bench.py --workload c2 --no-cpu-baseline > gpurun_out/r2s9_bench_c2.json 2> gpurun_out/r2s9_bench_c2.err; echo "rc=$?"; show gpurun_out/r2s9_bench_c2.json; tail -3 gpurun_out/r2s9_bench_c2.err
echo "== ncu launch list (c3)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2s9_launches_c3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2s9_ncu_list.log 2>&1; echo "rc=$?"
echo "== compute-sanitizer (256^3 through march + fft_cols + fftx_power, and the sweep kernel)"
cat > /tmp/san.py <<'PY'
import sys, torch, numpy as np
sys.path.insert(0, ".")
import genpk_b200 as gp
from genpk_b200 import api
n_side = dims = 256
n = n_side ** 3
d = torch.empty(3 * n, dtype=torch.float32, device="cuda")
api.synth_particles_dev(api.SYNTH_CLUSTERED, 42, n_side, 0, n, 1000.0, dims, d.data_ptr()); torch.cuda.synchronize()
for mode in (api.DEPOSIT_MARCH, api.DEPOSIT_SWEEP):
    with gp.Context(dims) as ctx:
        ctx.set_deposit_mode(mode)
        ctx.grid_zero(); ctx.deposit_dev(d.data_ptr(), n, 0, 1.0, 1000.0)
        p, c, k = ctx.fft_power(dims, float(n), float(n)); ctx.synchronize()
        print("mode", mode, "counts", int(c.astype(np.int64).sum()), "P[10]", p[10])
PY
timeout 900 compute-sanitizer --tool memcheck python /tmp/san.py > gpurun_out/r2s9_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/r2s9_sanitizer_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck python /tmp/san.py > gpurun_out/r2s9_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/r2s9_sanitizer_racecheck.log
